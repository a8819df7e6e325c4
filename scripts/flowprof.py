import sys, os
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
B,T,S = 16, 4096, 1024
g = torch.Generator(device="cuda").manual_seed(0)
init = torch.randn(B,S,device="cuda",generator=g).log_softmax(-1); A = torch.randn(S,S,device="cuda",generator=g).log_softmax(-1); obs = torch.randn(B,T,S,device="cuda",generator=g)
z = fb.hmm_forward(init, A, obs); torch.cuda.synchronize()
os.environ["FB2_FLOW_PROF"] = "1"
z = fb.hmm_forward(init, A, obs); torch.cuda.synchronize()
