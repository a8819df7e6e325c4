import sys, os
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
B,T,S = 16, 4096, 1024
g = torch.Generator(device="cuda").manual_seed(0)
init = torch.randn(B,S,device="cuda",generator=g).log_softmax(-1); A = torch.randn(S,S,device="cuda",generator=g).log_softmax(-1); obs = torch.randn(B,T,S,device="cuda",generator=g)
for i in range(3):
    z = fb.hmm_forward(init, A, obs); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): z = fb.hmm_forward(init, A, obs)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("FB2_FLOW_MODE"), os.environ.get("FB2_FLOW_LEGACY"), "us/step", e0.elapsed_time(e1)/3*1e3/T, z[:2].tolist(), flush=True)
