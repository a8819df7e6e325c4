import sys, time, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import funsor_b200 as fb
from oracle import semiring_ref as R
torch.manual_seed(0)
for (B,T,S) in [(16,40,1024),(5,33,65),(1,64,700),(40,25,257)]:
    init = torch.randn(B,S).log_softmax(-1); A = torch.randn(S,S).log_softmax(-1); obs = torch.randn(B,T,S)
    z = fb.hmm_forward(init.cuda(), A.cuda(), obs.cuda()).cpu().numpy()
    w = R.hmm_forward(R.LOG, init.double().numpy(), A.double().numpy(), obs.double().numpy())
    print(B,T,S, "maxabs", np.abs(z-w).max(), "rel", (np.abs(z-w)/(1+np.abs(w))).max(), flush=True)
B,T,S = 16, 4096, 1024
g = torch.Generator(device="cuda").manual_seed(0)
init = torch.randn(B,S,device="cuda",generator=g).log_softmax(-1); A = torch.randn(S,S,device="cuda",generator=g).log_softmax(-1); obs = torch.randn(B,T,S,device="cuda",generator=g)
for name, env in (("flow", None), ("generic", "1")):
    if env: os.environ["FB2_DISABLE_FLOW"] = env
    z = fb.hmm_forward(init, A, obs); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = fb.hmm_forward(init, A, obs); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(name, "T=4096 ms", ms, "us/step", ms*1e3/T, z[:3].tolist(), flush=True)

os.environ.pop("FB2_DISABLE_FLOW", None)
os.environ["FB2_FLOW_PROF"] = "1"
z = fb.hmm_forward(init, A, obs); torch.cuda.synchronize()
os.environ.pop("FB2_FLOW_PROF")
for (B2,T2,S2) in [(16,8192,512),(16,8192,256),(64,4096,1024)]:
    init2 = torch.randn(B2,S2,device="cuda").log_softmax(-1); A2 = torch.randn(S2,S2,device="cuda").log_softmax(-1); obs2 = torch.randn(B2,T2,S2,device="cuda")
    z = fb.hmm_forward(init2, A2, obs2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = fb.hmm_forward(init2, A2, obs2); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("flow", (B2,T2,S2), "ms", ms, "us/step", ms*1e3/T2, flush=True)
