"""Per-step bias of the fp32 paths against a float64 recurrence (linear domain, rescaled) on one long chain."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
from funsor_b200 import distributed as fd

dev = torch.device("cuda")
T, S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device=dev); g.manual_seed(7)
A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
init = torch.log_softmax(torch.randn(1, S, device=dev, generator=g), -1)
obs = torch.randn(1, T, S, device=dev, generator=g)
P = A.double().exp()
E = obs[0].double().exp()
v = init[0].double().exp()
acc = 0.0
accs = torch.zeros((), dtype=torch.float64, device=dev)
for t in range(T):
    v = (v @ P) * E[t]
    s = v.sum()
    v = v / s
    accs += s.log()
ref = accs.item()
z = fb.hmm_forward(init, A, obs)
print("float64 ref", ref)
print("flow fp32 out", z.item(), "diff", z.item() - ref, "per step", (z.item() - ref) / T)
summ = fb.hmm_chunk_summary(A, obs, return_offsets=True)
val, _ = fd.hmm_forward_time_sharded(init, A, obs, summary_fn=lambda a, o: summ)
print("tree summary", val.item(), "diff", val.item() - ref, "per step", (val.item() - ref) / T)
summ = fb.hmm_chunk_summary(A, obs[:, :8192].contiguous(), return_offsets=True, method="scan")
