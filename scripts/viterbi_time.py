import sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
S, T = 4096, 512
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(S, S, device="cuda", generator=g); obs = torch.randn(1, T, S, device="cuda", generator=g); init = torch.randn(1, S, device="cuda", generator=g)
for _ in range(2):
    best, path = fb.hmm_viterbi(init, A, obs); torch.cuda.synchronize()
print("viterbi ok", float(best[0]))
