"""Timing of the other scope rows (C3 Viterbi, C5 marginals, C4 Kalman, chunk summary) at reduced T."""
import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
import funsor_b200.gaussian as fg
dev = "cuda"
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device=dev).manual_seed(0)
S, T = 4096, 1024
A = torch.randn(S, S, device=dev, generator=g); obs = torch.randn(1, T, S, device=dev, generator=g); init = torch.randn(1, S, device=dev, generator=g)
ms = timeit(lambda: fb.hmm_viterbi(init, A, obs)); print("viterbi S=4096 B=1 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
ms = timeit(lambda: fb.hmm_forward(init, A, obs, sum_op="max")); print("maxplus fwd S=4096 B=1 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
S, T = 512, 4096
A = torch.randn(S, S, device=dev, generator=g).log_softmax(-1); obs = torch.randn(1, T, S, device=dev, generator=g); init = torch.randn(1, S, device=dev, generator=g).log_softmax(-1)
ms = timeit(lambda: fb.hmm_marginals(init, A, obs, want_xi=False)); print("marginals (no xi, flow) S=512 B=1 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
ms = timeit(lambda: fb.hmm_forward(init, A, obs)); print("log fwd (flow) S=512 B=1 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
ms = timeit(lambda: fb.hmm_forward(init, A, obs, return_alpha=True)); print("log fwd+alpha (generic) S=512 B=1 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
Tc = 256
ms = timeit(lambda: fb.hmm_chunk_summary(A, obs[:, :Tc].contiguous())); print("chunk summary (generic) S=512 T=%d: %.2f ms, %.2f us/step" % (Tc, ms, ms * 1e3 / Tc), flush=True)
Tc2 = 2048
ms = timeit(lambda: fb.hmm_chunk_summary(A, obs[:, :Tc2].contiguous(), return_offsets=True, method="scan")); print("chunk summary (dataflow scan) S=512 T=%d: %.2f ms, %.2f us/step" % (Tc2, ms, ms * 1e3 / Tc2), flush=True)
ms = timeit(lambda: fb.hmm_chunk_summary(A, obs, return_offsets=True, method="tree")); print("chunk summary (tcgen05 tree) S=512 T=%d: %.2f ms, %.2f us/step" % (T, ms, ms * 1e3 / T), flush=True)
Bk, Tk, D, O = 64, 8192, 32, 16
F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=dev, generator=g))[0]
H = torch.randn(Bk, D, O, device=dev, generator=g) / D ** 0.5
eye = lambda n: torch.eye(n, device=dev).expand(Bk, n, n).contiguous()
y = torch.randn(Bk, Tk, O, device=dev, generator=g)
w, P, s = fg.kalman_elements_sqrt(F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O), y)
ms0 = timeit(lambda: fg.sqrt_to_info(w, P, s))
Lam, eta, c = fg.sqrt_to_info(w, P, s)
ms = timeit(lambda: fg.gaussian_sequential_sum_product(Lam, eta, c, validate=False))
print("kalman D=32 B=64 T=%d: sqrt_to_info %.2f ms, scan %.2f ms -> %.3g steps/s" % (Tk, ms0, ms, Bk * Tk / (ms * 1e-3)), flush=True)

kargs = (F, torch.zeros(Bk, D, device=dev), eye(D), H, torch.zeros(Bk, O, device=dev), eye(O), y)
ms = timeit(lambda: fg.kalman_chain(*kargs, validate=False))
print("kalman homogeneous path (elements + eta + scan) D=32 B=64 T=%d: %.2f ms -> %.3g steps/s" % (Tk, ms, Bk * Tk / (ms * 1e-3)), flush=True)
