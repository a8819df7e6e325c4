"""Time the chunk summary (the cost of breaking the dependency chain for time sharding): log-domain tree vs linear-domain GEMM tree."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
from funsor_b200.ops import hmm_chunk_summary_gemm

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
obs = torch.randn(1, T, S, device="cuda", generator=g)
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {"S": S, "T": T}
for name, fn in (("tree", lambda: fb.hmm_chunk_summary(A, obs, return_offsets=True, method="tree")),
                 ("gemm_4096", lambda: hmm_chunk_summary_gemm(A, obs, block_steps=4096)),
                 ("gemm_16384", lambda: hmm_chunk_summary_gemm(A, obs, block_steps=16384))):
    ms = timeit(fn)
    out[name] = {"ms": ms, "us_per_step": ms * 1e3 / T, "useful_tflops": 2.0 * S ** 3 * (T - 1) / (ms * 1e-3) / 1e12}
print(json.dumps(out))
