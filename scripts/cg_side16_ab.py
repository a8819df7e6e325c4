"""chain_gemm_tc2: 3xTF32 (default) vs TF32 + bf16 side products (FB2_CG_MODE=8), alternating in one process."""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
from funsor_b200.ops import hmm_chunk_summary_gemm
for S, T in ((512, 32768), (1024, 8192)):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device="cuda", generator=g)
    def timeit(n=3):
        hmm_chunk_summary_gemm(A, obs, block_steps=16384); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): hmm_chunk_summary_gemm(A, obs, block_steps=16384)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3 / T
    res = {"0": [], "8": []}
    for rep in range(4):
        for mode in ("0", "8"):
            os.environ["FB2_CG_MODE"] = mode
            res[mode].append(timeit())
    print("S=%d T=%d us/step: 3xTF32 %s   TF32 + bf16 side %s" % (S, T, ["%.3f" % x for x in res["0"]], ["%.3f" % x for x in res["8"]]), flush=True)
