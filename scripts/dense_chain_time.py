import torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
g = torch.Generator(device="cuda").manual_seed(0)
Bd, Td, Sd = 16, 256, 1024
trans = torch.randn(Bd, Td, Sd, Sd, device="cuda", generator=g)
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mode in ("1", "0"):
    os.environ["FB2_CHAIN_GEMM"] = mode
    ms = timeit(lambda: fb.sequential_sum_product("logaddexp", "add", trans))
    print("FB2_CHAIN_GEMM", mode, "ms", ms, "useful TFLOP/s", 2.0 * Bd * (Td - 1) * Sd ** 3 / (ms * 1e-3) / 1e12)
