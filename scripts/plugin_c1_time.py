"""C1 (T=100, S=16) and a few other sizes evaluated by funsor itself on CUDA: stock rules vs the plugin's kernel branch."""
import sys, os, json, time
from collections import OrderedDict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_boot
funsor = ref_boot.boot("float32")
import funsor.ops as ops
from funsor.domains import Bint
from funsor.sum_product import sequential_sum_product
from funsor.tensor import Tensor
from funsor.terms import Variable
from funsor_b200 import funsor_plugin
funsor_plugin.register()

def run(B, T, S, fast):
    funsor_plugin.set_hooks()
    if not fast:
        funsor_plugin.set_hooks(fast_ok=lambda *d: False)
    g = torch.Generator(device="cuda").manual_seed(0)
    data = torch.randn(B, T, S, S, device="cuda", generator=g)
    def once():
        f = Tensor(data.clone(), OrderedDict(b=Bint[B], time=Bint[T], prev=Bint[S], curr=Bint[S]))
        return sequential_sum_product(ops.logaddexp, ops.add, f, Variable("time", Bint[T]), {"prev": "curr"}).data
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n): r = once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r

out = {}
for (B, T, S) in [(1, 100, 16), (16, 100, 16), (1, 1000, 64), (4, 256, 256), (2, 64, 1024)]:
    ms_stock, r0 = run(B, T, S, False)
    ms_fast, r1 = run(B, T, S, True)
    out["B%d_T%d_S%d" % (B, T, S)] = {"stock_ms": ms_stock, "plugin_ms": ms_fast, "maxdiff": float((r0 - r1).abs().max())}
print(json.dumps(out, indent=1))
