import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import funsor_b200 as fb
def ref(x, y):
    return torch.logsumexp(x.double().unsqueeze(-1) + y.double().unsqueeze(-3), -2)
torch.manual_seed(0)
for shape in [(1,1,128,128,128),(2,3,128,64,128),(1,2,200,100,130),(1,1,1024,1024,1024),(3,1,64,33,64),(1,1,257,260,129)]:
    n0,n1,I,K,J = shape
    x = torch.randn(n0,n1,I,K,device="cuda")*3; y = torch.randn(n0,n1,K,J,device="cuda")*3
    got = fb.semiring_matmul("logaddexp","add",x,y)
    want = ref(x,y) if I*K*J <= 2**27 else None
    if want is None:
        os.environ["FB2_DISABLE_TC_MATMUL"]="1"; want = fb.semiring_matmul("logaddexp","add",x,y).double(); os.environ.pop("FB2_DISABLE_TC_MATMUL")
    err = (got.double()-want).abs().max().item()
    print(shape, "max abs err", err, "launches", fb.launch_count(reset=True), flush=True)
x = torch.randn(4,16,1024,1024,device="cuda"); y = torch.randn(4,16,1024,1024,device="cuda")
for name, env in (("tc", None), ("simt", "1")):
    if env: os.environ["FB2_DISABLE_TC_MATMUL"]=env
    fb.semiring_matmul("logaddexp","add",x,y); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fb.semiring_matmul("logaddexp","add",x,y); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); print(name, "64 x 1024^3 log-matmul:", ms, "ms", 64*2*1024**3/ms/1e9, "TFLOP/s", flush=True)
