import sys, os
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
x = torch.randn(4,16,1024,1024,device="cuda"); y = torch.randn(4,16,1024,1024,device="cuda")
fb.semiring_matmul("logaddexp","add",x,y); torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fb.semiring_matmul("logaddexp","add",x,y); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print("mode", os.environ.get("FB2_LM_MODE"), "64 x 1024^3 log-matmul:", ms, "ms", 64*2*1024**3/ms/1e9, "TFLOP/s", flush=True)
