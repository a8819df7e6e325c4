import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from funsor_b200.ops import hmm_chunk_summary_gemm
S, T = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
obs = torch.randn(1, T, S, device="cuda", generator=g)
hmm_chunk_summary_gemm(A, obs, block_steps=4096)
torch.cuda.synchronize()
