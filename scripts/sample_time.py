import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")
for (B, T, S, N) in [(1, 8192, 512, 64), (4, 4096, 128, 256), (1, 2048, 2048, 16)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    u = torch.rand(B, N, T + 1, device=dev, generator=g)
    fb.hmm_posterior_sample(init, A, obs, uniforms=u); torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); fb.hmm_forward(init, A, obs, return_alpha=True); e1.record(); fb.hmm_posterior_sample(init, A, obs, uniforms=u); e2.record(); torch.cuda.synchronize()
    fwd, tot = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print("B=%d T=%d S=%d N=%d: forward+alpha %.1f ms, forward+sampling %.1f ms -> sampling %.2f us per step" % (B, T, S, N, fwd, tot, (tot - fwd) * 1e3 / T), flush=True)
