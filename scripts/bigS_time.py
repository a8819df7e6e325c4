import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb
dev = torch.device("cuda")
for (B, T, S) in [(1, 256, 2048), (16, 256, 2048), (1, 128, 4096), (16, 64, 4096)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    for name, fn in (("log forward", lambda: fb.hmm_forward(init, A, obs)), ("viterbi", lambda: fb.hmm_viterbi(init, A, obs))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("B=%d T=%d S=%d %s: %.2f ms, %.1f us/step" % (B, T, S, name, ms, ms * 1e3 / T), flush=True)
for (B, T, S) in [(1, 128, 2048), (4, 64, 4096)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = torch.randn(B, T, S, device=dev, generator=g)
    for stream in ("1", "0"):
        os.environ["FB2_DISABLE_STREAM"] = stream
        fn = lambda: fb.hmm_marginals(init, A, obs, want_xi=False)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        print("B=%d T=%d S=%d marginals (%s): %.2f ms, %.1f us/step" % (B, T, S, "generic" if stream == "1" else "stream", e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / T), flush=True)
os.environ.pop("FB2_DISABLE_STREAM")
