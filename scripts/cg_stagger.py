import sys, os, json, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from funsor_b200.ops import hmm_chunk_summary_gemm
    S, T = int(os.environ.get("CG_S", 512)), 16384
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
    obs = torch.randn(1, T, S, device="cuda", generator=g)
    hmm_chunk_summary_gemm(A, obs); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): hmm_chunk_summary_gemm(A, obs)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"S": S, "stagger": os.environ.get("FB2_CG_STAGGER"), "mode": os.environ.get("FB2_CG_MODE", "0"), "us_per_step": e0.elapsed_time(e1) / 3 * 1e3 / T}))
else:
    for st, m in (("0", "0"), ("8000", "0"), ("16000", "0"), ("24000", "0"), ("0", "4"), ("0", "1"), ("0", "2"), ("0", "7")):
        env = dict(os.environ, FB2_CG_STAGGER=st, FB2_CG_MODE=m)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=120).stdout.strip())
