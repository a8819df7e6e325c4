"""C3 shape (Viterbi, S = 4096, B = 1): register-resident rows on / off; paths must be identical."""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
for S, T in ((4096, 2048), (3000, 2048), (2500, 2048)):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(S, S, device="cuda", generator=g); obs = torch.randn(1, T, S, device="cuda", generator=g); init = torch.randn(1, S, device="cuda", generator=g)
    res = {}
    for rr in ("0", "1"):
        os.environ["FB2_VITERBI_REGROWS"] = rr
        for _ in range(2):
            best, path = fb.hmm_viterbi(init, A, obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            best, path = fb.hmm_viterbi(init, A, obs)
        e1.record(); torch.cuda.synchronize()
        res[rr] = (e0.elapsed_time(e1) / 3, best.clone(), path.clone())
    print("S=%d T=%d: streamed + smem rows %.2f us/step   + register rows %.2f us/step   identical value %s path %s" % (
        S, T, res["0"][0] * 1e3 / T, res["1"][0] * 1e3 / T, torch.equal(res["0"][1], res["1"][1]), torch.equal(res["0"][2], res["1"][2])), flush=True)
