"""C3 shape (Viterbi, S = 4096, B = 1): lines in flight per thread in the streaming kernel (8 / 16 / 32); paths must be identical.
The template variants behind FB2_VITERBI_DEEP were removed from the library after this measurement (DESIGN.md section 4.2 records
the numbers); with the current library the three columns time the same kernel."""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
for S, T in ((4096, 2048), (3000, 2048)):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(S, S, device="cuda", generator=g); obs = torch.randn(1, T, S, device="cuda", generator=g); init = torch.randn(1, S, device="cuda", generator=g)
    res = {}
    for mode in ("0", "1", "2"):
        os.environ["FB2_VITERBI_DEEP"] = mode
        for _ in range(2):
            best, path = fb.hmm_viterbi(init, A, obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            best, path = fb.hmm_viterbi(init, A, obs)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 3, best.clone(), path.clone())
    print("S=%d T=%d us/step: 16 lines %.2f   32 lines %.2f   8 lines %.2f   identical %s" % (
        S, T, res["0"][0] * 1e3 / T, res["1"][0] * 1e3 / T, res["2"][0] * 1e3 / T,
        all(torch.equal(res["0"][i], res[m][i]) for m in ("1", "2") for i in (1, 2))), flush=True)
