"""Where the end-to-end (host buffers) time of the headline shape goes: device-resident pass vs fb2_hmm_forward_f32_host at several T."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
import funsor_b200 as fb
from funsor_b200 import _lib
lib = _lib.load()
S, B = 1024, 16
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(S, S, device="cuda", generator=g).log_softmax(-1)
init = torch.randn(B, S, device="cuda", generator=g).log_softmax(-1)
for T in (4096, 16384, 65536):
    obs = torch.randn(B, T, S, device="cuda", generator=g)
    for _ in range(2): z = fb.hmm_forward(init, A, obs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): z = fb.hmm_forward(init, A, obs)
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) / 3 * 1e3
    h_init, h_A = init.cpu().pin_memory(), A.cpu().pin_memory()
    h_obs = torch.empty(obs.shape, dtype=torch.float32, pin_memory=True); h_obs.copy_(obs)
    h_out = torch.empty(B, dtype=torch.float32, pin_memory=True)
    def step():
        _lib.check(lib.fb2_hmm_forward_f32_host(0, h_init.data_ptr(), S, h_A.data_ptr(), 0, 0, h_obs.data_ptr(), B, T, S, h_out.data_ptr()))
    step(); step()
    t0 = time.perf_counter()
    for _ in range(3): step()
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    print("T=%d: device-resident %.2f ms   host buffers %.2f ms   gap %.2f ms (%.1f %%)   H2D bytes %.0f MB" % (
        T, dev_ms, e2e_ms, e2e_ms - dev_ms, 100 * (e2e_ms - dev_ms) / dev_ms, obs.numel() * 4 / 1e6), flush=True)
    del obs, h_obs
