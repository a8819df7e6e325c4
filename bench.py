#!/usr/bin/env python
"""Benchmark of the hot path on the configuration BASELINE.json's metric is quoted on.

Workload (config.workload = "c2"): batched discrete-HMM forward log-likelihood in the log semiring,
T=65536, S=1024, batch=16 per GPU, fp32, factored inputs A[S,S] + obs[B,T,S] (the
dense trans[B,T,S,S] of the reference API would be 4.4 TB).  One step = one pass of the hot path over
the whole batch.  Unit of work = one (t, batch, prev, curr) semiring multiply-add, so
value = N * T*S^2*B / seconds.  Multi-GPU: batch sharding, no data-path collective ("weak" scaling:
16 chains per GPU).  For N > 1 the same line also carries (extra.multi_gpu) the STRONG-scaling number of C2
(16 chains in total, split over the ranks) and the time-sharded forward-backward of BASELINE config 5
(T = 1,048,576, S = 512, one all-gather of chunk summaries) with its one-GPU time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hmm_forward_logsemiring_TxS2xB_per_s"
UNIT = "semiring_madd/s"
WORKLOADS = {
    "c2": dict(T=65536, S=1024, B=16),
    "c2_small": dict(T=2048, S=1024, B=16),  # debugging only
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(T, S, B, device, seed):
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    A = torch.randn(S, S, device=device, generator=g).log_softmax(-1)  # normalised like hmm.py:92
    init = torch.randn(B, S, device=device, generator=g).log_softmax(-1)
    obs = torch.randn(B, T, S, device=device, generator=g)
    return init, A, obs


def cpu_reference_sample(S, B, T_sample, repeats=1):
    """Time the reference's own data flow (oracle/cpu_port.py) on all host threads on a bounded sample."""
    import torch

    from oracle import cpu_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T_sample, S, generator=g)
    cpu_port.hmm_log_prob(init, A, obs[:, :2])  # warm-up (thread pool, allocator)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_port.hmm_log_prob(init, A, obs)
        best = min(best, time.perf_counter() - t0)
    return dict(value=T_sample * S * S * B / best, seconds=best, cores=cores, kind="port",
                sample="T={} of the workload's S={}, B={} (dense trans+obs, pairwise log-matmul tree, torch CPU fp32, {} threads)".format(T_sample, S, B, cores))


_FUNSOR = {}


def funsor_reference_sample(S, B, T_sample):
    """The UNMODIFIED reference (funsor, installed under baseline/_ref; absent third-party deps stood in for by oracle/shims) evaluating
    DiscreteHMM.log_prob's expression (pyro/hmm.py:128-137) on torch CPU tensors with all host threads, on a T-sample of the workload.
    Returns None when the package is not present."""
    import torch

    from oracle import ref_boot

    if not ref_boot.available():
        return None
    if not _FUNSOR:
        _FUNSOR["funsor"] = ref_boot.boot("float32")
    from collections import OrderedDict

    import funsor.ops as ops
    from funsor.domains import Bint
    from funsor.sum_product import sequential_sum_product
    from funsor.tensor import Tensor
    from funsor.terms import Variable

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T_sample, S, generator=g)

    def log_prob(T):
        fA = Tensor(A.clone(), OrderedDict([("state", Bint[S]), ("state(time=1)", Bint[S])]))
        fobs = Tensor(obs[:, :T].clone(), OrderedDict([("b", Bint[B]), ("time", Bint[T]), ("state(time=1)", Bint[S])]))
        finit = Tensor(init.clone(), OrderedDict([("b", Bint[B]), ("state", Bint[S])]))
        result = fA + fobs  # hmm.py:130
        result = sequential_sum_product(ops.logaddexp, ops.add, result, Variable("time", Bint[T]), {"state": "state(time=1)"})
        result = finit + result.reduce(ops.logaddexp, "state(time=1)")
        return result.reduce(ops.logaddexp, "state").data

    log_prob(2)  # warm-up: dispatch caches, thread pool
    t0 = time.perf_counter()
    z = log_prob(T_sample)
    sec = time.perf_counter() - t0
    assert bool(torch.isfinite(z).all())
    return dict(value=T_sample * S * S * B / sec, seconds=sec, cores=cores, kind="reference",
                sample="funsor 0.4.8 (unmodified, baseline/_ref): hmm.py:128-137 expression on T={} of the workload's S={}, B={} "
                       "(torch CPU fp32, {} threads; log-semiring einsum backend = the reference's in-tree funsor.einsum.numpy_log, "
                       "pyro's torch_log being absent)".format(T_sample, S, B, cores))


def cpu_vector_recurrence_sample(S, B, T_sample):
    """The SAME algorithm as the GPU path (exp-domain vector recurrence, O(T S^2 B)) on the host with torch CPU fp32: separates the
    algorithmic part of the GPU/CPU ratio (the reference executes S-fold more arithmetic: matrix-chain products) from the hardware part."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(S, S, generator=g).log_softmax(-1)
    init = torch.randn(B, S, generator=g).log_softmax(-1)
    obs = torch.randn(B, T_sample, S, generator=g)
    P = A.exp()

    def run(T):
        v = init.exp()
        acc = torch.zeros(B)
        E = obs[:, :T].exp()
        for t in range(T):
            v = (v @ P) * E[:, t]
            s = v.sum(-1, keepdim=True)
            v = v / s
            acc = acc + s[:, 0].log()
        return acc

    run(4)
    t0 = time.perf_counter()
    run(T_sample)
    sec = time.perf_counter() - t0
    return dict(value=T_sample * S * S * B / sec, seconds=sec, cores=cores,
                sample="T={} steps of v <- (v P) * exp(obs_t), S={}, B={}, torch CPU fp32, {} threads".format(T_sample, S, B, cores))


def reference_sample(S, B, T_sample):
    r = funsor_reference_sample(S, B, T_sample)
    return r if r is not None else cpu_reference_sample(S, B, T_sample)


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    T_sample = args.ref_T
    samples = []
    for i in range(args.warmup + args.steps):
        r = reference_sample(wl["S"], wl["B"], T_sample)
        if i >= args.warmup:
            samples.append(r)
    sec = sum(r["seconds"] for r in samples) / len(samples)
    value = T_sample * wl["S"] ** 2 * wl["B"] / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "T": wl["T"], "S": wl["S"], "B_per_gpu": wl["B"], "layout": "factored A[S,S] + obs[B,T,S]",
                   "parallelism": "host threads", "reference_sample_T": T_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": samples[0]["cores"], "kind": samples[0]["kind"], "sample": samples[0]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-T", type=int, default=16, help="time steps in the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours" and args.workload == "c2":
        pass  # allowed for debugging, but the driver always passes W >= 3

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    import funsor_b200 as fb
    from funsor_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    T, S, B = wl["T"], wl["S"], wl["B"]
    init, A, obs = make_inputs(T, S, B, device, seed=rank)
    units_per_step_rank = T * S * S * B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing (value) ----------------
    for _ in range(args.warmup):
        z = fb.hmm_forward(init, A, obs)
    barrier()
    fb.launch_count(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            z = fb.hmm_forward(init, A, obs)
            ev[i + 1].record()
        barrier()
    launches = fb.launch_count()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    kernel_ms = sum(step_ms) / len(step_ms)  # a step = 2 small memsets + colmax + the dataflow forward kernel + finalise;
    # the forward kernel is > 99.9 % of it (profiles/: ncu launch list), so the step time is used as its duration
    value = world * units_per_step_rank * args.steps / (total_ms * 1e-3)
    assert bool(torch.isfinite(z).all()), "non-finite log-likelihood"

    # ---------------- end-to-end through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        lib = _lib.load()
        h_init, h_A = init.cpu().pin_memory(), A.cpu().pin_memory()
        h_obs = torch.empty(obs.shape, dtype=torch.float32, pin_memory=True)
        h_obs.copy_(obs)
        h_out = torch.empty(B, dtype=torch.float32, pin_memory=True)

        def e2e_step():
            _lib.check(lib.fb2_hmm_forward_f32_host(0, h_init.data_ptr(), S, h_A.data_ptr(), 0, 0, h_obs.data_ptr(), B, T, S, h_out.data_ptr()))

        n_e2e = max(2, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert torch.allclose(h_out.to(device), z, rtol=1e-5, atol=1e-2), "e2e result differs from the device-resident result"
        e2e = {"value": world * units_per_step_rank * n_e2e / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(h_obs.numel() * 4 + h_A.numel() * 4 + h_init.numel() * 4), "d2h_bytes_per_step": int(B * 4),
               "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e}
        try:  # the same bytes as one plain pinned -> device copy: when this is close to ms_per_step the end-to-end number is the PCIe link, not the kernel
            d_tmp = torch.empty_like(obs)
            d_tmp.copy_(h_obs, non_blocking=True)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            d_tmp.copy_(h_obs, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            e2e["h2d_copy_alone_ms"] = c0.elapsed_time(c1)
            e2e["h2d_copy_alone_gbs"] = h_obs.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
            del d_tmp
        except Exception as e:
            e2e["h2d_copy_alone_ms"] = repr(e)
        lib.fb2_host_arena_release()
        del h_obs

    multi = None
    if world > 1 and not args.no_extra:
        del obs, init, z
        torch.cuda.empty_cache()
        try:
            multi = multi_gpu_rows(device, rank, world, barrier, max_over_ranks)
        except Exception as e:
            multi = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    flops = 2.0 * units_per_step_rank
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    hbm_bytes = (B * T * S + S * S + B * S) * 4
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if args.workload == "c2" and os.path.exists(tpath):  # dram bytes of one launch of the dominant kernel, from the committed ncu capture
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["tflops"],
                "traffic": traffic, "kernel": "hmm_flow_forward<8,2,8,true>", "kernel_ms": kernel_ms, "peak_source": peaks["src"],
                "algorithmic_flops_per_launch": flops, "us_per_time_step": kernel_ms * 1e3 / T,
                "latency_floor_us": 1.7, "latency_floor_source": "scripts/microbench/pingpong.cu: one store -> remote-SM load hand-over through L2 = 936 cycles "
                "(0.48 us), DSMEM hop ~215 cycles + 17-21 B/cycle; a step of this decomposition needs one L2 hop + one DSMEM reduce + the contraction + two short "
                "ALU chains (DESIGN.md section 3)", "achieved_over_floor": kernel_ms * 1e3 / T / 1.7,
                "note": "T dependent steps of a [16x1024]x[1024x1024] product: latency-bound (one L2 hand-over is 936 cycles "
                        "measured, plus a DSMEM reduce per step), DESIGN.md section 3; dram traffic = algorithmic (profiles/)",
                "hbm": {"achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                        "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": hbm_bytes}}
    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline and the secondary rows are N=1 items
        r = reference_sample(S, B, args.ref_T)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        try:  # like-for-like: the GPU path's own algorithm on the host cores (what part of the ratio is the algorithm)
            v = cpu_vector_recurrence_sample(S, B, 256)
            cpu["same_algorithm_on_cpu"] = {"value": v["value"], "unit": UNIT, "cores": v["cores"], "sample": v["sample"],
                                            "note": "the reference arm executes matrix-chain products (S-fold more arithmetic per unit, SURVEY 8(d)); "
                                                    "this row runs the vector recurrence the GPU kernel runs"}
        except Exception as e:
            cpu["same_algorithm_on_cpu"] = {"error": repr(e)}

    extra = {}
    if not args.no_extra and world == 1:
        try:
            extra = extra_measurements(device)
        except Exception as e:  # extras never invalidate the headline line
            extra = {"error": repr(e)}
    if multi is not None:
        extra["multi_gpu"] = multi

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "T": T, "S": S, "B_per_gpu": B, "layout": "factored A[S,S] + obs[B,T,S]",
                   "parallelism": "batch-sharded x{}".format(world), "l2": "inputs (4.3 GB obs) larger than L2, no flush needed"},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_rows(device, rank, world, barrier, max_over_ranks):
    """The two multi-GPU designs of SURVEY 8(e) on the BASELINE shapes, timed on the device (CUDA events, max over ranks).

    strong_c2        C2 with 16 chains IN TOTAL, batch-sharded (16 / N chains per rank, no collective).
    c5_time_sharded  forward-backward marginals of ONE chain of T = 1,048,576 steps, S = 512, time-sharded: chunk summary per rank
                     (tensor-core tree) -> ONE all-gather -> folds -> local forward-backward; rank 0 then runs the whole chain
                     alone (the one-GPU time the speed-up is quoted against) and checks logZ / gamma."""
    import torch
    import torch.distributed as dist

    import funsor_b200 as fb
    from funsor_b200 import distributed as fd

    out = {}

    def timed(fn, reps=2):
        best = float("inf")
        for i in range(reps + 1):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            barrier()
            t = max_over_ranks(e0.elapsed_time(e1))
            if i > 0:
                best = min(best, t)
        return best, r

    # ---- strong scaling of C2 by batch sharding
    T, S, Btot = WORKLOADS["c2"]["T"], WORKLOADS["c2"]["S"], WORKLOADS["c2"]["B"]
    lo, hi = fd.shard_bounds(Btot, rank, world)
    g = torch.Generator(device=device).manual_seed(100)
    A = torch.randn(S, S, device=device, generator=g).log_softmax(-1)
    g.manual_seed(200 + rank)
    init = torch.randn(max(hi - lo, 1), S, device=device, generator=g).log_softmax(-1)
    obs = torch.randn(max(hi - lo, 1), T, S, device=device, generator=g)
    ms, _ = timed(lambda: fb.hmm_forward(init, A, obs))
    out["strong_c2"] = {"scaling": "strong", "workload": "c2 with B=16 in total", "chains_per_rank": hi - lo, "ms": ms,
                        "units_per_s": T * S * S * Btot / (ms * 1e-3), "us_per_time_step": ms * 1e3 / T,
                        "note": "every rank still walks all T dependent steps; the step latency does not depend on the number of chains "
                                "in a 16-chain group (DESIGN.md section 3), so batch sharding below 16 chains per GPU cannot speed one pass up"}
    del obs, init
    torch.cuda.empty_cache()

    # ---- time-sharded forward-backward (BASELINE config 5)
    T5, S5 = 1 << 20, 512
    Tr = T5 // world

    def chunk_obs(r):
        gg = torch.Generator(device=device).manual_seed(1000 + r)
        return torch.randn(1, Tr, S5, device=device, generator=gg)

    g.manual_seed(7)
    A5 = torch.randn(S5, S5, device=device, generator=g).log_softmax(-1)
    init5 = torch.randn(1, S5, device=device, generator=g).log_softmax(-1)
    obs5 = chunk_obs(rank)
    parts = {}

    def sharded():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        summ = fb.hmm_chunk_summary(A5, obs5, return_offsets=True)
        e[1].record()
        parts["ev"] = e
        return fd.hmm_marginals_time_sharded(init5, A5, obs5, summary_fn=lambda A_, o_: summ)

    ms5, (logZ, gamma) = timed(sharded)
    torch.cuda.synchronize()
    ms_summary = max_over_ranks(parts["ev"][0].elapsed_time(parts["ev"][1]))
    row = {"scaling": "strong", "workload": "c5 forward-backward marginals, time-sharded", "T": T5, "S": S5, "B": 1, "ms": ms5,
           "ms_chunk_summary": ms_summary, "us_per_time_step": ms5 * 1e3 / T5, "units_per_s": 3.0 * T5 * S5 * S5 / (ms5 * 1e-3),
           "collective": "one all_gather of [S,S] float64 summaries ({} MiB in total)".format(world * S5 * S5 * 8 >> 20)}
    if rank == 0:  # the same chain on ONE GPU
        full = torch.cat([chunk_obs(r) for r in range(world)], 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fb.hmm_marginals(init5, A5, full[:, :4096], want_xi=False)
        e0.record()
        z1, g1, _ = fb.hmm_marginals(init5, A5, full, want_xi=False)
        e1.record()
        torch.cuda.synchronize()
        row["ms_one_gpu"] = e0.elapsed_time(e1)
        row["speedup_vs_one_gpu"] = row["ms_one_gpu"] / ms5
        row["logZ_rel_err_vs_one_gpu"] = abs(z1[0].item() - logZ[0].item()) / abs(z1[0].item())
        row["gamma_max_abs_err_chunk0"] = (g1[:, :Tr] - gamma).abs().max().item()
    dist.barrier()
    out["c5_time_sharded"] = row
    return out


def extra_measurements(device):
    """Secondary numbers of the same metric family (not the headline): Kalman steps/s and dense chain product."""
    import torch

    import funsor_b200 as fb
    import funsor_b200.gaussian as fg

    out = {}
    peaks = measured_peaks()
    g = torch.Generator(device=device).manual_seed(0)
    l2 = measure_l2_bandwidth(device)
    out["l2_bandwidth_probe"] = l2

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    # Kalman (C4 at full size): D=32, O=16, B=64, T=100000, homogeneous dynamics; factor construction + eta + parallel scan
    Bk, Tk, D, O = 64, 100000, 32, 16
    F = 0.9 * torch.linalg.qr(torch.randn(Bk, D, D, device=device, generator=g))[0]
    H = torch.randn(Bk, D, O, device=device, generator=g) / D ** 0.5
    eye = lambda n: torch.eye(n, device=device).expand(Bk, n, n).contiguous()  # noqa: E731
    y = torch.randn(Bk, Tk, O, device=device, generator=g)
    kargs = (F, torch.zeros(Bk, D, device=device), eye(D), H, torch.zeros(Bk, O, device=device), eye(O), y)
    sec = timeit(lambda: fg.kalman_chain(*kargs, validate=False))
    kalman_bytes = (y.numel() + Bk * (2 * D) * (2 * D + 1) + Bk) * 4.0  # read y once, write one (Lam, eta, c) per chain
    out["kalman_scan"] = {"steps_per_s": Bk * Tk / sec, "D": D, "O": O, "B": Bk, "T": Tk, "ms": sec * 1e3,
                          "kernel": "gaussian_kalman_constants + gaussian_homog_levels_cta<32> (precision operators per tree level) + gaussian_kalman_tile_tc<12> "
                                    "(per tile of 32 steps: elements from y and the first 5 tree levels as 3xTF32 mma.sync GEMMs against the shared operators, "
                                    "16 tiles per warp) + gaussian_homog_upper<32> (levels 5..16 in one launch)",
                          "roofline": {"bound": "hbm", "achieved": kalman_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                       "frac": kalman_bytes / sec / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": kalman_bytes,
                                       "note": "algorithmic bytes = the observations y[B,T,O] read once + the [B] results; the tile kernel (half of the time) "
                                               "runs on the tensor pipe (ncu: tensor pipe 69 % active, DRAM read = y), the rest is the serial chain of per-level "
                                               "precision operators (0.18 ms, one CTA per chain) and the upper levels of the tree (one launch)"},
                          "fp32": {"mac_per_step": 1280 + 2600, "useful_tflops": 2.0 * (1280 + 2600) * Bk * Tk / sec / 1e12,
                                   "note": "1280 multiply-adds per element (u = wo + y Pr, eta = a0 + Po u) + ~2600 per pair contraction (triangular L^-1 z, U'v), "
                                           "issued as 3xTF32 mma.sync (three tensor instructions per useful product)"},
                          "note": "BASELINE config C4 at full size; the general (time-varying) scan runs gaussian_pair_combine_warp<32> at 3.6e7 pair contractions/s"}
    del y, kargs
    # dense chain product (literal sequential_sum_product semantics at the C2 state count): trans[16,256,1024,1024] = 17 GB,
    # tcgen05 3xTF32 log-matmul per tree level
    Bd, Td, Sd = 16, 256, 1024
    trans = torch.randn(Bd, Td, Sd, Sd, device=device, generator=g)
    sec = timeit(lambda: fb.sequential_sum_product("logaddexp", "add", trans), n=2)
    tf = 2.0 * Bd * (Td - 1) * Sd ** 3 / sec / 1e12
    dense_bytes = 4.0 * Bd * Td * Sd * Sd
    out["dense_chain_log"] = {"pair_products_per_s": Bd * (Td - 1) / sec, "useful_tflops": tf, "B": Bd, "T": Td, "S": Sd, "ms": sec * 1e3,
                              "units_per_s": Bd * Td * Sd * Sd / sec,
                              "kernel": "chain_gemm_tc2 (linear-domain tree: TMA + tcgen05.mma cta_group::2 kind::tf32 on CTA pairs, 3 tensor products per useful flop, lo operand "
                                        "formed in shared memory) + cg_elem_max / cg_exp_orient / cg_finish_dense",
                              "roofline": {"bound": "tensor", "achieved": 3 * tf, "peak": peaks_tf32(), "unit": "TFLOP/s (tf32 issued)",
                                           "frac": 3 * tf / peaks_tf32(), "peak_source": _TF32.get("how"),
                                           "hbm": {"achieved_gbs": dense_bytes / sec / 1e9, "peak_gbs": peaks["hbm_gbs"],
                                                   "frac": dense_bytes / sec / 1e9 / peaks["hbm_gbs"],
                                                   "note": "dense input read once (algorithmic bytes); the matrix-chain product is tensor-bound at S=1024 (S/2 flop per byte)"}}}
    # chunk summary of a time-sharded chain (the cost of breaking the dependency chain): S x S matrix-chain product of
    # T steps of A + obs[t], linear-domain GEMM tree; this is what bounds the multi-GPU speed-up of BASELINE config 5
    Sc, Tc = 512, 32768
    Ac = torch.randn(Sc, Sc, device=device, generator=g).log_softmax(-1)
    oc = torch.randn(1, Tc, Sc, device=device, generator=g)
    sec_c = timeit(lambda: fb.hmm_chunk_summary(Ac, oc, return_offsets=True), n=2)
    tfc = 2.0 * Sc ** 3 * (Tc - 1) / sec_c / 1e12
    out["chunk_summary_S512"] = {"us_per_step": sec_c * 1e6 / Tc, "useful_tflops": tfc, "S": Sc, "T": Tc, "ms": sec_c * 1e3,
                                 "kernel": "chain_gemm_tc2 (CTA pairs, tcgen05.mma cta_group::2; the leaf level scales P by exp(obs) in shared memory; nothing T x S x S is materialised)",
                                 "roofline": {"bound": "tensor", "achieved": 3 * tfc, "peak": peaks_tf32(), "unit": "TFLOP/s (tf32 issued)",
                                              "frac": 3 * tfc / peaks_tf32(), "peak_source": _TF32.get("how")}}
    del Ac, oc
    trans_small = trans[:4, :64].contiguous()
    del trans
    sec = timeit(lambda: fb.sequential_sum_product("max", "add", trans_small), n=1)
    out["dense_chain_max"] = {"pair_products_per_s": 4 * 63 / sec, "fp32_tflops": 2.0 * 4 * 63 * Sd ** 3 / sec / 1e12, "B": 4, "T": 64, "S": Sd,
                              "ms": sec * 1e3, "kernel": "semiring_matmul_tiled<MAX> (CUDA cores: max-plus has no tensor-core form)"}
    del trans_small
    # Viterbi (C3 shape at reduced T): S=4096, B=1, bit-exact max-plus forward + back-pointers + backtrace
    Sv, Tv = 4096, 2048
    Av = torch.randn(Sv, Sv, device=device, generator=g)
    ov = torch.randn(1, Tv, Sv, device=device, generator=g)
    iv = torch.randn(1, Sv, device=device, generator=g)
    sec = timeit(lambda: fb.hmm_viterbi(iv, Av, ov), n=2)
    out["viterbi"] = {"units_per_s": Tv * Sv * Sv / sec, "us_per_step": sec * 1e6 / Tv, "S": Sv, "B": 1, "T": Tv, "ms": sec * 1e3,
                      "kernel": "hmm_maxplus_stream<1,true>",
                      "roofline": {"bound": "l2", "achieved": 4.0 * Sv * (Sv - 1664) * Tv / sec / 1e9, "peak": l2["gbs"], "unit": "GB/s",
                                   "frac": 4.0 * Sv * (Sv - 1664) * Tv / sec / 1e9 / l2["gbs"], "peak_source": l2["how"],
                                   "on_chip_gbs_including_shared_memory": 4.0 * Sv * Sv * Tv / sec / 1e9,
                                   "alu": {"achieved_tflops": 2.0 * Sv * Sv * Tv / sec / 1e12, "peak_tflops": fp32_alu_peak(),
                                           "frac": 2.0 * Sv * Sv * Tv / sec / 1e12 / fp32_alu_peak(),
                                           "note": "one add + one max per (prev, curr) pair against 148 SMs x 128 lanes x 1.965 GHz (non-FMA issue rate)"},
                                   "note": "A (64 MB) is re-read by every step: 1664 of the 4096 row slots of each strip stay in shared memory for the whole pass, "
                                           "the other 2432 stream from L2 (ncu: L2 hit 97 %, DRAM throughput 0.14 %), so the bytes are L2 bytes and the "
                                           "denominator is the L2 read rate measured in this run, not HBM"}}
    del Av, ov, iv
    # forward-backward marginals (C5 shape at reduced T): S=512, B=1
    Sm, Tm = 512, 8192
    Am = torch.randn(Sm, Sm, device=device, generator=g).log_softmax(-1)
    om = torch.randn(1, Tm, Sm, device=device, generator=g)
    im = torch.randn(1, Sm, device=device, generator=g).log_softmax(-1)
    sec = timeit(lambda: fb.hmm_marginals(im, Am, om, want_xi=True), n=2)
    out["marginals"] = {"units_per_s": 3.0 * Tm * Sm * Sm / sec, "us_per_step": sec * 1e6 / Tm, "S": Sm, "B": 1, "T": Tm, "ms": sec * 1e3,
                        "note": "forward + backward dataflow passes + gamma + xi_sum (outer-product GEMM over time); 3 T S^2 units"}
    sec = timeit(lambda: fb.hmm_forward(im, Am, om), n=2)
    out["forward_S512_B1"] = {"units_per_s": Tm * Sm * Sm / sec, "us_per_step": sec * 1e6 / Tm, "ms": sec * 1e3,
                              "kernel": "hmm_flow_forward<4,1,8,true> (two-hop dataflow kernel, 4-byte words; the single-hop kernel serves forward-backward)"}
    # many short chains with a small state space (the batched DiscreteHMM.log_prob of the reference's examples: hidden_dim 16)
    Bs, Ts, Ss = 16384, 1000, 16
    As = torch.randn(Ss, Ss, device=device, generator=g).log_softmax(-1)
    os_ = torch.randn(Bs, Ts, Ss, device=device, generator=g)
    is_ = torch.randn(Bs, Ss, device=device, generator=g).log_softmax(-1)
    sec = timeit(lambda: fb.hmm_forward(is_, As, os_), n=3)
    small_bytes = 4.0 * Bs * Ts * Ss
    out["forward_S16_B16384"] = {"units_per_s": Bs * Ts * Ss * Ss / sec, "chain_steps_per_s": Bs * Ts / sec, "ms": sec * 1e3, "S": Ss, "B": Bs, "T": Ts,
                                 "kernel": "hmm_small_forward_tc<2> (16 chains per warp as mma.sync rows, 3xTF32 against P in shared memory)",
                                 "roofline": {"bound": "hbm", "achieved": small_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                              "frac": small_bytes / sec / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": small_bytes,
                                              "note": "the observations are read once; a step of a 16-chain warp is a dependent chain of ~0.3 us"}}
    del As, os_, is_
    # forward-filter backward-sample on the same chain: 64 posterior paths
    Ns = 64
    us = torch.rand(1, Ns, Tm + 1, device=device, generator=g)
    sec = timeit(lambda: fb.hmm_posterior_sample(im, Am, om, uniforms=us), n=2)
    out["posterior_sampling"] = {"path_steps_per_s": Ns * Tm / sec, "samples": Ns, "S": Sm, "T": Tm, "ms": sec * 1e3,
                                 "kernel": "hmm_direct_forward<8> (alpha history) + hmm_backward_sample (warp per path)"}
    del Am, om, im, us
    # log-semiring forward beyond the dataflow kernels: S = 4096 (linear-domain streaming GEMV over strip-major P)
    Sl, Tl = 4096, 512
    Al = torch.randn(Sl, Sl, device=device, generator=g).log_softmax(-1)
    ol = torch.randn(1, Tl, Sl, device=device, generator=g)
    il = torch.randn(1, Sl, device=device, generator=g).log_softmax(-1)
    sec = timeit(lambda: fb.hmm_forward(il, Al, ol), n=2)
    out["forward_S4096_B1"] = {"units_per_s": Tl * Sl * Sl / sec, "us_per_step": sec * 1e6 / Tl, "ms": sec * 1e3, "kernel": "hmm_logsum_stream<1>",
                               "roofline": {"bound": "l2", "achieved": 4.0 * Sl * Sl * Tl / sec / 1e9, "peak": l2["gbs"], "unit": "GB/s",
                                            "frac": 4.0 * Sl * Sl * Tl / sec / 1e9 / l2["gbs"], "peak_source": l2["how"],
                                            "note": "P (64 MB) is re-read every step from L2 / shared memory, not from DRAM"}}
    return out


_TF32 = {}


def peaks_tf32():
    """TF32 tensor peak MEASURED in this run: torch.matmul of 8192^3 fp32 operands with TF32 allowed (cuBLAS), best of 5.  Falls
    back to half of the measured sustained bf16 rate (TF32 is half-rate on B200) if the probe fails."""
    if "v" in _TF32:
        return _TF32["v"]
    try:
        import torch

        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        torch.matmul(a, b)
        best = float("inf")
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        _TF32["v"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        _TF32["how"] = "measured in this run: cuBLAS TF32 matmul 8192^3, best of 5"
    except Exception:
        _TF32["v"] = measured_peaks()["tflops"] / 2.0
        _TF32["how"] = "fallback: half of the measured sustained bf16 rate"
    return _TF32["v"]


def fp32_alu_peak():
    return 148 * 128 * 1.965e9 / 1e12  # one non-FMA fp32 instruction per lane per cycle, TFLOP/s


def measure_l2_bandwidth(device):
    """Copy rate of an L2-resident working set (32 MB source + 32 MB destination inside the 126 MB L2): read + write bytes / time,
    best of 20 back-to-back pairs of copies (two per timed region so that launch latency does not dominate)."""
    import torch

    n = 8 << 20
    a = torch.randn(n, device=device)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    best = float("inf")
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        a.copy_(b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 2)
    return {"gbs": 2.0 * n * 4 / (best * 1e-3) / 1e9, "how": "measured in this run: torch copy inside a 64 MB L2-resident working set (read + write bytes), best of 20"}


if __name__ == "__main__":
    main()
