"""Time-sharded forward-backward of ONE long chain across the GPUs of a box (BASELINE config 5 shape).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/time_shard_bench.py --T 1048576 --S 512

Every rank generates its own contiguous chunk of observations (seeded by chunk index), reduces it to a summary (tensor-core
tree), ONE all-gather moves the summaries, then the local forward-backward runs from the folded boundaries.  Timing: CUDA
events, max over ranks.  Rank 0 then recomputes the whole chain alone and compares logZ and a slice of gamma."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import funsor_b200 as fb  # noqa: E402
from funsor_b200 import distributed as fd  # noqa: E402


def chunk_obs(r, B, Tr, S, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + r)
    return torch.randn(B, Tr, S, device=dev, generator=g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=1048576)
    ap.add_argument("--S", type=int, default=512)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, S = a.B, a.T, a.S
    assert T % world == 0
    Tr = T // world
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    A = torch.log_softmax(torch.randn(S, S, device=dev, generator=g), -1)
    init = torch.log_softmax(torch.randn(B, S, device=dev, generator=g), -1)
    obs = chunk_obs(rank, B, Tr, S, dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times, parts = [], []
    for rep in range(a.reps + 1):
        sync()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        summ = fb.hmm_chunk_summary(A, obs, return_offsets=True)
        e[1].record()
        logZ, gamma = fd.hmm_marginals_time_sharded(init, A, obs, summary_fn=lambda A_, o_: summ)
        e[2].record()
        sync()
        t = torch.tensor([e[0].elapsed_time(e[2]), e[0].elapsed_time(e[1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(t[0].item())
            parts.append(t[1].item())
    out = {
        "workload": "time-sharded forward-backward",
        "T": T,
        "S": S,
        "B": B,
        "n_gpus": world,
        "ms_total": min(times),
        "ms_summary": min(parts),
        "us_per_step": min(times) * 1e3 / T,
        "logZ": logZ[0].item(),
    }
    if rank == 0 and not a.no_check:
        full = torch.cat([chunk_obs(r, B, Tr, S, dev) for r in range(world)], 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z1, g1 = fb.hmm_marginals(init, A, full, want_xi=False)[:2]
        e1.record()
        torch.cuda.synchronize()
        out["ms_one_gpu"] = e0.elapsed_time(e1)
        out["logZ_one_gpu"] = z1[0].item()
        out["logZ_rel_err"] = abs(z1[0].item() - logZ[0].item()) / abs(z1[0].item())
        out["gamma_max_abs_err_chunk0"] = (g1[:, :Tr] - gamma).abs().max().item()
        out["gamma_rowsum_err"] = (gamma.exp().sum(-1) - 1).abs().max().item()  # gamma is log p(z_t | obs)
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
