"""CPU, world_size 2 over gloo: the host-side logic of batch- and time-sharding (funsor_b200/distributed.py)
with the numpy oracle standing in for the kernels.  The N>1 GPU path uses the same functions with the CUDA
entry points and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import semiring_ref as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_fns(sid):
    def summary(A, obs):
        """Stands in for ops.hmm_chunk_summary(return_offsets=True): rows relative to a float64 row offset."""
        trans = A.double().numpy()[None, None] + obs.double().numpy()[:, :, None, :]
        m = torch.from_numpy(R.chain_product(sid, trans))
        off = m.amax(-1)
        return (m - off[..., None]).float(), off

    def forward(init, A, obs, sum_op="logaddexp"):
        return torch.from_numpy(R.hmm_forward(sid, init.double().numpy(), A.double().numpy(), obs.double().numpy()))

    def marginals(init, A, obs):
        z, gamma, _ = R.hmm_marginals(init.double().numpy(), A.double().numpy(), obs.double().numpy())
        return torch.from_numpy(z), torch.from_numpy(gamma)

    return summary, forward, marginals


def _worker(rank, world, port, sname, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from funsor_b200 import distributed as D

        sid = R.LOG if sname == "logaddexp" else R.MAX
        summary, forward, marginals = _oracle_fns(sid)
        g = torch.Generator().manual_seed(3)
        B, T, S = 3, 23, 5  # ragged: 23 steps over 2 ranks, 3 chains over 2 ranks
        init = torch.randn(B, S, generator=g).double()
        A = torch.randn(S, S, generator=g).double()
        obs = torch.randn(B, T, S, generator=g).double() - 300.0  # large log-magnitudes: exercises offsets / re-centring
        lo, hi = D.shard_bounds(T, rank, world)
        z_time, boundary = D.hmm_forward_time_sharded(init, A, obs[:, lo:hi], sum_op=sname, summary_fn=summary)
        z_batch = D.hmm_forward_batch_sharded(init, A, obs, sum_op=sname, forward=forward)
        want = torch.from_numpy(R.hmm_forward(sid, init.numpy(), A.numpy(), obs.numpy()))
        # the boundary message entering this rank's chunk equals alpha at the end of the previous chunk
        if lo > 0:
            _, alphas = R.hmm_forward(sid, init.numpy(), A.numpy(), obs.numpy(), want_alpha=True)
            wb = torch.from_numpy(alphas[:, lo - 1])
        else:
            wb = init
        errs = [float((z_time - want).abs().max() / want.abs().max()), float((z_batch - want).abs().max()),
                float((boundary - wb).abs().max() / wb.abs().max())]
        if sname == "logaddexp":  # time-sharded forward-backward: global logZ and this chunk's slice of gamma
            zg, gamma = D.hmm_marginals_time_sharded(init, A, obs[:, lo:hi], summary_fn=summary, marginals_fn=marginals)
            wz, wgamma, _ = R.hmm_marginals(init.numpy(), A.numpy(), obs.numpy())
            errs.append(float((zg - torch.from_numpy(wz)).abs().max() / abs(wz).max()))
            errs.append(float((gamma - torch.from_numpy(wgamma[:, lo:hi])).abs().max()))
        out[rank] = tuple(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sname", ["logaddexp", "max"])
def test_time_and_batch_sharding_world2(sname):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), sname, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert max(out[r]) < 1e-6, (r, out[r])  # the summaries travel as fp32 rows + float64 offsets


def test_shard_bounds_cover_and_balance():
    from funsor_b200.distributed import shard_bounds

    for n in (0, 1, 7, 16, 65536, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
