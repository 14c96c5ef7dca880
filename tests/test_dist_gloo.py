"""World-size-2 (and 3, ragged) runs of the sharding layer on CPU with the gloo backend.  The per-rank operator is
injected, so here it is the CPU oracle; on the GPU box bench.py runs the same layer with the CUDA kernels over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idealgan import dist as igdist
from idealgan import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [igdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        igdist.shard_bounds(4, 2, 2)


def _oracle_loss(acqs, pm, te, inv_n, field=1.5):
    sys.path.insert(0, ROOT)
    from oracle import ideal_oracle as orc
    rho, recon = orc.acq_to_acq(acqs, pm, te=te, field=field)
    recon = torch.where(acqs != 0, recon, torch.zeros_like(recon))
    return ((acqs - recon) ** 2).sum() * inv_n


def _worker(rank, world, port, nb, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from oracle import ideal_oracle as orc
        rng = np.random.default_rng(0)                                     # identical data on every rank, then sharded
        H = W = 8
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = torch.from_numpy(synth.te_random(nb, 6, rng))
        acqs = torch.from_numpy(synth.add_noise(orc.IDEAL_model(torch.from_numpy(maps), [1.5, te]).numpy(), rng))
        pm = torch.from_numpy(np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.9))
        p_full = pm.clone().requires_grad_(True)
        full = _oracle_loss(acqs, p_full, te, 1.0 / acqs.numel())
        (g_full,) = torch.autograd.grad(full, [p_full])
        p_loc = igdist.shard(pm).clone().requires_grad_(True)
        total, local = igdist.sharded_physics_loss(_oracle_loss, igdist.shard(acqs), p_loc, igdist.shard(te), acqs.numel())
        (g_loc,) = torch.autograd.grad(local, [p_loc])
        gathered = igdist.gather_batch(g_loc, nb)
        ok = (abs(total.item() - full.item()) <= 1e-6 * full.item()
              and torch.allclose(gathered, g_full, rtol=1e-5, atol=1e-9)
              and tuple(igdist.shard(acqs).shape[1:]) == tuple(acqs.shape[1:]))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nb", [(2, 4), (3, 5)])
def test_sharded_objective_equals_global_objective(world, nb):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, nb, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}


def _reducer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        red = igdist.AsyncLossReducer(torch.device("cpu"), depth=2)
        got = []
        for step in range(5):
            buf = red.acquire()
            if step >= 2:                       # the slot handed out again holds the reduced scalar of step - 2
                got.append((step - 2, float(buf.item())))
            buf.fill_(float((rank + 1) * (step + 1)))          # this rank's local objective term of the step
            red.submit()
        last = float(red.last().item())
        want = lambda st: float(sum((r + 1) * (st + 1) for r in range(world)))
        out[rank] = all(abs(v - want(st)) < 1e-6 for st, v in got) and abs(last - want(4)) < 1e-6 and len(got) == 3
    finally:
        dist.destroy_process_group()


def test_async_loss_reducer_overlaps_and_orders():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_reducer_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


def test_async_loss_reducer_single_process_is_a_no_op():
    red = igdist.AsyncLossReducer(torch.device("cpu"), depth=3)
    for step in range(4):
        red.acquire().fill_(step + 1.0)
        red.submit()
    assert red.last().item() == 4.0


def _peer_setup_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            igdist.PeerLossExchange(torch.device("cuda", 0))      # no CUDA device here: mailbox creation fails
            out[rank] = "created"
        except RuntimeError as e:
            out[rank] = str(e)
        dist.barrier()                                             # nobody was left behind in a collective
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.is_available(), reason="failure path: needs a box without a GPU")
def test_peer_exchange_setup_fails_on_every_rank_together():
    """PeerLossExchange's set-up is collective: when the mailbox cannot be created (no GPU here) every rank raises, none hangs."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_peer_setup_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world
        for r in range(world):
            assert "set-up failed" in out[r], out[r]


def _row_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from oracle import ideal_oracle as orc
        rng = np.random.default_rng(1)
        nb, H, W = 1, 9, 6                                                 # one slice, fewer slices than ranks: split the rows (ragged)
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = torch.from_numpy(synth.te_random(nb, 6, rng))
        acqs = torch.from_numpy(synth.add_noise(orc.IDEAL_model(torch.from_numpy(maps), [1.5, te]).numpy(), rng))
        pm = torch.from_numpy(np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.9))
        p_full = pm.clone().requires_grad_(True)
        full = _oracle_loss(acqs, p_full, te, 1.0 / acqs.numel())
        (g_full,) = torch.autograd.grad(full, [p_full])
        p_loc = igdist.shard(pm, axis=2).clone().requires_grad_(True)
        total, local = igdist.sharded_physics_loss(_oracle_loss, igdist.shard(acqs, axis=2).contiguous(), p_loc, te, acqs.numel())
        (g_loc,) = torch.autograd.grad(local, [p_loc])
        gathered = igdist.gather_batch(g_loc, H, axis=2)
        out[rank] = bool(abs(total.item() - full.item()) <= 1e-6 * full.item() and torch.allclose(gathered, g_full, rtol=1e-5, atol=1e-9)
                         and gathered.shape == g_full.shape)
    finally:
        dist.destroy_process_group()


def test_row_sharding_for_batches_smaller_than_the_world():
    """for nb < #GPUs (config 1, config 4 at nb = 3) the rows H are split instead; pointwise operators need no halo."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_row_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}
