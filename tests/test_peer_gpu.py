"""The scalar exchange fused into the objective's kernel (ig_a2a_loss_peer, idealgan/dist.py:PeerLossExchange): single
rank, and two ranks of one process on one device through ig_peer_connect_local (the multi-process CUDA-IPC path is
exercised by `bench.py --gpus N`, whose JSON line carries the exchanged loss next to the host-pipeline loss)."""
import ctypes

import numpy as np
import pytest
import torch

from idealgan import _lib as L
from idealgan import dist as igdist
from idealgan import ops, synth

pytestmark = pytest.mark.gpu


def _case(nb=4, ne=6, H=16, W=24, seed=3):
    rng = np.random.default_rng(seed)
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = torch.from_numpy(synth.te_random(nb, ne, rng)).cuda()
    tab = ops.gen_tables(te, 1.5)
    sig = ops.ideal_fwd(L.MODEL_WFPM, torch.from_numpy(maps).cuda(), tab, ne)
    noise = torch.from_numpy(rng.standard_normal(tuple(sig.shape)).astype(np.float32)).cuda()
    acqs = (sig + 0.02 * noise * (sig != 0)).contiguous()
    pm = torch.from_numpy(np.ascontiguousarray(maps[:, 2:3]) * 0.9).cuda()
    return acqs, pm, tab


def test_single_rank_exchange_equals_plain_objective():
    acqs, pm, tab = _case()
    ref_loss, ref_g, _, _ = ops.a2a_loss(acqs, pm, tab)
    ex = igdist.PeerLossExchange(torch.device("cuda", 0))
    losses = []
    for i in range(12):                                  # more steps than mailbox slots
        scale = 1.0 + 0.1 * i
        loss, g = ex.a2a_loss(acqs, pm * scale, tab)
        losses.append(loss.item())
        if i == 0:
            assert torch.equal(g, ref_g)
            np.testing.assert_allclose(loss.item(), ref_loss.item(), rtol=1e-6)      # block partials follow the dynamic tile schedule
        else:
            assert ex.prev.item() == losses[i - 1]       # the previous step's scalar, bit for bit
    assert ex.last().item() == losses[-1]
    ex.close()


@pytest.mark.parametrize("lag", [1, 2])
def test_two_ranks_in_one_process(lag):
    acqs, pm, tab = _case(nb=6)
    n_glob = acqs.numel()
    full, _, _, _ = ops.a2a_loss(acqs, pm, tab)
    lib = L.load()
    handles = (ctypes.c_void_p * 2)()
    for r in range(2):
        h = ctypes.c_void_p()
        L.check(lib.ig_peer_create(r, 2, ctypes.byref(h)), "ig_peer_create")
        handles[r] = h
    L.check(lib.ig_peer_connect_local(handles, 2), "ig_peer_connect_local")
    st = torch.cuda.current_stream().cuda_stream
    shards = [(acqs[:2].contiguous(), pm[:2].contiguous(), tab[:2].contiguous()), (acqs[2:].contiguous(), pm[2:].contiguous(), tab[2:].contiguous())]
    prev = [torch.full((1,), -1.0, device="cuda") for _ in range(2)]
    local = [torch.zeros(1, device="cuda") for _ in range(2)]
    for step in range(11):                               # more steps than mailbox slots
        for r, (a, p, t) in enumerate(shards):           # step-major: a rank never waits for a scalar that is not yet launched
            nb, ne, H, W, _ = a.shape
            g = torch.empty((nb, 1, H, W, 2), device="cuda")
            scr = ops.loss_scratch(a.device, nb, H * W)
            L.check(lib.ig_a2a_loss_peer(a.data_ptr(), p.data_ptr(), H * W * 2, t.data_ptr(), nb, ne, H * W, 200.0, 1.0 / n_glob, g.data_ptr(), 0, 0,
                                         local[r].data_ptr(), scr.data_ptr(), scr.numel(), handles[r], step, lag, prev[r].data_ptr(), st), "ig_a2a_loss_peer")
        if step >= lag:
            assert prev[0].item() == prev[1].item()                                     # same bits on both ranks
            np.testing.assert_allclose(prev[0].item(), full.item(), rtol=2e-6)          # = the objective of the whole batch
    out = torch.zeros(1, device="cuda")
    L.check(lib.ig_peer_reduce(handles[1], 10, out.data_ptr(), st), "ig_peer_reduce")
    np.testing.assert_allclose(out.item(), local[0].item() + local[1].item(), rtol=1e-7)
    # a step nobody has published: NaN after the time-out instead of a hang
    L.check(lib.ig_peer_reduce(handles[0], 1000, out.data_ptr(), st), "ig_peer_reduce")
    assert np.isnan(out.item())
    for r in range(2):
        lib.ig_peer_destroy(handles[r])


def test_peer_argument_errors():
    lib = L.load()
    h = ctypes.c_void_p()
    assert lib.ig_peer_create(2, 2, ctypes.byref(h)) == -1
    L.check(lib.ig_peer_create(0, 2, ctypes.byref(h)), "ig_peer_create")
    out = torch.zeros(1, device="cuda")
    assert lib.ig_peer_reduce(h, 0, out.data_ptr(), 0) == -1                            # not connected
    lib.ig_peer_destroy(h)


def test_publish_for_the_other_objectives():
    """ig_peer_publish: the exchange for a scalar that is already on the device (here the uncertainty-aware objective's)."""
    acqs, pm, tab = _case(nb=4)
    nb, ne, H, W, _ = acqs.shape
    rng = np.random.default_rng(4)
    pv = torch.from_numpy(rng.uniform(1e-5, 4e-3, (nb, 1, H, W, 1)).astype(np.float32)).cuda()
    rv = torch.from_numpy(rng.uniform(1e-5, 3e-3, (nb, 1, H, W, 1)).astype(np.float32)).cuda()
    rm = pm[..., 1:2].contiguous()
    lib = L.load()
    handles = (ctypes.c_void_p * 2)()
    for r in range(2):
        h = ctypes.c_void_p()
        L.check(lib.ig_peer_create(r, 2, ctypes.byref(h)), "ig_peer_create")
        handles[r] = h
    L.check(lib.ig_peer_connect_local(handles, 2), "ig_peer_connect_local")
    st = torch.cuda.current_stream().cuda_stream
    inv_n = 1.0 / acqs.numel()
    whole = ops.a2a_uq_loss(acqs, pm, pv, rm, rv, tab, inv_n=inv_n)[0]
    prev = [torch.zeros(1, device="cuda") for _ in range(2)]
    parts = []
    for step in range(3):
        parts = []
        for r, sl in enumerate((slice(0, 1), slice(1, 4))):           # ragged shards
            loc = ops.a2a_uq_loss(acqs[sl].contiguous(), pm[sl].contiguous(), pv[sl].contiguous(), rm[sl].contiguous(), rv[sl].contiguous(),
                                  tab[sl].contiguous(), inv_n=inv_n)[0]
            parts.append(loc)
            L.check(lib.ig_peer_publish(handles[r], step, 1, loc.data_ptr(), prev[r].data_ptr(), st), "ig_peer_publish")
        if step:
            assert prev[0].item() == prev[1].item()
            np.testing.assert_allclose(prev[0].item(), whole.item(), rtol=3e-6)
    assert lib.ig_peer_publish(handles[0], 3, 0, parts[0].data_ptr(), prev[0].data_ptr(), st) == -1     # lag outside [1, 3]
    for r in range(2):
        lib.ig_peer_destroy(handles[r])
