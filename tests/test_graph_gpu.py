"""CUDA-graph capture of the small-batch step (config 4 runs at nb = 3, config 1 at nb = 1: launch-bound through Python).
The fused objectives keep their reduction state in a scratch buffer that the kernel re-arms itself, so a captured step
must replay correctly any number of times with new data written into the same buffers."""
import numpy as np
import pytest
import torch

from idealgan import _lib as L
from idealgan import ops, synth

pytestmark = pytest.mark.gpu


def _data(nb, H, W, ne, seed):
    rng = np.random.default_rng(seed)
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = synth.te_random(nb, ne, rng)
    return torch.from_numpy(maps).cuda(), torch.from_numpy(np.ascontiguousarray(te[:, :, 0])).cuda()


@pytest.mark.parametrize("objective", ["a2a", "uq", "magpha"])
def test_captured_step_replays_with_new_data(objective):
    nb, H, W, ne = 3, 64, 64, 6
    maps, te = _data(nb, H, W, ne, 0)
    acqs = torch.empty((nb, ne, H, W, 2), device="cuda")
    pm = torch.empty((nb, 1, H, W, 2), device="cuda")
    pv = torch.full((nb, 1, H, W, 1), 2e-3, device="cuda")
    rv = torch.full((nb, 1, H, W, 1), 1e-3, device="cuda")
    mp = torch.empty((nb, 2, H, W, 4), device="cuda")
    g = torch.Generator(device="cuda")

    def fill(seed):                                        # new data, same buffers
        m, t = _data(nb, H, W, ne, seed)
        te.copy_(t)
        tab = ops.gen_tables(te, 1.5)
        g.manual_seed(seed)
        s = ops.ideal_fwd(L.MODEL_WFPM, m, tab, ne)
        acqs.copy_(s + 0.02 * torch.randn(s.shape, device="cuda", generator=g) * (s != 0))
        pm.copy_(m[:, 2:3] * 0.9)
        mp.copy_(torch.from_numpy(synth.magpha_maps(nb, H, W, np.random.default_rng(seed))).cuda())

    def step():
        tab = ops.gen_tables(te, 1.5)
        if objective == "a2a":
            loss, grad, _, _ = ops.a2a_loss(acqs, pm, tab)
        elif objective == "uq":
            out = ops.a2a_uq_loss(acqs, pm, pv, pm[..., 1:2].contiguous(), rv, tab)
            loss, grad = out[0], out[1]
        else:
            loss, grad, _ = ops.ideal_loss(L.MODEL_MAGPHA, mp, acqs, tab)
        return loss, grad

    fill(1)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):                                 # warm-up off the capture (scratch allocation, lazy module load)
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_g, grad_g = step()
    for seed in (2, 3, 4):
        fill(seed)
        graph.replay()
        torch.cuda.synchronize()
        got_l, got_g = loss_g.clone(), grad_g.clone()
        ref_l, ref_g = step()
        np.testing.assert_allclose(got_l.item(), ref_l.item(), rtol=1e-6)
        assert torch.equal(got_g, ref_g)


def test_drop_in_call_captured_in_a_graph_rebuilds_its_table_on_replay():
    """`wf.*` calls cache the per-sample table per echo train -- but not inside a capture: the table belongs to the captured step and
    must follow the echo-time buffer on every replay (a cached one would be baked into the graph; one built inside the capture lives
    in the graph's private pool and may not be handed to later eager calls)."""
    import wflib as wf
    from idealgan import torch_ops as TO
    nb, H, W, ne = 2, 32, 48, 6
    maps, te2d = _data(nb, H, W, ne, 10)
    te = te2d[:, :, None].contiguous()
    tab = ops.gen_tables(te, 1.5)
    acqs = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne)
    pm = maps[:, 2:3].contiguous()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            wf.get_rho(acqs, pm, te=te)                        # eager warm-up: this DOES cache a table for (te, version)
    torch.cuda.current_stream().wait_stream(side)
    cached = len(TO.table_cache.entries)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        rho_g = wf.get_rho(acqs, pm, te=te)
    assert len(TO.table_cache.entries) == cached               # nothing stored from inside the capture
    for seed in (11, 12):
        m2, t2 = _data(nb, H, W, ne, seed)
        te.copy_(t2[:, :, None])                               # new echo trains in the same buffer
        acqs.copy_(ops.ideal_fwd(L.MODEL_WFPM, m2, ops.gen_tables(te, 1.5), ne))
        pm.copy_(m2[:, 2:3])
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(rho_g, wf.get_rho(acqs, pm, te=te)), "the captured step used a stale table"
