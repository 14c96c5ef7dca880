"""ig_gen_tables_ahead: the look-ahead table build of a training loop (programmatic dependent launch in front of the objective that
uses the PREVIOUS table).  Bit-identical tables, and a chained run with a different echo train per step gives the scalars and gradients
of the plain, fully serialised sequence."""
import numpy as np
import pytest
import torch

from idealgan import _lib as L
from idealgan import ops, synth

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


@pytest.mark.parametrize("nb,ne", [(1, 6), (5, 3), (64, 6), (7, 12)])
def test_same_table_bit_for_bit(nb, ne):
    lib = L.load()
    rng = np.random.default_rng(nb * 100 + ne)
    te = dev(synth.te_random(nb, ne, rng, d_te_min=0.9e-3, d_te_d=0.4e-3)[:, :, 0])
    a = torch.full((nb, L.TAB_FLOATS), -7.0, device="cuda")
    b = torch.full((nb, L.TAB_FLOATS), -7.0, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.ig_gen_tables(te.data_ptr(), nb, ne, 1.5, a.data_ptr(), st), "ig_gen_tables")
    L.check(lib.ig_gen_tables_ahead(te.data_ptr(), nb, ne, 3.0, b.data_ptr(), st), "ig_gen_tables_ahead")       # another field: not a stale copy
    assert not torch.equal(a, b)
    L.check(lib.ig_gen_tables_ahead(te.data_ptr(), nb, ne, 1.5, b.data_ptr(), st), "ig_gen_tables_ahead")
    assert torch.equal(a, b)
    c = torch.empty_like(b)
    assert ops.gen_tables_ahead(te.unsqueeze(-1), 1.5, c) is c and torch.equal(a, c)     # the Python wrapper
    with pytest.raises(ValueError):
        ops.gen_tables_ahead(te, 1.5, c[:, :-1])
    assert lib.ig_gen_tables_ahead(0, nb, ne, 1.5, b.data_ptr(), st) == -1                  # IG_E_ARG
    assert lib.ig_gen_tables_ahead(te.data_ptr(), nb, 17, 1.5, b.data_ptr(), st) == -2     # IG_E_NE


def test_chained_steps_equal_the_serialised_sequence():
    lib = L.load()
    nb, H, W, ne, steps = 8, 96, 128, 6, 12
    nv = H * W
    rng = np.random.default_rng(77)
    tes = [dev(synth.te_random(nb, ne, rng, d_te_min=1.4e-3, d_te_d=0.6e-3)[:, :, 0]) for _ in range(steps + 1)]
    maps = dev(synth.wfpm_maps(nb, H, W, rng))
    batches = []
    for i in range(steps):
        sig = ops.ideal_fwd(L.MODEL_WFPM, maps, ops.gen_tables(tes[i].unsqueeze(-1), 1.5), ne)
        batches.append((sig + 0.01 * torch.randn_like(sig)).contiguous())
    pm = (maps[:, 2:3] + 0.05 * torch.randn_like(maps[:, 2:3])).contiguous()
    inv_n = 1.0 / batches[0].numel()
    st = torch.cuda.current_stream().cuda_stream
    scratch = ops.loss_scratch(torch.device("cuda"), nb, nv)

    def objective(acqs, tab, loss, g):
        L.check(lib.ig_a2a_loss(acqs.data_ptr(), pm.data_ptr(), nv * 2, tab.data_ptr(), nb, ne, nv, 200.0, inv_n, g.data_ptr(), 0, 0,
                                loss.data_ptr(), scratch.data_ptr(), scratch.numel(), st), "ig_a2a_loss")

    # serialised: table, synchronise, objective, synchronise
    ref_loss, ref_g = [], []
    tab = torch.empty((nb, L.TAB_FLOATS), device="cuda")
    for i in range(steps):
        L.check(lib.ig_gen_tables(tes[i].data_ptr(), nb, ne, 1.5, tab.data_ptr(), st), "ig_gen_tables")
        torch.cuda.synchronize()
        loss, g = torch.zeros(1, device="cuda"), torch.empty((nb, 1, H, W, 2), device="cuda")
        objective(batches[i], tab, loss, g)
        torch.cuda.synchronize()
        ref_loss.append(loss.item())
        ref_g.append(g)
    assert len(set(ref_loss)) == steps                                          # every step has its own echo train: a stale table would show

    # chained: the table of step i + 1 goes in front of objective i, three buffers in rotation, nothing synchronises in between
    for _ in range(3):                                                          # repeated: an ordering bug need not show every time
        tabs = [torch.zeros((nb, L.TAB_FLOATS), device="cuda") for _ in range(3)]
        losses = [torch.zeros(1, device="cuda") for _ in range(steps)]
        gs = [torch.empty((nb, 1, H, W, 2), device="cuda") for _ in range(steps)]
        L.check(lib.ig_gen_tables(tes[0].data_ptr(), nb, ne, 1.5, tabs[0].data_ptr(), st), "ig_gen_tables")
        for i in range(steps):
            L.check(lib.ig_gen_tables_ahead(tes[i + 1].data_ptr(), nb, ne, 1.5, tabs[(i + 1) % 3].data_ptr(), st), "ig_gen_tables_ahead")
            objective(batches[i], tabs[i % 3], losses[i], gs[i])
        torch.cuda.synchronize()
        for i in range(steps):
            assert abs(losses[i].item() - ref_loss[i]) <= 1e-6 * abs(ref_loss[i]), (i, losses[i].item(), ref_loss[i])     # schedule-dependent partial sums
            assert torch.equal(gs[i], ref_g[i]), f"gradient of step {i}"
