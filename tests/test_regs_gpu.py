"""Script-level reductions on the GPU (ig_mag_regs, ig_roi_maps) against the vectors made by the reference scripts' own
statements (tests/golden/regs.npz) and against the oracle at other shapes (vector and scalar kernels, ragged blocks)."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from idealgan import ops, torch_ops
from oracle import ideal_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5
NAMES = ("Ad_TV", "LS_NZ", "WF_NZ", "LS_cond", "R2_TV")


def dev(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("name", ["reg6", "reg3_odd"])          # W = 12: float4 kernel; W = 9: scalar kernel
def test_mag_regularisers_golden(golden, name):
    g = golden("regs")
    ls, demod, r2 = (dev(g[f"{name}_{k}"], grad=True) for k in ("ls", "demod", "r2"))
    w = [float(x) for x in g[f"{name}_weights"]]
    total, logs = torch_ops.mag_regularisers(ls, demod, r2, *w)
    sums = np.array([logs[k].item() for k in NAMES])
    np.testing.assert_allclose(sums, g[f"{name}_sums"], rtol=TOL)
    assert sums[2] == 0.0
    np.testing.assert_allclose(total.item(), g[f"{name}_total"], rtol=TOL)
    (3.0 * total).backward()
    for t, key in ((ls, "g_ls"), (demod, "g_demod"), (r2, "g_r2")):
        assert_close(t.grad.cpu().numpy() / 3.0, g[f"{name}_{key}"], TOL, key)


@pytest.mark.parametrize("nb,ne,H,W", [(1, 1, 1, 1), (2, 4, 5, 8), (3, 6, 33, 28), (2, 5, 17, 31), (1, 16, 64, 48)])
def test_mag_regularisers_vs_oracle(nb, ne, H, W):
    rng = np.random.default_rng(11)
    ls = rng.standard_normal((nb, 3, H, W, 1)).astype(np.float32)
    demod = rng.uniform(0, 1, (nb, ne, H, W, 1)).astype(np.float32)
    demod[:, :, : H // 3] = 0.0                                   # flat background: |0|' = 0 on both sides
    r2 = rng.uniform(0, 1, (nb, 1, H, W, 1)).astype(np.float32)
    w = (0.25, 1.5, 0.75, 2.0)
    sums, g_ls, g_demod, g_r2 = ops.mag_regs(dev(ls), dev(demod), dev(r2), w)
    t = [torch.from_numpy(a).double().requires_grad_(True) for a in (ls, demod, r2)]
    ref = orc.mag_regularisers(*t)
    np.testing.assert_allclose(sums.cpu().numpy(), [ref[k].item() for k in NAMES], rtol=TOL, atol=1e-30)
    total = ref["Ad_TV"] * w[0] + ref["LS_NZ"] * w[1] + ref["LS_cond"] * w[2] + ref["R2_TV"] * w[3]
    grads = torch.autograd.grad(total, t)
    for got, want, what in zip((g_ls, g_demod, g_r2), grads, ("g_ls", "g_demod", "g_r2")):
        assert_close(got.cpu().numpy(), want.numpy(), TOL, what)
    # bit-reproducible: block partials are added in block order
    again = ops.mag_regs(dev(ls), dev(demod), dev(r2), w)[0]
    assert torch.equal(sums, again)


def test_mag_regularisers_partial_inputs_and_errors():
    rng = np.random.default_rng(12)
    demod = rng.uniform(0, 1, (2, 3, 8, 8, 1)).astype(np.float32)
    sums, g_ls, g_demod, g_r2 = ops.mag_regs(None, dev(demod), None, (1.0, 0.0, 0.0, 0.0))
    assert g_ls is None and g_r2 is None
    ref = orc.mag_regularisers(None, torch.from_numpy(demod).double(), None)
    np.testing.assert_allclose(sums.cpu().numpy(), [ref[k].item() for k in NAMES], rtol=TOL)
    with pytest.raises(ValueError):
        ops.mag_regs(dev(np.zeros((2, 2, 8, 8, 1), np.float32)), dev(demod), None)
    with pytest.raises(ValueError):
        ops.mag_regs()
    with pytest.raises(ValueError):
        ops.mag_regs(None, torch.zeros(2, 3, 8, 8, 1), None)                 # host tensor: no CPU path


def test_roi_maps_golden(golden):
    g = golden("regs")
    maps, var = dev(g["roi_maps"]), dev(g["roi_var"])
    assert_close(torch_ops.roi_maps(maps).cpu().numpy(), g["roi_out4"], TOL)
    out5 = torch_ops.roi_maps(maps, var, "PDFF-var").cpu().numpy()
    assert_close(out5[..., :4], g["roi_out4"], TOL)
    # the variance is a three-term cancellation: compare with the fp64 restatement of the same formula, and with the
    # reference's fp32 result at the looser tolerance its own rounding allows
    ref64 = orc.roi_maps(torch.from_numpy(g["roi_maps"]).double(), torch.from_numpy(g["roi_var"]).double(), "PDFF-var").numpy()
    assert_close(out5[..., 4], ref64[..., 4], TOL, "PDFF variance vs fp64")
    assert_close(out5[..., 4], g["roi_out5"][..., 4], 3e-5, "PDFF variance vs reference fp32")
    mag = torch_ops.roi_maps(maps, var, "PDFF-var-Mag").cpu().numpy()
    assert_close(mag[..., 4], g["roi_var"][:, 1, :, :, 0], 2e-7)                     # hypotf(x, 0): within an ulp of |x|


@pytest.mark.parametrize("nb,H,W", [(1, 1, 1), (3, 17, 15), (2, 64, 48)])
def test_roi_maps_vs_oracle(nb, H, W):
    rng = np.random.default_rng(13)
    maps = rng.uniform(-0.5, 0.5, (nb, 3, H, W, 2)).astype(np.float32)
    var = rng.uniform(1e-6, 1e-3, (nb, 5, H, W, 2)).astype(np.float32)
    var[:, :4, :, :, 1] = 0
    out = torch_ops.roi_maps(dev(maps), dev(var), "PDFF-var").cpu().numpy()
    ref = orc.roi_maps(torch.from_numpy(maps).double(), torch.from_numpy(var).double(), "PDFF-var").numpy()
    assert_close(out[..., :4], ref[..., :4], TOL)
    assert_close(out[..., 4], ref[..., 4], TOL, "PDFF variance")
    # background voxel: the reference's 0/0 is NaN; same here (IEEE divisions)
    maps[0, :2, 0, 0] = 0
    out = torch_ops.roi_maps(dev(maps), dev(var), "PDFF-var").cpu().numpy()
    assert np.isnan(out[0, 0, 0, 4]) and out[0, 0, 0, 0] == 0


def test_full_size_properties():
    """BASELINE size (64 x 384 x 384 x 6): TV of a constant image is 0 with zero gradient; scaling the input scales TV;
    LS penalties vanish for a positive-definite fit (a, c > 0, b^2 < 4ac)."""
    nb, ne, H, W = 64, 6, 384, 384
    const = torch.full((nb, ne, H, W, 1), 0.7, device="cuda")
    ls = torch.empty((nb, 3, H, W, 1), device="cuda")
    ls[:, 0], ls[:, 1], ls[:, 2] = 1.0, 0.5, 2.0
    sums, g_ls, g_demod, _ = ops.mag_regs(ls, const, None, (1.0, 1.0, 1.0, 1.0))
    assert sums.abs().max().item() == 0.0 and g_ls.abs().max().item() == 0.0 and g_demod.abs().max().item() == 0.0
    x = torch.rand((nb, ne, H, W, 1), device="cuda")
    s1 = ops.mag_regs(None, x, None, want_grads=False)[0][0].item()
    s2 = ops.mag_regs(None, 2.0 * x, None, want_grads=False)[0][0].item()
    np.testing.assert_allclose(s2, 2.0 * s1, rtol=1e-6)
    ref = ((x[:, :, 1:] - x[:, :, :-1]).abs().double().sum() + (x[:, :, :, 1:] - x[:, :, :, :-1]).abs().double().sum()).item()
    np.testing.assert_allclose(s1, ref, rtol=1e-6)
