"""ig_ideal_decode (config 5's consumer, gen_LDM_dataset.py:156-158,216-237), the pinned-buffer allocator and the copy probe.

The decode kernel is held to the vectors the script's own statements produce (tests/golden/ldm.npz), to the oracle on odd
shapes (scalar-lane path) and at the BASELINE slice size, for the three map parameterisations; the streamed shard is held
to the direct call."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import assert_close
from idealgan import _lib as L
from idealgan import dist as igdist
from idealgan import ops, synth, torch_ops
from oracle import ideal_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def _same_with_nan(x, ref, tol, what):
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(x), nan), f"{what}: NaN pattern differs"
    assert_close(np.nan_to_num(x), np.nan_to_num(ref), tol, what)


def test_decode_vs_the_scripts_own_statements(golden):
    g = golden("ldm")
    shat, mag, pdff, r2s = torch_ops.ldm_decode(dev(g["ldm_maps"]), dev(g["ldm_te"]))
    assert_close(host(shat), g["ldm_sig"], TOL, "signals")
    assert_close(host(mag), g["ldm_mag"], TOL, "magnitude images")
    assert np.isnan(g["ldm_pdff"]).any()                                   # 0 / 0 on background stays NaN through the clip
    _same_with_nan(host(pdff), g["ldm_pdff"], TOL, "PDFF")
    assert np.array_equal(host(r2s), g["ldm_r2s"])
    assert host(mag).max() == 1.0 and host(r2s).min() == 0.0               # both ends of the clip are exercised
    # without the complex output the field-map phasor is skipped: same images
    none, mag2, pdff2, r2s2 = torch_ops.ldm_decode(dev(g["ldm_maps"]), dev(g["ldm_te"]), want_shat=False)
    assert none is None and torch.equal(r2s2, r2s) and torch.equal(torch.nan_to_num(pdff2), torch.nan_to_num(pdff))
    assert_close(host(mag2), g["ldm_mag"], TOL, "magnitude images without the complex output")


@pytest.mark.parametrize("hw", [(384, 384), (17, 13)], ids=["384x384", "odd-17x13"])
@pytest.mark.parametrize("ch", [3, 4])
def test_decode_magpha_vs_oracle(hw, ch):
    H, W = hw
    nb, ne = 2, 6
    rng = np.random.default_rng(31 + ch)
    maps = synth.magpha_maps(nb, H, W, rng, bipolar=(ch == 4))
    maps[:, 0, :, :, :2] *= 1.5
    te = synth.te_random(nb, ne, rng, te_ini_d=0.4e-3, d_te_min=1.0e-3, d_te_d=0.3e-3)
    sig_r, mag_r, pdff_r, r2s_r = orc.ldm_images(torch.from_numpy(maps), te=torch.from_numpy(te))
    shat, mag, pdff, r2s = torch_ops.ldm_decode(dev(maps), dev(te))
    assert_close(host(shat), host(sig_r), TOL, "signals")
    assert_close(host(mag), host(mag_r), TOL, "magnitudes")
    _same_with_nan(host(pdff), host(pdff_r), TOL, "PDFF")
    assert np.array_equal(host(r2s), host(r2s_r))
    # unclipped magnitudes = |forward model|
    _, raw, _, _ = torch_ops.ldm_decode(dev(maps), dev(te), want_shat=False, clip=False)
    assert_close(host(raw), np.sqrt((host(sig_r) ** 2).sum(-1)), TOL, "unclipped magnitudes")


@pytest.mark.parametrize("model,maker", [(L.MODEL_WFPM, "wfpm"), (L.MODEL_FFPD, "ffpd"), (L.MODEL_MAGPHA, "magpha")])
def test_forward_only_mag_flag(model, maker):
    """ig_ideal_fwd with IG_F_ONLY_MAG: |forward model|, one channel (the forward models' counterpart of acq_to_acq(only_mag=True))."""
    nb, H, W, ne = 2, 24, 32, 6
    rng = np.random.default_rng(8)
    maps = {"wfpm": synth.wfpm_maps, "ffpd": synth.ffpd_maps, "magpha": synth.magpha_maps}[maker](nb, H, W, rng)
    te = dev(synth.te_orig(nb, ne))
    tab = ops.gen_tables(te, 1.5)
    full = ops.ideal_fwd(model, dev(maps), tab, ne)
    mag = ops.ideal_fwd(model, dev(maps), tab, ne, flags=L.F_ONLY_MAG)
    assert tuple(mag.shape) == (nb, ne, H, W, 1)
    assert_close(host(mag[..., 0]), np.sqrt((host(full) ** 2).sum(-1)), TOL, "|S_hat|")
    with pytest.raises(ValueError):
        ops.ideal_fwd(model, dev(maps), tab, ne, flags=L.F_ONLY_MAG | L.F_FLAT)


@pytest.mark.parametrize("model,maker", [(L.MODEL_WFPM, "wfpm"), (L.MODEL_FFPD, "ffpd")])
def test_decode_complex_row_models(model, maker):
    nb, H, W, ne = 2, 48, 64, 6
    rng = np.random.default_rng(5)
    maps = synth.wfpm_maps(nb, H, W, rng) if maker == "wfpm" else synth.ffpd_maps(nb, H, W, rng)
    te = synth.te_orig(nb, ne)
    fwd = orc.IDEAL_model if maker == "wfpm" else orc.IDEAL_mag
    sig_r = fwd(torch.from_numpy(maps), [1.5, torch.from_numpy(te)])
    shat, mag, pdff, r2s = torch_ops.ldm_decode(dev(maps), dev(te), model=model)
    assert_close(host(shat), host(sig_r), TOL, "signals")
    assert_close(host(mag), np.clip(np.sqrt((host(sig_r) ** 2).sum(-1)), 0, 1), TOL, "magnitudes")
    m = torch.from_numpy(maps)
    if maker == "wfpm":
        w, f = (m[:, 0] ** 2).sum(-1).sqrt(), (m[:, 1] ** 2).sum(-1).sqrt()
        r2_raw = m[:, 2, :, :, 1]
    else:
        w, f = ((1 - m[:, 0, :, :, 0]) * m[:, 1, :, :, 0]).abs(), (m[:, 0, :, :, 0] * m[:, 1, :, :, 0]).abs()
        r2_raw = m[:, 1, :, :, 1]
    _same_with_nan(host(pdff), torch.clamp(f / (w + f), 0, 1).numpy(), TOL, "PDFF")
    assert np.array_equal(host(r2s), torch.clamp(r2_raw, 0, 1).numpy())


def test_wfpm_accepts_more_than_four_rows_last_row_is_bipolar():
    """IDEAL_model reads out_maps[:, -1] whenever shape[1] > 3 (IDEAL_model.py:246-247): a 5-row tensor works like the
    4-row one with the same last row; the rows in between get zero gradient."""
    import wflib as wf
    nb, H, W, ne = 2, 16, 16, 6
    rng = np.random.default_rng(6)
    maps4 = synth.wfpm_maps(nb, H, W, rng, bipolar=True)
    maps5 = np.concatenate([maps4[:, :3], rng.standard_normal((nb, 1, H, W, 2)).astype(np.float32), maps4[:, 3:]], axis=1)
    te = dev(synth.te_orig(nb, ne))
    m4, m5 = dev(maps4).requires_grad_(True), dev(maps5).requires_grad_(True)
    y4, y5 = wf.IDEAL_model(m4, [1.5, te]), wf.IDEAL_model(m5, [1.5, te])
    assert torch.equal(y4, y5)
    up = torch.randn_like(y4)
    (g4,), (g5,) = torch.autograd.grad(y4, [m4], up), torch.autograd.grad(y5, [m5], up)
    assert torch.equal(g5[:, :3], g4[:, :3]) and torch.equal(g5[:, 4], g4[:, 3]) and not g5[:, 3].any()
    ref = orc.IDEAL_model(torch.from_numpy(maps5), [1.5, te.cpu()])
    assert_close(host(y5), host(ref), TOL, "5-row forward vs oracle")


def test_streamed_shard_images_equal_direct_call_and_buffers_are_pinned():
    nb, H, W, ne = 11, 32, 48, 6
    rng = np.random.default_rng(7)
    maps = synth.magpha_maps(nb, H, W, rng, bipolar=False)
    te = synth.te_orig(nb, ne)
    maps_h = igdist.pinned_empty(maps.shape)
    assert maps_h.is_pinned() and maps_h.dtype == torch.float32
    maps_h.copy_(torch.from_numpy(maps))
    sig_h, img = igdist.synthesize_to_host(L.MODEL_MAGPHA, maps_h, te, chunk_nb=4, images=True)       # ragged last chunk
    shat, mag, pdff, r2s = torch_ops.ldm_decode(dev(maps), dev(te))
    assert torch.equal(sig_h, shat.cpu()) and torch.equal(img["mag"], mag.cpu()) and torch.equal(img["r2s"], r2s.cpu())
    assert torch.equal(torch.nan_to_num(img["pdff"]), torch.nan_to_num(pdff.cpu()))
    assert all(t.is_pinned() for t in img.values()) and sig_h.is_pinned()
    none, img2 = igdist.synthesize_to_host(L.MODEL_MAGPHA, maps_h, te, chunk_nb=5, images=True, out_host=False)
    assert none is None and torch.equal(img2["mag"], img["mag"])
    view = img2["mag"][3:]
    del img2
    import gc
    gc.collect()
    assert torch.equal(view, img["mag"][3:])                                  # a view keeps the ig_host_alloc buffer alive


def test_copy_probe_and_numa_node_report():
    lib = L.load()
    node = lib.ig_host_numa_node(torch.cuda.current_device())
    assert node >= -1
    n = 8 << 20
    h = igdist.pinned_empty((n,))
    h2 = igdist.pinned_empty((n,), write_combined=True)
    d, d2 = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    sec = ctypes.c_double()
    for direction in (0, 1, 2):
        L.check(lib.ig_copy_probe(h.data_ptr(), d.data_ptr(), h2.data_ptr(), d2.data_ptr(), n * 4, 4, direction, ctypes.byref(sec)), "ig_copy_probe")
        assert 1.0 < n * 4 * 4 / sec.value / 1e9 < 200.0, (direction, sec.value)      # a PCIe-class rate, not zero and not HBM
    h.fill_(3.0)
    L.check(lib.ig_copy_probe(h.data_ptr(), d.data_ptr(), 0, 0, n * 4, 1, 0, ctypes.byref(sec)), "ig_copy_probe")
    assert float(d[-1]) == 3.0
