"""BASELINE.json's configurations at their own image sizes, held to the ORACLE (not to the repo's other kernels).

tests/test_full_size_gpu.py checks the full batches through size-independent properties; this file pins the kernels
bench.py / tools/kernel_bench.py launch -- the tensor-map ring of the fused objectives, ideal_kernel<6, pk>, the
get_rho / acq_to_acq forward and adjoint kernels, the persistent fused forward objective -- to oracle/ideal_oracle.py
on a few slices of 384 x 384 x 6 (C1, C2, C4, C5) and 192 x 192 x 6 with per-sample random echo times (C3), masked and
unmasked, plus slices with ragged components, at the north-star tolerance of 1e-5 (relative to the tensor's maximum).
A tile-index or stride error that only shows beyond 2^16 voxels per plane cannot pass here.

Reference call sites: train-IDEAL-unsup.py:214-218,236,255 (C2), train-IDEAL-TEaug.py:217,304 (C3),
train-IDEAL-single.py:154-157,175 (C4), gen_LDM_dataset.py:156-158 (C5), IDEAL_model.py:220-311,527-624 (C1).
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from idealgan import _lib as L
from idealgan import ops, synth
from oracle import ideal_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5          # BASELINE.json north_star: relative error 1e-5 (fp32) on signals, maps and gradients


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()


def cpu(x, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return t.requires_grad_(True) if grad else t


def host(t):
    return t.detach().cpu().numpy()


def _c2_inputs(nb, H, W, ne, rng, masked, te):
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0, masked=masked)
    with torch.no_grad():
        sig = host(orc.IDEAL_model(cpu(maps), [1.5, cpu(te)]))
    acqs = synth.add_noise(sig, rng)
    pm = np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.95)
    return maps, acqs, pm


# ---------------------------------------------------------------------------------------------------------------
# C2: the headline kernel (fused acq_to_acq -> mask -> MSE -> d/dPM on the tensor-map ring)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("masked", [True, False], ids=["disc-masked", "unmasked"])
def test_c2_fused_objective_384_vs_oracle(masked):
    nb, H, W, ne = 3, 384, 384, 6
    rng = np.random.default_rng(2024 + masked)
    te = synth.te_orig(nb, ne)
    _, acqs, pm = _c2_inputs(nb, H, W, ne, rng, masked, te)
    # slice 2 carries ragged voxels: single zeroed components inside the object, a zeroed echo along a row segment,
    # and one far beyond voxel 2^16 of the plane (the per-component mask path of train-IDEAL-unsup.py:218)
    acqs[2, 0, 200, 100, 1] = 0.0
    acqs[2, ne - 1, 190, 50:130, :] = 0.0
    acqs[2, 3, 300, 191, 0] = 0.0
    for rdtype, tol_l in ((torch.float32, TOL), (torch.float64, 2e-6)):
        p = cpu(pm).to(rdtype).requires_grad_(True)
        lref, _, _ = orc.physics_loss_a2a(cpu(acqs), p, te=cpu(te), rdtype=rdtype)
        (gref,) = torch.autograd.grad(lref, [p])
        tab = ops.gen_tables(dev(te), 1.5)
        loss, g, _, _ = ops.a2a_loss(dev(acqs), dev(pm), tab)
        assert abs(loss.item() - lref.item()) <= tol_l * lref.item(), (loss.item(), lref.item())
        assert_close(host(g), host(gref.float()), TOL, f"d loss / d PM ({rdtype})")
    if masked:
        bg = (acqs[:, 0, :, :, 0] == 0) & (acqs[:, 0, :, :, 1] == 0) & (acqs[:, 1, :, :, 0] == 0)
        assert np.abs(host(g)[:, 0][bg]).max() == 0.0


@pytest.mark.parametrize("masked", [True, False], ids=["disc-masked", "unmasked"])
def test_c2_unmodified_script_path_384_vs_oracle(masked):
    """wf.acq_to_acq -> where -> MSE -> autodiff as train-IDEAL-unsup.py:216-218,236,255 writes it: ig_a2a_fwd + ig_a2a_bwd."""
    nb, H, W, ne = 2, 384, 384, 6
    rng = np.random.default_rng(77 + masked)
    te = synth.te_orig(nb, ne)
    _, acqs, pm = _c2_inputs(nb, H, W, ne, rng, masked, te)
    a, p = cpu(acqs, True), cpu(pm, True)
    rho_r, s_r = orc.acq_to_acq(a, p, te=cpu(te))
    up_rho = rng.standard_normal(rho_r.shape).astype(np.float32)
    up_s = rng.standard_normal(s_r.shape).astype(np.float32)
    ga_r, gp_r = torch.autograd.grad((rho_r * cpu(up_rho)).sum() + (s_r * cpu(up_s)).sum(), [a, p])
    tab = ops.gen_tables(dev(te), 1.5)
    rho, shat = ops.a2a_fwd(dev(acqs), dev(pm), tab)
    assert_close(host(rho), host(rho_r), TOL, "rho_hat")
    assert_close(host(shat), host(s_r), TOL, "S_hat")
    ga, gp = ops.a2a_bwd(dev(acqs), dev(pm), tab, dev(up_rho), dev(up_s))
    assert_close(host(ga), host(ga_r), TOL, "grad acqs")
    assert_close(host(gp), host(gp_r), TOL, "grad pm")
    # the fused kernel with materialised outputs (the OUT instantiation of the ring) gives the same tensors
    _, _, rho2, shat2 = ops.a2a_loss(dev(acqs), dev(pm), tab, want_rho=True, want_shat=True)
    assert_close(host(rho2), host(rho_r), TOL, "rho_hat (ring, OUT)")
    assert_close(host(shat2), host(s_r), TOL, "S_hat (ring, OUT)")


# ---------------------------------------------------------------------------------------------------------------
# C1: IDEAL_Layer forward + LS solve on one 384 x 384 x 6 slice (and the solve's adjoint)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bipolar", [False, True], ids=["3-row", "bipolar-row"])
def test_c1_forward_and_solve_384_vs_oracle(bipolar):
    nb, H, W, ne = 2, 384, 384, 6
    rng = np.random.default_rng(11 + bipolar)
    maps = synth.wfpm_maps(nb, H, W, rng, bipolar=bipolar, neg_r2_frac=0.1)
    te = synth.te_orig(nb, ne)
    m = cpu(maps, True)
    s_r = orc.IDEAL_model(m, [1.5, cpu(te)])
    up = rng.standard_normal(s_r.shape).astype(np.float32)
    (gm_r,) = torch.autograd.grad((s_r * cpu(up)).sum(), [m])
    tab = ops.gen_tables(dev(te), 1.5)
    S = ops.ideal_fwd(L.MODEL_WFPM, dev(maps), tab, ne)
    assert_close(host(S), host(s_r), TOL, "IDEAL_Layer forward")
    gm = ops.ideal_bwd(L.MODEL_WFPM, dev(maps), tab, ne, dev(up))
    assert_close(host(gm), host(gm_r), TOL, "IDEAL_Layer adjoint")
    # LS solve of noisy echoes with an imperfect (phi, R2*) estimate, and its adjoint
    acqs = synth.add_noise(host(s_r), rng)
    pmaps = maps.copy()
    pmaps[:, 2] *= np.float32(0.97)
    pm_full = np.ascontiguousarray(pmaps[:, 2:])            # (phi, R2*) row [+ bipolar row]
    a, p = cpu(acqs, True), cpu(pm_full, True)
    rho_r, dem_r = orc.get_rho(a, p, te=cpu(te), acq_demod=True)
    up_rho = rng.standard_normal(rho_r.shape).astype(np.float32)
    up_dem = rng.standard_normal(dem_r.shape).astype(np.float32)
    ga_r, gp_r = torch.autograd.grad((rho_r * cpu(up_rho)).sum() + (dem_r * cpu(up_dem)).sum(), [a, p])
    rho, dem = ops.get_rho_fwd(dev(acqs), dev(pm_full), tab, want_demod=True)
    assert_close(host(rho), host(rho_r), TOL, "get_rho")
    assert_close(host(dem), host(dem_r), TOL, "get_rho demodulated echoes")
    ga, gp = ops.get_rho_bwd(dev(acqs), dev(pm_full), tab, dev(up_rho), dev(up_dem))
    assert_close(host(ga), host(ga_r), TOL, "get_rho grad acqs")
    assert_close(host(gp), host(gp_r), TOL, "get_rho grad pm")


# ---------------------------------------------------------------------------------------------------------------
# C3: TE-augmented training shapes, 192 x 192 x 6 with a different random echo train per sample
# ---------------------------------------------------------------------------------------------------------------
def test_c3_per_sample_echo_times_192_vs_oracle():
    nb, H, W, ne = 4, 192, 192, 6
    rng = np.random.default_rng(303)
    te = synth.te_random(nb, ne, rng)
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.1)
    m = cpu(maps, True)
    s_r = orc.IDEAL_model(m, [1.5, cpu(te)])
    up = rng.standard_normal(s_r.shape).astype(np.float32)
    (gm_r,) = torch.autograd.grad((s_r * cpu(up)).sum(), [m])
    tab = ops.gen_tables(dev(te), 1.5)
    assert_close(host(ops.ideal_fwd(L.MODEL_WFPM, dev(maps), tab, ne)), host(s_r), TOL, "TE-augmented forward")
    assert_close(host(ops.ideal_bwd(L.MODEL_WFPM, dev(maps), tab, ne, dev(up))), host(gm_r), TOL, "TE-augmented adjoint")
    acqs = synth.add_noise(host(s_r), rng)
    pm = np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.95)
    a, p = cpu(acqs, True), cpu(pm, True)
    rho_r = orc.get_rho(a, p, te=cpu(te))
    up_rho = rng.standard_normal(rho_r.shape).astype(np.float32)
    ga_r, gp_r = torch.autograd.grad((rho_r * cpu(up_rho)).sum(), [a, p])
    rho, _ = ops.get_rho_fwd(dev(acqs), dev(pm), tab)
    assert_close(host(rho), host(rho_r), TOL, "get_rho (train-IDEAL-TEaug.py:304)")
    ga, gp = ops.get_rho_bwd(dev(acqs), dev(pm), tab, dev(up_rho), None)
    assert_close(host(ga), host(ga_r), TOL, "get_rho grad acqs")
    assert_close(host(gp), host(gp_r), TOL, "get_rho grad pm")
    # and the fused objective on the same per-sample tables (36 864 voxels per plane: 72 ring tiles per sample)
    p2 = cpu(pm, True)
    lref, _, _ = orc.physics_loss_a2a(cpu(acqs), p2, te=cpu(te))
    (gref,) = torch.autograd.grad(lref, [p2])
    loss, g, _, _ = ops.a2a_loss(dev(acqs), dev(pm), tab)
    assert abs(loss.item() - lref.item()) <= TOL * lref.item()
    assert_close(host(g), host(gref), TOL, "fused objective grad pm")


# ---------------------------------------------------------------------------------------------------------------
# C4: bipolar mag/phase model at the script's own batch (3 slices of 384 x 384 x 6), fused objective and operators
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("masked", [True, False], ids=["disc-masked", "unmasked"])
def test_c4_bipolar_mag_phase_384_vs_oracle(masked):
    nb, H, W, ne = 3, 384, 384, 6
    rng = np.random.default_rng(404 + masked)
    maps = synth.magpha_maps(nb, H, W, rng, bipolar=True, masked=masked)
    te = synth.te_orig(nb, ne)
    m = cpu(maps, True)
    s_r = orc.IDEAL_mag_phase(m, [1.5, cpu(te)])
    up = rng.standard_normal(s_r.shape).astype(np.float32)
    (gm_r,) = torch.autograd.grad((s_r * cpu(up)).sum(), [m])
    tab = ops.gen_tables(dev(te), 1.5)
    assert_close(host(ops.ideal_fwd(L.MODEL_MAGPHA, dev(maps), tab, ne)), host(s_r), TOL, "IDEAL_mag_phase forward")
    assert_close(host(ops.ideal_bwd(L.MODEL_MAGPHA, dev(maps), tab, ne, dev(up))), host(gm_r), TOL, "IDEAL_mag_phase adjoint")
    acqs = synth.add_noise(host(s_r), rng)
    acqs[1, 2, 250, 250, 0] = 0.0                            # a single zeroed component inside the object
    est = (maps + 0.02 * rng.standard_normal(maps.shape).astype(np.float32) * (maps != 0)).astype(np.float32)
    for rdtype, tol_l in ((torch.float32, TOL), (torch.float64, 2e-6)):
        e = cpu(est).to(rdtype).requires_grad_(True)
        lref, _ = orc.physics_loss_fwd(cpu(acqs), e, cpu(te), model="magpha", rdtype=rdtype)
        (gref,) = torch.autograd.grad(lref, [e])
        loss, gm, _ = ops.ideal_loss(L.MODEL_MAGPHA, dev(est), dev(acqs), tab)
        assert abs(loss.item() - lref.item()) <= tol_l * lref.item(), (loss.item(), lref.item())
        assert_close(host(gm), host(gref.float()), TOL, f"fused bipolar objective grad ({rdtype})")


# ---------------------------------------------------------------------------------------------------------------
# C5: dataset synthesis decoders (IDEAL_mag_Layer on 2-row mag/phase tensors without the bipolar channel, and IDEAL_mag)
# ---------------------------------------------------------------------------------------------------------------
def test_c5_synthesis_decoders_384_vs_oracle():
    nb, H, W, ne = 3, 384, 384, 6
    rng = np.random.default_rng(505)
    te = synth.te_orig(nb, ne)
    tab = ops.gen_tables(dev(te), 1.5)
    mp = synth.magpha_maps(nb, H, W, rng, bipolar=False, masked=False)
    with torch.no_grad():
        ref = orc.IDEAL_mag_Layer()(cpu(mp), te=cpu(te))
    assert_close(host(ops.ideal_fwd(L.MODEL_MAGPHA, dev(mp), tab, ne)), host(ref), TOL, "IDEAL_mag_Layer (2-row)")
    ff = synth.ffpd_maps(nb, H, W, rng)
    with torch.no_grad():
        ref = orc.IDEAL_mag(cpu(ff), [1.5, cpu(te)])
    assert_close(host(ops.ideal_fwd(L.MODEL_FFPD, dev(ff), tab, ne)), host(ref), TOL, "IDEAL_mag")
    with torch.no_grad():
        flat_ref = orc.A_from_MEBCRN(ref)
    flat = ops.ideal_fwd(L.MODEL_FFPD, dev(ff), tab, ne, flags=L.F_FLAT)
    assert_close(host(flat), host(flat_ref), TOL, "IDEAL_mag, interleaved output")


# ---------------------------------------------------------------------------------------------------------------
# the published model's objectives at 384 x 384
# ---------------------------------------------------------------------------------------------------------------
def _uq_inputs(nb, H, W, ne, rng, masked):
    te = synth.te_orig(nb, ne)
    maps, acqs, _ = _c2_inputs(nb, H, W, ne, rng, masked, te)
    pm = (maps[:, 2:3] + 0.03 * rng.standard_normal(maps[:, 2:3].shape).astype(np.float32) * (maps[:, 2:3] != 0)).astype(np.float32)
    tissue = (maps[:, 0:1, :, :, 0:1] != 0).astype(np.float32)
    phi_v = rng.uniform(1e-5, 4e-3, size=(nb, 1, H, W, 1)).astype(np.float32) * tissue
    r2_m = np.ascontiguousarray(pm[..., 1:2])
    r2_v = rng.uniform(1e-5, 3e-3, size=(nb, 1, H, W, 1)).astype(np.float32) * tissue
    return te, acqs, pm, phi_v, r2_m, r2_v


@pytest.mark.parametrize("masked", [True, False], ids=["disc-masked", "unmasked"])
def test_uq_objective_384_vs_fp64_oracle(masked):
    nb, H, W, ne = 2, 384, 384, 6
    rng = np.random.default_rng(606 + masked)
    te, acqs, pm, phi_v, r2_m, r2_v = _uq_inputs(nb, H, W, ne, rng, masked)
    dt = torch.float64
    p, pv, rm, rv = (cpu(x).to(dt).requires_grad_(True) for x in (pm, phi_v, r2_m, r2_v))
    lref, rho_r, _, _ = orc.physics_loss_a2a_uq(cpu(acqs), p, pv, rm, rv, te=cpu(te), rdtype=dt)
    gref = torch.autograd.grad(lref, [p, pv, rm, rv])
    tab = ops.gen_tables(dev(te), 1.5)
    loss, g_pm, g_pv, g_rm, g_rv, rho = ops.a2a_uq_loss(dev(acqs), dev(pm), dev(phi_v), dev(r2_m), dev(r2_v), tab, want_rho=True)
    assert abs(loss.item() - lref.item()) <= TOL * abs(lref.item())
    assert_close(host(rho), host(rho_r.float()), TOL, "rho")
    # gradient tolerance as in test_parity_gpu.py::test_uq_objective_vs_fp64_oracle (fp32 1 - exp(-x) on the kernel side)
    for got, want, what in zip((g_pm, g_pv, g_rm, g_rv), gref, ("pm", "phi var", "r2 mean", "r2 var")):
        assert_close(host(got).reshape(want.shape), host(want.float()), 2e-5, "grad " + what)


def test_rician_objective_384_vs_fp64_oracle():
    nb, H, W, ne = 2, 384, 384, 6
    rng = np.random.default_rng(707)
    te, acqs, pm, phi_v, r2_m, r2_v = _uq_inputs(nb, H, W, ne, rng, True)
    dt = torch.float64
    p, pv, rm, rv = (cpu(x).to(dt).requires_grad_(True) for x in (pm, phi_v, r2_m, r2_v))
    lref = orc.physics_loss_a2a_rician(cpu(acqs), p, pv, rm, rv, te=cpu(te), rdtype=dt)[0]
    gref = torch.autograd.grad(lref, [p, pv, rm, rv])
    gref = [torch.nan_to_num(g, nan=0.0) for g in gref]
    tab = ops.gen_tables(dev(te), 1.5)
    loss, g_pm, g_pv, g_rm, g_rv, _ = ops.a2a_rician_loss(dev(acqs), dev(pm), dev(phi_v), dev(r2_m), dev(r2_v), tab)
    assert abs(loss.item() - lref.item()) <= TOL * abs(lref.item()), (loss.item(), lref.item())
    for got, want, what in zip((g_pm, g_pv, g_rm, g_rv), gref, ("pm", "phi var", "r2 mean", "r2 var")):
        assert_close(host(got).reshape(want.shape), host(want.float()), 5e-5, "grad " + what)
