"""Generate tests/golden/*.npz by running the UNMODIFIED reference source.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (it needs /root/reference):

    python oracle/gen_golden.py

How: `oracle/tf_shim` (a torch-CPU stand-in for the tf.* ops the reference touches) is put first on
sys.path, then /root/reference, and the reference's own `wflib` and `tf2gan/loss.py` are imported and
executed on seeded inputs from `idealgan.synth`.  Gradients are torch autograd through the reference's
own code.  Every fixture stores inputs, outputs and gradients as float32/complex64 arrays; sizes are
tiny (H = W = 12) so the fixtures stay a few hundred KB in total.

The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("IDEALGAN_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, REF)
sys.path.insert(2, os.path.join(ROOT, "ideal-gan_b200"))

import torch  # noqa: E402
import tensorflow as tf_shim  # noqa: E402  (oracle/tf_shim)
import wflib as wf  # noqa: E402  (the reference's package, running on the shim)
from idealgan import synth  # noqa: E402

assert os.path.realpath(wf.__file__).startswith(os.path.realpath(REF)), wf.__file__

_spec = importlib.util.spec_from_file_location("ref_tf2gan_loss", os.path.join(REF, "tf2gan", "loss.py"))
ref_loss = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_loss)

OUT = os.path.join(ROOT, "tests", "golden")
H = W = 12


def T(a, grad=False):
    t = tf_shim.convert_to_tensor(np.ascontiguousarray(a))      # shim tensor: immutable-style `x *= y`
    return t.requires_grad_(True) if grad else t


def N(t):
    return t.detach().as_subclass(torch.Tensor).numpy().copy()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {len(arrays)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


def vjp(outputs, inputs, rng):
    """Random upstream gradients (one per output) and the resulting input gradients."""
    ups = [rng.standard_normal(tuple(o.shape)).astype(np.float32) for o in outputs]
    loss = sum((o * T(u)).sum() for o, u in zip(outputs, ups))
    grads = torch.autograd.grad(loss, inputs, allow_unused=True)
    grads = [N(g) if g is not None else np.zeros(tuple(i.shape), np.float32) for g, i in zip(grads, inputs)]
    return ups, grads


class Moments:
    def __init__(self, mean, variance):
        self._m, self._v = mean, variance

    def mean(self):
        return self._m

    def variance(self):
        return self._v


def gen_tables():
    out = {}
    out["te_orig6"] = N(wf.gen_TEvar(6, 2, orig=True))
    out["te_orig3"] = N(wf.gen_TEvar(3, 1, orig=True))
    out["te_orig12"] = N(wf.gen_TEvar(12, 1, orig=True))
    out["te_3T"] = N(wf.gen_TEvar(6, 2, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None))
    np.random.seed(7)
    out["te_rand_seed7"] = N(wf.gen_TEvar(8, 3))
    np.random.seed(7)
    out["te_rand_bip_seed7"] = N(wf.gen_TEvar(6, 2, TE_ini_d=0.4e-3, d_TE_min=1.0e-3, d_TE_d=0.3e-3))
    rng = np.random.default_rng(11)
    cases = {"orig6_1p5": (synth.te_orig(2, 6), 1.5), "rand6_3p0": (synth.te_random(3, 6, rng), 3.0),
             "rand12_1p5": (synth.te_random(2, 12, rng), 1.5), "rand3_1p5": (synth.te_random(2, 3, rng), 1.5)}
    for k, (te, field) in cases.items():
        M, Mp, Hp = wf.gen_M(T(te), field=field, get_H=True)
        _, P0, _ = wf.gen_M(T(te), field=field, get_P0=True)
        A, Ap, AtAp = wf.gen_A(M, gen_AtA_pinv=True)
        out[f"{k}_te"], out[f"{k}_field"] = te, np.float32(field)
        out[f"{k}_M"], out[f"{k}_Mpinv"], out[f"{k}_Hpinv"], out[f"{k}_P0"] = N(M), N(Mp), N(Hp), N(P0)
        out[f"{k}_A"], out[f"{k}_Apinv"], out[f"{k}_AtApinv"] = N(A), N(Ap), N(AtAp)
    X = rng.uniform(0.0, 1.0, size=(2, 40, 3)).astype(np.float32)
    X[0, 0] = (0.7, 0.0, 0.0)      # pure water quirk (Appendix A.7)
    X[0, 1] = (0.0, 0.0, 0.0)      # background quirk
    xy, ratio = wf.eigenvals(T(X))
    out["eig_X"], out["eig_xy"], out["eig_ratio"] = X, N(xy), N(ratio)
    save("tables", **out)


def gen_forward():
    rng = np.random.default_rng(0)
    out = {}
    # (name, nb, ne, bipolar, te kind, field, r2_sc)
    cases = [("wfpm_orig6", 2, 6, False, "orig", 1.5, 200.0), ("wfpm_bip_rand6", 3, 6, True, "rand", 3.0, 150.0),
             ("wfpm_rand3", 2, 3, False, "rand", 1.5, 200.0), ("wfpm_bip_rand12", 2, 12, True, "rand", 1.5, 200.0)]
    for name, nb, ne, bip, kind, field, r2 in cases:
        maps = synth.wfpm_maps(nb, H, W, rng, bipolar=bip)
        te = synth.te_orig(nb, ne) if kind == "orig" else synth.te_random(nb, ne, rng)
        m = T(maps, grad=True)
        y = wf.IDEAL_Layer(field=field, r2_sc=r2)(m, te=T(te), training=False)
        (up,), (g,) = vjp([y], [m], rng)
        out.update({f"{name}_maps": maps, f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_r2sc": np.float32(r2), f"{name}_out": N(y), f"{name}_up": up, f"{name}_gmaps": g})
    # default-te path of the layer (ne kwarg)
    maps = synth.wfpm_maps(1, H, W, rng)
    out["wfpm_default_maps"] = maps
    out["wfpm_default_out"] = N(wf.IDEAL_Layer()(T(maps), ne=4))
    # IDEAL_mag (ff / pd / phase)
    for name, nb, ne, kind, field in [("ffpd_orig6", 2, 6, "orig", 1.5), ("ffpd_rand5", 2, 5, "rand", 3.0)]:
        maps = synth.ffpd_maps(nb, H, W, rng)
        te = synth.te_orig(nb, ne) if kind == "orig" else synth.te_random(nb, ne, rng)
        m = T(maps, grad=True)
        y = wf.IDEAL_mag_Layer(field=field)(m, te=T(te), training=False)
        (up,), (g,) = vjp([y], [m], rng)
        out.update({f"{name}_maps": maps, f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_out": N(y), f"{name}_up": up, f"{name}_gmaps": g})
    # IDEAL_mag_phase (bipolar, 4 channels)
    for name, nb, ne, kind, field in [("magpha_rand6", 3, 6, "bip", 1.5), ("magpha_orig4", 1, 4, "orig", 3.0)]:
        maps = synth.magpha_maps(nb, H, W, rng, bipolar=True)
        te = (synth.te_orig(nb, ne) if kind == "orig" else
              synth.te_random(nb, ne, rng, te_ini_d=0.4e-3, d_te_min=1.0e-3, d_te_d=0.3e-3))
        m = T(maps, grad=True)
        y = wf.IDEAL_mag_Layer(field=field, sep_phase=True)(m, T(te), training=False)
        (up,), (g,) = vjp([y], [m], rng)
        out.update({f"{name}_maps": maps, f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_out": N(y), f"{name}_up": up, f"{name}_gmaps": g})
    save("forward", **out)


def _acqs_from(maps, te, field, rng, r2=200.0, sigma=0.02):
    with torch.no_grad():
        clean = N(wf.IDEAL_Layer(field=field, r2_sc=r2)(T(maps), te=T(te)))
    return synth.add_noise(clean, rng, sigma=sigma)


def gen_solve():
    rng = np.random.default_rng(1)
    out = {}
    # get_rho, MEBCRN layout
    for name, nb, ne, field, r2, pc in [("rho_orig6", 2, 6, 1.5, 200.0, False), ("rho_rand6_pc", 2, 6, 3.0, 200.0, True),
                                        ("rho_rand9", 2, 9, 1.5, 150.0, False)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = synth.te_orig(nb, ne) if "orig" in name else synth.te_random(nb, ne, rng)
        acqs = _acqs_from(maps, te, field, rng, r2)
        a, p = T(acqs, grad=True), T(maps[:, 2:3].copy(), grad=True)
        rho, dem = wf.get_rho(a, p, field=field, te=T(te), r2_sc=r2, phase_constraint=pc, acq_demod=True)
        ups, gs = vjp([rho, dem], [a, p], rng)
        out.update({f"{name}_acqs": acqs, f"{name}_pm": maps[:, 2:3], f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_r2sc": np.float32(r2), f"{name}_pc": np.bool_(pc), f"{name}_rho": N(rho), f"{name}_demod": N(dem),
                    f"{name}_up_rho": ups[0], f"{name}_up_demod": ups[1], f"{name}_gacqs": gs[0], f"{name}_gpm": gs[1]})
    # get_rho with a bipolar row: literal reference semantics -- row 0 is the (phi, R2*) pair, the LAST row
    # carries the bipolar phase when the tensor has more than 3 rows (IDEAL_model.py:556-557,567-568)
    nb, ne, field = 2, 6, 1.5
    maps = synth.wfpm_maps(nb, H, W, rng, bipolar=True, neg_r2_frac=0.0)
    te = synth.te_random(nb, ne, rng, te_ini_d=0.4e-3, d_te_min=1.0e-3, d_te_d=0.3e-3)
    acqs = _acqs_from(maps, te, field, rng)
    pm4 = np.ascontiguousarray(maps[:, [2, 0, 1, 3]])
    a, p = T(acqs, grad=True), T(pm4, grad=True)
    rho = wf.get_rho(a, p, field=field, te=T(te))
    ups, gs = vjp([rho], [a, p], rng)
    out.update({"rho_bip_acqs": acqs, "rho_bip_pm": pm4, "rho_bip_te": te, "rho_bip_rho": N(rho),
                "rho_bip_up_rho": ups[0], "rho_bip_gacqs": gs[0], "rho_bip_gpm": gs[1]})
    # flat layout (nb,H,W,2ne) with PM (nb,H,W,2) = (R2*, phi).  The reference's own flat branch cannot run
    # for H > 3: `param_maps.shape[1] > 3` (IDEAL_model.py:567) reads H as a row count and then indexes a
    # 4-D tensor with 5 indices.  Its evident intent (no bipolar term; :548-549,559-560,606-614) is pinned
    # here by running the reference's MEBCRN branch on the same data and re-laying the result out.
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = synth.te_orig(nb, ne)
    acqs = _acqs_from(maps, te, 1.5, rng)
    flat = np.ascontiguousarray(acqs.transpose(0, 2, 3, 1, 4).reshape(nb, H, W, 2 * ne))
    pm_flat = np.ascontiguousarray(np.stack([maps[:, 2, :, :, 1], maps[:, 2, :, :, 0]], axis=-1))
    try:
        wf.get_rho(T(flat), T(pm_flat), MEBCRN=False)
        raise SystemExit("reference flat get_rho unexpectedly works: regenerate this fixture from it")
    except IndexError:
        pass
    a, p = T(acqs, grad=True), T(maps[:, 2:3].copy(), grad=True)
    rho = wf.get_rho(a, p)
    up = rng.standard_normal((nb, H, W, 4)).astype(np.float32)
    up5 = T(np.ascontiguousarray(up.reshape(nb, H, W, 2, 2).transpose(0, 3, 1, 2, 4)))
    ga, gp = torch.autograd.grad((rho * up5).sum(), [a, p])
    out.update({"rho_flat_acqs": flat, "rho_flat_pm": pm_flat,
                "rho_flat_rho": np.ascontiguousarray(N(rho).transpose(0, 2, 3, 1, 4).reshape(nb, H, W, 4)),
                "rho_flat_up_rho": up,
                "rho_flat_gacqs": np.ascontiguousarray(N(ga).transpose(0, 2, 3, 1, 4).reshape(nb, H, W, 2 * ne)),
                "rho_flat_gpm": np.ascontiguousarray(np.stack([N(gp)[:, 0, :, :, 1], N(gp)[:, 0, :, :, 0]], axis=-1))})
    # acq_to_acq (library single-tensor form) with explicit and default echo times
    for name, nb, ne, field, te_kind in [("a2a_orig6", 2, 6, 1.5, None), ("a2a_3T", 2, 6, 3.0, None),
                                         ("a2a_rand7", 3, 7, 1.5, "rand")]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = synth.te_random(nb, ne, rng) if te_kind else N(
            wf.gen_TEvar(ne, nb, orig=True) if field == 1.5 else
            wf.gen_TEvar(ne, nb, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None))
        acqs = _acqs_from(maps, te, field, rng)
        pm = maps[:, 2:3].copy()
        pm += rng.uniform(-0.05, 0.05, size=pm.shape).astype(np.float32) * (pm != 0)   # imperfect estimate
        a, p = T(acqs, grad=True), T(pm, grad=True)
        y = wf.acq_to_acq(a, p, te=T(te) if te_kind else None, field=field)
        ups, gs = vjp([y], [a, p], rng)
        # config-2 objective on the reference's own ops (train-IDEAL-unsup.py:216-218,236)
        a2, p2 = T(acqs), T(pm, grad=True)
        y2 = wf.acq_to_acq(a2, p2, te=T(te) if te_kind else None, field=field)
        y2 = torch.where(a2 != 0.0, y2, torch.zeros_like(y2))
        loss = ref_loss.tf.losses.MeanSquaredError()(a2, y2)
        (gl,) = torch.autograd.grad(loss, [p2])
        out.update({f"{name}_acqs": acqs, f"{name}_pm": pm, f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_out": N(y), f"{name}_up": ups[0], f"{name}_gacqs": gs[0], f"{name}_gpm": gs[1],
                    f"{name}_loss": np.float32(loss.item()), f"{name}_loss_gpm": N(gl)})
    save("solve", **out)


def gen_losses():
    """Config-4 objective and the UQ losses on the reference's own ops."""
    rng = np.random.default_rng(2)
    out = {}
    nb, ne = 3, 6
    maps = synth.magpha_maps(nb, H, W, rng, bipolar=True)
    te = synth.te_random(nb, ne, rng, te_ini_d=0.4e-3, d_te_min=1.0e-3, d_te_d=0.3e-3)
    op = wf.IDEAL_mag_Layer(sep_phase=True)
    with torch.no_grad():
        acqs = synth.add_noise(N(op(T(maps), T(te))), rng)
    est = maps + rng.uniform(-0.03, 0.03, size=maps.shape).astype(np.float32) * (maps != 0)
    m, a = T(est, grad=True), T(acqs)
    y = op(m, T(te), training=False)
    y = torch.where(a != 0.0, y, torch.zeros_like(y))
    loss = ref_loss.tf.losses.MeanSquaredError()(a, y)
    (g,) = torch.autograd.grad(loss, [m])
    out.update({"c4_maps": est, "c4_te": te, "c4_acqs": acqs, "c4_loss": np.float32(loss.item()), "c4_gmaps": N(g)})
    # same with the WF-PM forward model (IDEAL_Layer) incl. bipolar row
    maps = synth.wfpm_maps(nb, H, W, rng, bipolar=True)
    te = synth.te_random(nb, ne, rng)
    opw = wf.IDEAL_Layer(field=3.0)
    with torch.no_grad():
        acqs = synth.add_noise(N(opw(T(maps), te=T(te))), rng)
    est = maps + rng.uniform(-0.03, 0.03, size=maps.shape).astype(np.float32) * (maps != 0)
    m, a = T(est, grad=True), T(acqs)
    y = torch.where(a != 0.0, opw(m, te=T(te)), torch.zeros_like(a))
    loss = ref_loss.tf.losses.MeanSquaredError()(a, y)
    (g,) = torch.autograd.grad(loss, [m])
    out.update({"wl_maps": est, "wl_te": te, "wl_acqs": acqs, "wl_loss": np.float32(loss.item()), "wl_gmaps": N(g)})
    # VarMeanSquaredError / VarMeanSquaredErrorR2 (tf2gan/loss.py:130-162)
    yt = rng.uniform(0.0, 1.0, size=(2, 6, H, W, 2)).astype(np.float32)
    yp = (yt + rng.normal(0, 0.05, size=yt.shape)).astype(np.float32)
    var = rng.uniform(0.0, 0.02, size=yt.shape).astype(np.float32)
    var[0, 0] = 0.0                                           # exercises the 1e-5 floor
    ypv = T(np.concatenate([yp, var], axis=-1), grad=True)
    l1 = ref_loss.VarMeanSquaredError()(T(yt), ypv)
    (g1,) = torch.autograd.grad(l1, [ypv])
    yt1 = yt[..., :1].copy()
    yt1[0, 0, :2] = 0.0
    ypv2 = T(np.concatenate([np.abs(yp[..., :1]), var[..., :1]], axis=-1), grad=True)
    l2 = ref_loss.VarMeanSquaredErrorR2()(T(yt1), ypv2)
    (g2,) = torch.autograd.grad(l2, [ypv2])
    out.update({"vm_true": yt, "vm_predvar": N(ypv), "vm_loss": np.float32(l1.item()), "vm_grad": N(g1),
                "vr_true": yt1, "vr_predvar": N(ypv2), "vr_loss": np.float32(l2.item()), "vr_grad": N(g2)})
    save("losses", **out)


def gen_tier2():
    rng = np.random.default_rng(3)
    out = {}
    nb, ne = 2, 6
    # CSE_mag on noiseless magnitudes (keeps fit > 1e-6 inside the disc, Appendix A.7)
    for name, field, r2sc in [("cse_1p5", 1.5, 200.0), ("cse_3p0", 3.0, 150.0)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0, masked=False)
        maps[:, 0:2, :, :, 1] = 0.0                          # in-phase water/fat so that magnitudes are informative
        maps[:, 0:2, :, :, 0] = np.abs(maps[:, 0:2, :, :, 0]) + 0.1
        te = synth.te_random(nb, ne, rng)
        with torch.no_grad():
            sig = N(wf.IDEAL_Layer(field=field, r2_sc=r2sc)(T(maps), te=T(te)))
        mag = np.sqrt((sig ** 2).sum(-1, keepdims=True)).astype(np.float32)
        r2 = maps[:, 2:3, :, :, 1:2].copy()
        a, r = T(mag, grad=True), T(r2, grad=True)
        rho, fit, demod, ls = wf.CSE_mag(a, r, [field, T(te)], r2_sc=r2sc, demod_signal=True)
        _, _, unc, _ = wf.CSE_mag(a, r, [field, T(te)], r2_sc=r2sc, uncertainty=True)
        ups, gs = vjp([rho, fit, demod, ls, unc], [a, r], rng)
        out.update({f"{name}_mag": mag, f"{name}_r2": r2, f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_r2sc": np.float32(r2sc), f"{name}_rho": N(rho), f"{name}_fit": N(fit), f"{name}_demod": N(demod),
                    f"{name}_ls": N(ls), f"{name}_unc": N(unc), f"{name}_gmag": gs[0], f"{name}_gr2": gs[1]})
        for i, k in enumerate(["rho", "fit", "demod", "ls", "unc"]):
            out[f"{name}_up_{k}"] = ups[i]
    # acq_uncertainty / PDFF_uncertainty with moment holders in place of tfp distributions
    for name, field, rem in [("unc_1p5", 1.5, False), ("unc_3p0_rem", 3.0, True)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = synth.te_random(nb, ne, rng)
        phi_m = maps[:, 2:3, :, :, 0:1].copy()
        r2_m = maps[:, 2:3, :, :, 1:2].copy()
        phi_v = rng.uniform(1e-5, 2e-3, size=phi_m.shape).astype(np.float32)
        r2_v = rng.uniform(1e-5, 2e-3, size=phi_m.shape).astype(np.float32)
        pv, rm, rv = T(phi_v, grad=True), T(r2_m, grad=True), T(r2_v, grad=True)
        var = wf.acq_uncertainty(T(maps[:, :2].copy()), Moments(T(phi_m), pv), Moments(rm, rv), ne=ne, te=T(te),
                                 field=field, rem_R2=rem)
        var1 = wf.acq_uncertainty(T(maps[:, :2].copy()), Moments(T(phi_m), pv), Moments(rm, rv), ne=ne, te=T(te),
                                  field=field, rem_R2=rem, only_mag=True)
        ups, gs = vjp([var], [pv, rm, rv], rng)
        out.update({f"{name}_rho": maps[:, :2], f"{name}_te": te, f"{name}_field": np.float32(field),
                    f"{name}_rem": np.bool_(rem), f"{name}_phi_m": phi_m, f"{name}_phi_v": phi_v, f"{name}_r2_m": r2_m,
                    f"{name}_r2_v": r2_v, f"{name}_var": N(var), f"{name}_var_mag": N(var1), f"{name}_up": ups[0],
                    f"{name}_g_phi_v": gs[0], f"{name}_g_r2_m": gs[1], f"{name}_g_r2_v": gs[2]})
    for name, rem in [("pdffu", False), ("pdffu_rem", True)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0, masked=False)
        te = synth.te_random(nb, ne, rng)
        acqs = _acqs_from(maps, te, 1.5, rng)
        phi_m = maps[:, 2:3, :, :, 0:1].copy()
        r2_m = maps[:, 2:3, :, :, 1:2].copy()
        phi_v = rng.uniform(1e-5, 2e-3, size=phi_m.shape).astype(np.float32)
        r2_v = rng.uniform(1e-5, 2e-3, size=phi_m.shape).astype(np.float32)
        with torch.no_grad():
            rho, rvar = wf.PDFF_uncertainty(T(acqs), Moments(T(phi_m), T(phi_v)), Moments(T(r2_m), T(r2_v)),
                                            te=T(te), rem_R2=rem)
        out.update({f"{name}_acqs": acqs, f"{name}_te": te, f"{name}_rem": np.bool_(rem), f"{name}_phi_m": phi_m,
                    f"{name}_phi_v": phi_v, f"{name}_r2_m": r2_m, f"{name}_r2_v": r2_v, f"{name}_rho": N(rho),
                    f"{name}_rho_var": N(rvar)})
    save("tier2", **out)


def gen_uq():
    """The UQ training objective of train-IDEAL-unsup.py:214-231 composed from the reference's own functions:
    get_rho (= the A2B_WF its callers unpack from acq_to_acq), acq_to_acq, acq_uncertainty on the stop-gradient
    estimate, VarMeanSquaredError."""
    rng = np.random.default_rng(4)
    out = {}
    for name, nb, ne, field, rem in [("uq_orig6", 2, 6, 1.5, False), ("uq_rand5_rem", 2, 5, 1.5, True), ("uq_3T", 2, 6, 3.0, False)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
        te = synth.te_orig(nb, ne) if name == "uq_orig6" else synth.te_random(nb, ne, rng)
        if field == 3.0:
            te = (te * np.float32(0.5)).astype(np.float32)
        with torch.no_grad():
            acqs = synth.add_noise(N(wf.IDEAL_Layer(field=field)(T(maps), te=T(te))), rng)
        acqs[0, 1, H // 2, W // 2, 0] = 0.0                                   # one ragged voxel: per-component mask
        pm = (maps[:, 2:3] + 0.03 * rng.standard_normal(maps[:, 2:3].shape).astype(np.float32) * (maps[:, 2:3] != 0)).astype(np.float32)
        tissue = (maps[:, 0:1, :, :, 0:1] != 0).astype(np.float32)
        phi_v = (rng.uniform(1e-5, 4e-3, size=(nb, 1, H, W, 1)).astype(np.float32) * tissue)
        r2_m = np.ascontiguousarray(pm[:, :, :, :, 1:2])
        r2_v = (rng.uniform(1e-5, 3e-3, size=(nb, 1, H, W, 1)).astype(np.float32) * tissue)
        a = T(acqs)
        p, pv, rm, rv = T(pm, grad=True), T(phi_v, grad=True), T(r2_m, grad=True), T(r2_v, grad=True)
        rho = wf.get_rho(a, p, field=field, te=T(te))
        recon = wf.acq_to_acq(a, p, te=T(te), field=field)
        recon = torch.where(a != 0.0, recon, torch.zeros_like(recon))
        var = wf.acq_uncertainty(rho.detach(), Moments(None, pv), Moments(rm, rv), ne=ne, te=T(te), field=field, rem_R2=rem)
        loss = ref_loss.VarMeanSquaredError()(a, tf_shim.concat([recon, var], axis=-1))
        grads = torch.autograd.grad(loss, [p, pv, rm, rv], allow_unused=True)
        g = [N(x) if x is not None else np.zeros(tuple(t.shape), np.float32) for x, t in zip(grads, [p, pv, rm, rv])]
        out.update({f"{name}_acqs": acqs, f"{name}_te": te, f"{name}_field": np.float32(field), f"{name}_rem": np.bool_(rem),
                    f"{name}_pm": pm, f"{name}_phi_v": phi_v, f"{name}_r2_m": r2_m, f"{name}_r2_v": r2_v,
                    f"{name}_loss": np.float32(loss.item()), f"{name}_var": N(var), f"{name}_rho": N(rho),
                    f"{name}_gpm": g[0], f"{name}_gphi_v": g[1], f"{name}_gr2_m": g[2], f"{name}_gr2_v": g[3]})
    save("uq", **out)


def gen_rician():
    """The R2* stage objective of train-IDEAL-unsup.py:267-292 from the reference's own functions.  The library's acq_to_acq
    has no only_mag argument: its callers' |S_hat| is formed here as they form |A| (:269), sqrt(sum(square))."""
    rng = np.random.default_rng(6)
    out = {}
    for name, nb, ne, field, rem in [("ric_orig6", 2, 6, 1.5, False), ("ric_rand5_rem", 2, 5, 1.5, True)]:
        maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0, masked=False)          # no background: sqrt'(0) is NaN in autodiff
        te = synth.te_orig(nb, ne) if name == "ric_orig6" else synth.te_random(nb, ne, rng)
        with torch.no_grad():
            acqs = synth.add_noise(N(wf.IDEAL_Layer(field=field)(T(maps), te=T(te))), rng)
        acqs[0, 1, 2:4, 3, 0] = 0.0                                                   # real channel zero: masked elements
        pm = (maps[:, 2:3] + 0.03 * rng.standard_normal(maps[:, 2:3].shape).astype(np.float32)).astype(np.float32)
        phi_v = rng.uniform(1e-5, 4e-3, size=(nb, 1, H, W, 1)).astype(np.float32)
        r2_m = np.ascontiguousarray(pm[:, :, :, :, 1:2])
        r2_v = rng.uniform(1e-5, 3e-3, size=(nb, 1, H, W, 1)).astype(np.float32)
        phi_v[1, 0, :3] = 0.0                                                          # variance floor
        r2_v[1, 0, :3] = 0.0
        a = T(acqs)
        p, pv, rm, rv = T(pm, grad=True), T(phi_v, grad=True), T(r2_m, grad=True), T(r2_v, grad=True)
        rho = wf.get_rho(a, p, field=field, te=T(te))
        recon = wf.acq_to_acq(a, p, te=T(te), field=field)
        mag = tf_shim.math.sqrt(tf_shim.reduce_sum(tf_shim.square(recon), axis=-1, keepdims=True))
        mag_m = torch.where(a[..., :1] != 0.0, mag, torch.zeros_like(mag))
        var = wf.acq_uncertainty(rho.detach(), Moments(None, pv), Moments(rm, rv), ne=ne, te=T(te), field=field, rem_R2=rem, only_mag=True)
        y = tf_shim.math.sqrt(tf_shim.reduce_sum(tf_shim.square(a), axis=-1, keepdims=True))
        loss = ref_loss.VarMeanSquaredErrorR2()(y, tf_shim.concat([mag_m, var], axis=-1))
        grads = torch.autograd.grad(loss, [p, pv, rm, rv], allow_unused=True)
        g = [N(x) if x is not None else np.zeros(tuple(t.shape), np.float32) for x, t in zip(grads, [p, pv, rm, rv])]
        assert all(np.isfinite(x).all() for x in g)
        out.update({f"{name}_acqs": acqs, f"{name}_te": te, f"{name}_field": np.float32(field), f"{name}_rem": np.bool_(rem),
                    f"{name}_pm": pm, f"{name}_phi_v": phi_v, f"{name}_r2_m": r2_m, f"{name}_r2_v": r2_v,
                    f"{name}_loss": np.float32(loss.item()), f"{name}_var": N(var), f"{name}_mag": N(mag),
                    f"{name}_gpm": g[0], f"{name}_gphi_v": g[1], f"{name}_gr2_m": g[2], f"{name}_gr2_v": g[3]})
    save("rician", **out)


def _reference_data_functions(names):
    """data.py cannot be imported here (h5py, pydicom, nibabel, skimage are absent), so the SOURCE of the named top-level
    functions is cut out of /root/reference/data.py with `ast` and executed unmodified against the shim."""
    import ast
    src = open(os.path.join(REF, "data.py")).read()
    ns = {"tf": tf_shim, "np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, "data.py"), "exec"), ns)
    return [ns[n] for n in names]


def gen_layout():
    """data.A_from_MEBCRN / B_from_MEBCRN / B_to_MEBCRN (data.py:262-329)."""
    A_from, B_from, B_to = _reference_data_functions(["A_from_MEBCRN", "B_from_MEBCRN", "B_to_MEBCRN"])
    rng = np.random.default_rng(5)
    out = {}
    for ne in (3, 6):
        a = rng.standard_normal((2, ne, H, W, 2)).astype(np.float32)
        out[f"a{ne}_in"] = a
        out[f"a{ne}_flat"] = N(A_from(T(a)))
    b = rng.standard_normal((2, 3, H, W, 2)).astype(np.float32)
    out["b_in"] = b
    out["b_flat"] = N(B_from(T(b)))
    for ch in (3, 4):
        bm = rng.uniform(-0.5, 0.5, size=(2, 2, H, W, ch)).astype(np.float32)
        out[f"bmp{ch}_in"] = bm
        out[f"bmp{ch}_flat"] = N(B_from(T(bm), mag_and_phase=True))
        out[f"bmp{ch}_flat_c1"] = N(B_from(T(bm), mag_and_phase=True, c_pha=1))
    for mode, c in (("All", 6), ("WF-PM", 4), ("WF", 2), ("PM", 2)):
        f = rng.standard_normal((2, H, W, c)).astype(np.float32)
        key = mode.replace("-", "")
        out[f"to_{key}_in"] = f
        out[f"to_{key}_out"] = N(B_to(T(f), mode=mode))
    save("layout", **out)


def _reference_script_lines(script, first_tokens):
    """Cut single statements out of a reference SCRIPT (which cannot be imported: it parses argv and builds networks) by the
    text each starts with, in file order, dedented, for `exec` on the shim.  A token that is missing raises (reference drift)."""
    import textwrap
    lines = open(os.path.join(REF, script)).read().splitlines()
    out, pos = [], 0
    for tok in first_tokens:
        for k in range(pos, len(lines)):
            if lines[k].strip().startswith(tok):
                out.append(lines[k].strip())
                pos = k + 1
                break
        else:
            raise RuntimeError(f"{script}: no statement starting with {tok!r} after line {pos}")
    return textwrap.dedent("\n".join(out))


def gen_regs():
    """Regularisers of train-IDEAL-mag.py:288-289,304,308-316 and the map assembly of ROI-analysis.py:301-322: the scripts' own
    statements executed on the shim, inputs from the reference's CSE_mag / PDFF_uncertainty on synthetic data."""
    import types as _types
    rng = np.random.default_rng(7)
    out = {}
    src = _reference_script_lines("train-IDEAL-mag.py", [
        "R2_TV = tf.reduce_sum(tf.image.total_variation(", "G_loss += R2_TV * args.R2_TV_weight", "Ad_aux = tf.reshape(A_demod",
        "Ad_TV = tf.reduce_sum(", "LS_NZ = tf.reduce_sum(", "WF_NZ = tf.reduce_sum(", "aux_cond = tf.square(", "LS_cond = tf.reduce_sum(",
        "G_loss += Ad_TV * args.A_demod_TV_weight"])
    for name, nb, ne, hh, ww in [("reg6", 2, 6, H, W), ("reg3_odd", 3, 3, 7, 9)]:
        maps = synth.wfpm_maps(nb, hh, ww, rng, neg_r2_frac=0.0, masked=(name == "reg6"))
        te = synth.te_orig(nb, ne)
        with torch.no_grad():
            acqs = synth.add_noise(N(wf.IDEAL_Layer(field=1.5)(T(maps), te=T(te))), rng)
            if name == "reg6":
                acqs *= (maps[:, :1, :, :, :1] != 0)
            mag = np.sqrt((acqs ** 2).sum(-1, keepdims=True)).astype(np.float32)
            r2 = np.ascontiguousarray(maps[:, 2:3, :, :, 1:2])
            _, fit, demod, ls = wf.CSE_mag(T(mag), T(r2), [1.5, T(te)], demod_signal=True)
        ls = (N(ls) + 0.05 * rng.standard_normal(tuple(ls.shape))).astype(np.float32)     # negatives and both discriminant signs
        demod = N(demod).astype(np.float32)
        weights = (0.3, 1.7, 0.9, 0.6)
        a_ls, a_demod, a_r2 = T(ls, grad=True), T(demod, grad=True), T(r2, grad=True)
        ns_ = {"tf": tf_shim, "A_demod": a_demod, "A2B_ls": a_ls, "A2B2A_mag": T(N(fit)), "R2_TV_aux": a_r2, "G_loss": 0.0,
               "args": _types.SimpleNamespace(A_demod_TV_weight=weights[0], LS_NZ_weight=weights[1], LS_cond_weight=weights[2], R2_TV_weight=weights[3])}
        exec(src, ns_)
        grads = torch.autograd.grad(ns_["G_loss"], [a_ls, a_demod, a_r2])
        out.update({f"{name}_ls": ls, f"{name}_demod": demod, f"{name}_r2": r2, f"{name}_weights": np.float32(weights),
                    f"{name}_sums": np.float32([ns_[k].item() for k in ("Ad_TV", "LS_NZ", "WF_NZ", "LS_cond", "R2_TV")]),
                    f"{name}_total": np.float32(ns_["G_loss"].item()),
                    f"{name}_g_ls": N(grads[0]), f"{name}_g_demod": N(grads[1]), f"{name}_g_r2": N(grads[2])})

    roi_src = _reference_script_lines("ROI-analysis.py", [
        "A2B_WF_abs = tf.math.sqrt(", "A2B_WF_abs = tf.transpose(", "A2B_WFsum_abs = tf.math.sqrt(", "A2B_WFsum_abs = tf.transpose(",
        "A2B_R2 = A2B[:,2,:,:,1:]", "A2B = tf.concat([A2B_WF_abs,A2B_WFsum_abs,A2B_R2],axis=-1)"])
    var_src = _reference_script_lines("ROI-analysis.py", [
        "W_var = tf.abs(", "WF_var = tf.abs(", "F_var = tf.abs(", "PDFF_var = W_var/(", "PDFF_var -= 2 * WF_var", "PDFF_var += (W_var + F_var",
        "PDFF_var *= A2B_WF_abs", "A2B = tf.concat([A2B,PDFF_var],axis=-1)"])
    nb = 2
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0, masked=False)
    var = rng.uniform(1e-6, 1e-3, size=(nb, 5, H, W, 2)).astype(np.float32)
    var[:, :4, :, :, 1] = 0.0                                           # ROI-analysis.py:244: covariance rows are zero-padded
    var[0, :, :2] = 1e-10                                               # the background fill of :248
    with torch.no_grad():
        ns_ = {"tf": tf_shim, "A2B": T(maps), "A2B_var": T(var)}
        exec(roi_src, ns_)
        out.update({"roi_maps": maps, "roi_var": var, "roi_out4": N(ns_["A2B"])})
        exec(var_src, ns_)
        out["roi_out5"] = N(ns_["A2B"])
    save("regs", **out)


def gen_ldm():
    """The per-slice images of gen_LDM_dataset.py:216-218,225-227,234-237 (PDFF, R2*, multi-echo magnitudes, each clipped to
    [0, 1]): the script's own statements on the shim.  Its forward call `IDEAL_op(Z2B)` cannot run in the reference for the
    2-row x 3-channel tensor the script builds (IDEAL_mag_Layer() dispatches to IDEAL_mag, and IDEAL_mag_phase
    needs a 4th channel); the signals are therefore produced by the reference's IDEAL_mag_phase with a zero bipolar channel
    appended -- the semantics the drop-in gives the 3-channel tensor."""
    rng = np.random.default_rng(9)
    src = _reference_script_lines("gen_LDM_dataset.py", [
        "X1 = tf.squeeze(Z2B[i,0,:,:,1]/(", "X1 = tf.clip_by_value(X1", "X2 = tf.squeeze(Z2B[i,0,:,:,2]", "X2 = tf.clip_by_value(X2",
        "X3 = tf.math.sqrt(tf.reduce_sum(tf.square(Z2B2A[i,...]),axis=-1))", "X3 = tf.squeeze(X3)", "X3 = tf.clip_by_value(X3"])
    out = {}
    nb, ne = 3, 6
    maps = synth.magpha_maps(nb, H, W, rng, bipolar=False)                      # (nb, 2, H, W, 3), zero outside the disc (0/0 -> NaN PDFF)
    maps[:, 0, :, :, :2] *= 1.6                                                # some magnitudes and signals above 1: the clip matters
    maps[0, 0, :3, :, 2] = -0.2                                                # and some R2* below 0
    te = synth.te_orig(nb, ne)
    maps4 = np.concatenate([maps, np.zeros_like(maps[..., :1])], axis=-1)
    with torch.no_grad():
        sig = wf.IDEAL_mag_Layer(sep_phase=True)(T(maps4), T(te), training=False)
        pdff, r2s, mags = [], [], []
        for i in range(nb):
            ns_ = {"tf": tf_shim, "Z2B": T(maps), "Z2B2A": sig, "i": i}
            exec(src, ns_)
            pdff.append(N(ns_["X1"])); r2s.append(N(ns_["X2"])); mags.append(N(ns_["X3"]))
    out.update({"ldm_maps": maps, "ldm_te": te, "ldm_sig": N(sig), "ldm_pdff": np.stack(pdff), "ldm_r2s": np.stack(r2s),
                "ldm_mag": np.stack(mags)})
    assert np.isnan(out["ldm_pdff"]).any() and (out["ldm_mag"] == 1.0).any() and (out["ldm_r2s"] == 0.0).any()
    save("ldm", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_tables()
    gen_forward()
    gen_solve()
    gen_losses()
    gen_tier2()
    gen_uq()
    gen_rician()
    gen_layout()
    gen_regs()
    gen_ldm()
