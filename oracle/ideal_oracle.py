"""CPU oracle for the IDEAL physics path.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

A restatement, on torch-CPU tensors, of the reference's algorithm in
/root/reference/wflib/IDEAL_model.py (cited per function below) and of the physics losses built on it
(/root/reference/tf2gan/loss.py:130-162, train-IDEAL-unsup.py:214-236, train-IDEAL-single.py:154-157).
It keeps the reference's op structure (batched complex matmul chains over (nb, ne, nv) tensors) so that
it doubles as the "port" CPU baseline timed by bench.py, and it is differentiable through torch autograd
so reference gradients come for free.  `rdtype=torch.float32` follows the reference's complex64
arithmetic (the parity target); `torch.float64` is the tie-breaker truth.

Parity pin: the reference ships no tests or golden vectors for this path, and TensorFlow
cannot be installed here.  The pin used instead is tests/golden/*.npz, produced by oracle/gen_golden.py
running the UNMODIFIED reference source under oracle/tf_shim (a torch-backed stand-in for the handful of
tf.* ops the file uses); tests/test_oracle_golden.py holds this restatement to those vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline legs may import this module.
"""
import numpy as np
import torch

# Reference constants (IDEAL_model.py:5-19)
PPM = (0.0, -3.80, -3.40, -2.60, -1.94, -0.39, 0.60)
AMP_FAT = (0.087, 0.693, 0.128, 0.004, 0.039, 0.048)
GAMMA_HZ_PER_T_PPM = 1e-6 * 42.58e6
fm_sc = 300.0
rho_sc = 1.4
ns = 2


def _cdtype(rdtype):
    return torch.complex64 if rdtype == torch.float32 else torch.complex128


def _t(x, rdtype):
    if isinstance(x, torch.Tensor):
        return x.to(rdtype)
    return torch.as_tensor(np.asarray(x), dtype=rdtype)


# ----------------------------------------------------------------------------------------------
# echo times and per-sample tables
# ----------------------------------------------------------------------------------------------
def gen_TEvar(n_ech, bs=1, orig=False, TE_ini_min=1.0e-3, TE_ini_d=1.4e-3, d_TE_min=1.6e-3, d_TE_d=1.0e-3):
    """IDEAL_model.py:21-45.  One echo train per call, tiled over the batch; numpy global RNG."""
    if orig:
        first, step = 1.3e-3, 2.1e-3
        te = np.arange(first, first + step * (n_ech - 1) + 1e-4, step)
    elif not TE_ini_d and not d_TE_d:
        te = np.arange(TE_ini_min, TE_ini_min + d_TE_min * (n_ech - 1) + 1e-4, d_TE_min)
    else:
        first = TE_ini_min + np.random.uniform(0, TE_ini_d)
        centre = d_TE_min + np.random.uniform(0, d_TE_d)
        steps = np.random.normal(centre, 1e-4, size=(n_ech - 1,))
        te = np.cumsum(np.concatenate((np.array([0.0]), steps))) + first
    te = torch.as_tensor(te, dtype=torch.float32)
    return te[None, :, None].repeat(bs, 1, 1)


def _default_te(ne, nb, field):
    """acq_to_acq / acq_uncertainty default echo trains (IDEAL_model.py:145-149, 713-717)."""
    if field == 3.0:
        return gen_TEvar(ne, bs=nb, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None)
    return gen_TEvar(ne, bs=nb, orig=True)


def model_matrix(te, field=1.5, rdtype=torch.float32):
    """M[b,e,:] = (1, sum_p alpha_p exp(2 pi i te field f_p))   (IDEAL_model.py:10-15,54)."""
    cd = _cdtype(rdtype)
    te = _t(te, rdtype)
    f_p = torch.as_tensor(np.array(PPM) * GAMMA_HZ_PER_T_PPM, dtype=cd)[None, :]      # (1,7), rounded once
    A_p = torch.zeros(7, 2, dtype=cd)
    A_p[0, 0] = 1.0
    A_p[1:, 1] = torch.as_tensor(AMP_FAT, dtype=cd)
    fld = torch.tensor(field, dtype=cd)
    phase = torch.matmul(2j * np.pi * te.to(cd), fld * f_p)                           # (nb,ne,7)
    return torch.matmul(torch.exp(phase), A_p)


def gen_M(te, field=1.5, get_Mpinv=True, get_P0=False, get_H=False, rdtype=torch.float32):
    """IDEAL_model.py:48-77 including its return-arity rules (None for the unsupported combinations)."""
    M = model_matrix(te, field, rdtype)
    ne = M.shape[1]
    Q, R = torch.linalg.qr(M)
    Qh = Q.transpose(1, 2).conj()
    if get_P0:
        P0 = torch.eye(ne, dtype=M.dtype) - torch.matmul(Q, Qh)
        P0 = 0.5 * (P0.transpose(1, 2).conj() + P0)
    if get_Mpinv:
        M_pinv = torch.linalg.solve(R, Qh)
        if get_H:
            H = torch.matmul(M_pinv, M).real
            Qr, Rr = torch.linalg.qr(H)
            H_pinv = torch.linalg.solve(Rr, Qr.transpose(1, 2)).to(M.dtype)
    if get_P0 and get_Mpinv:
        return M, P0, M_pinv
    if get_Mpinv and not get_P0 and not get_H:
        return M, M_pinv
    if get_Mpinv and not get_P0:
        return M, M_pinv, H_pinv
    if not get_Mpinv and not get_P0 and not get_H:
        return M
    return None


def gen_A(M, gen_AtA_pinv=False):
    """Magnitude design matrix and its pseudo-inverse (IDEAL_model.py:80-97)."""
    A = torch.cat([M.abs()[..., :1], M.real[..., 1:], M.abs()[..., 1:] ** 2], dim=-1)
    Q, R = torch.linalg.qr(A)
    A_pinv = torch.linalg.solve(R, Q.transpose(1, 2))
    if gen_AtA_pinv:
        Q2, R2 = torch.linalg.qr(torch.matmul(A.transpose(1, 2), A))
        return A, A_pinv, torch.linalg.solve(R2, Q2.transpose(1, 2))
    return A, A_pinv


def eigenvals(X):
    """Closed-form principal eigenpair of [[a, b/2], [b/2, c]] (IDEAL_model.py:100-138)."""
    a, b, c = X[..., :1], X[..., 1:2], X[..., 2:]
    eps = 1e-12
    hd = (a - c) * 0.5
    hb = b * 0.5
    delta = torch.sqrt(hd * hd + hb * hb + eps)
    lmax = (a + c) * 0.5 + delta
    lmin = (a + c) * 0.5 - delta
    lmax_p = torch.clamp_min(lmax, 0.0)
    lmin_p = torch.clamp_min(lmin, 0.0)
    vx, vy = hb, lmax - a
    norm = torch.sqrt(vx ** 2 + vy ** 2 + eps)
    v = torch.cat([_div_no_nan(vx, norm), _div_no_nan(vy, norm)], dim=-1)
    return torch.sqrt(lmax_p) * v, _div_no_nan(lmin_p, lmax_p)


def _div_no_nan(x, y):
    safe = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(x * safe), x / safe)


# ----------------------------------------------------------------------------------------------
# helpers shared by the operators
# ----------------------------------------------------------------------------------------------
def _xi(phi, r2s, cd):
    """xi = phi + i r2s / (2 pi), flattened to (nb, 1, nv)  (e.g. IDEAL_model.py:242-244)."""
    nb = phi.shape[0]
    xi = torch.complex(phi, r2s / (2 * np.pi)).to(cd)
    return xi.reshape(nb, 1, -1)


def _bipolar_exponent(bip_radians, ne, cd):
    """i * (-1)^e * beta for e = 1..ne  -> (nb, ne, nv)   (IDEAL_model.py:246-253)."""
    nb = bip_radians.shape[0]
    sign = (-1.0) ** torch.arange(1, ne + 1, dtype=bip_radians.dtype)
    col = torch.complex(torch.zeros_like(sign), sign).to(cd)[None, :, None].repeat(nb, 1, 1)
    return torch.matmul(col, bip_radians.reshape(nb, 1, -1).to(cd))


def _split_ri(x, shape):
    x = x.reshape(*shape)
    return torch.stack([x.real, x.imag], dim=-1)


def _modulator(te, xi, sign, cd):
    """exp(sign * 2 pi i te (x) xi) as a K=1 complex matmul, like the reference (IDEAL_model.py:183-184)."""
    te_c = torch.complex(torch.zeros_like(te), te).to(cd)
    return torch.matmul(sign * 2 * np.pi * te_c, xi)


# ----------------------------------------------------------------------------------------------
# forward models
# ----------------------------------------------------------------------------------------------
def IDEAL_model(out_maps, params, r2_sc=200.0, rdtype=torch.float32):
    """Forward model, WF-PM parameterisation (IDEAL_model.py:220-299).
    out_maps (nb, 3|4, H, W, 2); params = [field, te (nb,ne,1)] -> (nb, ne, H, W, 2)."""
    cd = _cdtype(rdtype)
    out_maps = _t(out_maps, rdtype)
    nb, rows, H, W, _ = out_maps.shape
    te = _t(params[1], rdtype)
    ne = te.shape[1]
    M = gen_M(te, field=params[0], get_Mpinv=False, rdtype=rdtype)
    rho = (torch.complex(out_maps[:, :2, :, :, 0], out_maps[:, :2, :, :, 1]) * rho_sc).reshape(nb, ns, -1)
    r2s = torch.relu(out_maps[:, 2, :, :, 1]) * r2_sc
    phi = out_maps[:, 2, :, :, 0] * fm_sc
    expo = _modulator(te, _xi(phi, r2s, cd), +1.0, cd)
    if rows > 3:
        expo = expo + _bipolar_exponent(out_maps[:, -1, :, :, 0] * np.pi, ne, cd)
    S = torch.exp(expo) * torch.matmul(M, rho)
    return _split_ri(S, (nb, ne, H, W))


def IDEAL_mag(out_maps, params, r2_sc=200.0, rdtype=torch.float32):
    """Forward model, PDFF / PD / common-phase parameterisation (IDEAL_model.py:404-453).
    out_maps (nb, 3, H, W, 2): row0=(ff,-) row1=(pd, R2*/r2_sc) row2=(phase/4pi, phi/300)."""
    cd = _cdtype(rdtype)
    out_maps = _t(out_maps, rdtype)
    nb, _, H, W, _ = out_maps.shape
    te = _t(params[1], rdtype)
    ne = te.shape[1]
    M = gen_M(te, field=params[0], get_Mpinv=False, rdtype=rdtype)
    ff = out_maps[:, 0, :, :, 0]
    pd = out_maps[:, 1, :, :, 0]
    r2s = out_maps[:, 1, :, :, 1] * r2_sc
    pha = torch.complex(torch.zeros_like(ff), out_maps[:, 2, :, :, 0]) * np.pi * 4
    phi = out_maps[:, 2, :, :, 1] * fm_sc
    u = torch.exp(pha)
    rho_w = torch.complex((1.0 - ff) * pd * rho_sc, torch.zeros_like(ff)) * u
    rho_f = torch.complex(ff * pd * rho_sc, torch.zeros_like(ff)) * u
    rho = torch.stack([rho_w, rho_f], dim=1).reshape(nb, ns, -1)
    S = torch.exp(_modulator(te, _xi(phi, r2s, cd), +1.0, cd)) * torch.matmul(M, rho)
    return _split_ri(S, (nb, ne, H, W))


def IDEAL_mag_phase(out_maps, params, r2_sc=200.0, rdtype=torch.float32):
    """Forward model, per-species magnitude / phase + bipolar term x 4 pi (IDEAL_model.py:456-509).
    out_maps (nb, 2, H, W, 3|4): row0=(|W|,|F|,R2*/r2_sc[,-]) row1=(pW/4pi,pF/4pi,phi/300[,bip/4pi]).
    The 3-channel (unipolar) case, which the reference cannot evaluate, means bipolar = 0."""
    cd = _cdtype(rdtype)
    out_maps = _t(out_maps, rdtype)
    nb, _, H, W, ch = out_maps.shape
    te = _t(params[1], rdtype)
    ne = te.shape[1]
    M = gen_M(te, field=params[0], get_Mpinv=False, rdtype=rdtype)
    mag = out_maps[:, 0, :, :, :2].permute(0, 3, 1, 2)
    pha = out_maps[:, 1, :, :, :2].permute(0, 3, 1, 2)
    rho = torch.complex(mag, torch.zeros_like(mag)) * rho_sc
    rho = (rho * torch.exp(torch.complex(torch.zeros_like(pha), pha) * 4 * np.pi)).reshape(nb, ns, -1)
    r2s = out_maps[:, 0, :, :, 2] * r2_sc
    phi = out_maps[:, 1, :, :, 2] * fm_sc
    expo = _modulator(te, _xi(phi, r2s, cd), +1.0, cd)
    if ch > 3:
        expo = expo + _bipolar_exponent(out_maps[:, 1, :, :, 3] * 4 * np.pi, ne, cd)
    S = torch.exp(expo) * torch.matmul(M, rho)
    return _split_ri(S, (nb, ne, H, W))


class IDEAL_Layer:
    """IDEAL_model.py:302-311."""

    def __init__(self, field=1.5, r2_sc=200.0):
        self.field, self.r2_sc = field, r2_sc

    def __call__(self, out_maps, te=None, ne=6, training=None, rdtype=torch.float32):
        if te is None:
            te = gen_TEvar(ne, out_maps.shape[0], orig=True)
        return IDEAL_model(out_maps, [self.field, te], r2_sc=self.r2_sc, rdtype=rdtype)


class IDEAL_mag_Layer:
    """IDEAL_model.py:512-524 (always r2_sc = 200).  A 2-row input selects the mag/phase model even
    when sep_phase is False (gen_LDM_dataset.py:156-158 feeds 2x3 tensors)."""

    def __init__(self, field=1.5, sep_phase=False):
        self.field, self.sep_phase = field, sep_phase

    def __call__(self, out_maps, te=None, ne=6, training=None, rdtype=torch.float32):
        if te is None:
            te = gen_TEvar(ne, out_maps.shape[0], orig=True)
        if self.sep_phase or out_maps.shape[1] == 2:
            return IDEAL_mag_phase(out_maps, [self.field, te], rdtype=rdtype)
        return IDEAL_mag(out_maps, [self.field, te], rdtype=rdtype)


# ----------------------------------------------------------------------------------------------
# least-squares water/fat solve and project-resynthesise
# ----------------------------------------------------------------------------------------------
def get_rho(acqs, param_maps, field=1.5, te=None, r2_sc=200.0, phase_constraint=False, MEBCRN=True,
            acq_demod=False, rdtype=torch.float32):
    """LS water/fat solve rho = M^+ (Wm S) / rho_sc  (IDEAL_model.py:527-624)."""
    cd = _cdtype(rdtype)
    acqs = _t(acqs, rdtype)
    param_maps = _t(param_maps, rdtype)
    if MEBCRN:
        nb, ne, H, W, _ = acqs.shape
        S = torch.complex(acqs[..., 0], acqs[..., 1])
    else:
        nb, H, W, two_ne = acqs.shape
        ne = two_ne // 2
        S = torch.complex(acqs[..., 0::2], acqs[..., 1::2]).permute(0, 3, 1, 2)
    if te is None:
        te = gen_TEvar(ne, bs=nb, orig=True)
    te = _t(te, rdtype)
    if phase_constraint:
        M, M_pinv, Hp = gen_M(te, field=field, get_H=True, rdtype=rdtype)
    else:
        M, M_pinv = gen_M(te, field=field, rdtype=rdtype)
    S = S.reshape(nb, ne, -1)
    if MEBCRN:
        r2s = param_maps[:, :1, :, :, 1:] * r2_sc
        phi = param_maps[:, :1, :, :, :1] * fm_sc
    else:                                        # flat layout carries (R2*, phi): order swapped (:559-560)
        r2s = param_maps[:, :, :, :1] * r2_sc
        phi = param_maps[:, :, :, 1:] * fm_sc
    expo = _modulator(te, _xi(phi, r2s, cd), -1.0, cd)
    # Reference tests `param_maps.shape[1] > 3` (:567).  For the flat layout that axis is H, which makes the
    # reference's own flat branch unrunnable for H > 3; its intent (no bipolar term in flat mode) is kept.
    if MEBCRN and param_maps.shape[1] > 3:
        expo = expo - _bipolar_exponent(param_maps[:, -1, :, :, 0] * np.pi, ne, cd)
    WmS = torch.exp(expo) * S
    rho = torch.matmul(M_pinv, WmS)
    if phase_constraint:
        q = torch.matmul(Hp, rho)
        theta = 0.5 * torch.angle(torch.sum(rho * q, dim=1, keepdim=True)).repeat(1, ns, 1)
        rot = torch.exp(torch.complex(torch.zeros_like(theta), -theta))
        mag = torch.matmul(Hp.abs(), (rho * rot).real)
        rho = torch.complex(mag, torch.zeros_like(mag)) * torch.exp(torch.complex(torch.zeros_like(theta), theta))
    rho = rho.reshape(nb, ns, H, W) / rho_sc
    if MEBCRN:
        res = torch.stack([rho.real, rho.imag], dim=-1)
    else:
        rho = rho.permute(0, 2, 3, 1)
        res = torch.stack([rho.real, rho.imag], dim=-1).reshape(nb, H, W, 2 * ns)
    if acq_demod:
        return res, _split_ri(WmS, (nb, ne, H, W))
    return res


def acq_to_acq(acqs, param_maps, te=None, field=1.5, r2_sc=200.0, only_mag=False, legacy_single=False,
               rdtype=torch.float32):
    """Project the measured echoes on the model subspace and resynthesise (IDEAL_model.py:142-200):
    S_hat = Wp M M^+ Wm S.  Returns the callers' 2-tuple (rho_hat / rho_sc, S_hat) -- SURVEY Q1:
    train-IDEAL-unsup.py:214-216,279 -- or, with legacy_single, the library's single S_hat tensor.
    only_mag turns the second result into |S_hat| with one channel."""
    cd = _cdtype(rdtype)
    acqs = _t(acqs, rdtype)
    param_maps = _t(param_maps, rdtype)
    nb, ne, H, W, _ = acqs.shape
    if te is None:
        te = _default_te(ne, nb, field)
    te = _t(te, rdtype)
    M, M_pinv = gen_M(te, field=field, rdtype=rdtype)
    S = torch.complex(acqs[..., 0], acqs[..., 1]).reshape(nb, ne, -1)
    r2s = param_maps[:, 0, :, :, 1] * r2_sc
    phi = param_maps[:, 0, :, :, 0] * fm_sc
    xi = _xi(phi, r2s, cd)
    Wm = torch.exp(_modulator(te, xi, -1.0, cd))
    Wp = torch.exp(_modulator(te, xi, +1.0, cd))
    rho = torch.matmul(M_pinv, Wm * S)
    S_hat = Wp * torch.matmul(M, rho)
    res = _split_ri(S_hat, (nb, ne, H, W))
    if legacy_single:
        return res
    if only_mag:
        res = torch.sqrt(torch.sum(res * res, dim=-1, keepdim=True))
    return _split_ri(rho / rho_sc, (nb, ns, H, W)), res


# ----------------------------------------------------------------------------------------------
# magnitude-only fit
# ----------------------------------------------------------------------------------------------
def CSE_mag(acqs, out_maps, params, r2_sc=200.0, demod_signal=False, R2_prob=False, uncertainty=False,
            rdtype=torch.float32):
    """Magnitude-only three-parameter fit + eigen-decomposition (IDEAL_model.py:314-401).
    acqs (nb,ne,H,W,1) magnitudes; out_maps (nb,1,H,W,1) R2*/r2_sc (or an object with `.nu` when R2_prob)."""
    maps = _t(out_maps if not R2_prob else out_maps.tensor, rdtype)
    acqs = _t(acqs, rdtype)
    nb, _, H, W, _ = maps.shape
    te = _t(params[1], rdtype)
    ne = te.shape[1]
    M = gen_M(te, field=params[0], get_Mpinv=False, rdtype=rdtype)
    A, A_pinv = gen_A(M)
    S = acqs.reshape(nb, ne, -1)
    r2s = (maps[:, 0, :, :, 0] * r2_sc).reshape(nb, 1, -1)
    Wm = torch.exp(torch.matmul(te, r2s))
    Wp = torch.exp(torch.matmul(-te, r2s))
    y = (Wm * S) ** 2
    if R2_prob:
        r2s_nu = (_t(out_maps.nu, rdtype)[:, 0, :, :, 0] * r2_sc).reshape(nb, 1, -1)
        y_nu = (torch.exp(torch.matmul(te, r2s_nu)) * S) ** 2
    abc = torch.matmul(A_pinv, y)
    fit = torch.matmul(A, abc)
    S_hat = Wp * torch.where(fit > 1e-6, torch.sqrt(torch.where(fit > 1e-6, fit, torch.ones_like(fit))),
                             torch.zeros_like(fit))
    rho_abc = abc.transpose(1, 2)
    rho_hat, rho_unc = eigenvals(rho_abc)
    res_rho = rho_hat.transpose(1, 2).reshape(nb, ns, H, W, 1) / rho_sc
    res_demod = (y_nu if R2_prob else y).reshape(nb, ne, H, W, 1)
    res_ls = rho_abc.transpose(1, 2).reshape(nb, 3, H, W, 1) / (rho_sc ** 2)
    res_gt = S_hat.reshape(nb, ne, H, W, 1)
    res_unc = rho_unc.transpose(1, 2).reshape(nb, 1, H, W, 1)
    if uncertainty and demod_signal:
        return res_rho, res_gt, res_demod, res_unc
    if uncertainty:
        return res_rho, res_gt, res_unc, res_ls
    if demod_signal:
        return res_rho, res_gt, res_demod, res_ls
    return res_rho, res_gt


# ----------------------------------------------------------------------------------------------
# uncertainty propagation
# ----------------------------------------------------------------------------------------------
class Moments:
    """Duck-typed stand-in for the tfp distributions the reference passes (`.mean()`, `.variance()`)."""

    def __init__(self, mean, variance):
        self._m, self._v = mean, variance

    def mean(self):
        return self._m

    def variance(self):
        return self._v


def acq_uncertainty(rho_maps, phi_tfp, r2s_tfp, ne=6, te=None, r2_sc=200.0, field=1.5, rem_R2=False,
                    only_mag=False, rdtype=torch.float32):
    """Signal-domain variance Var_e = V_e |M rho|_e^2  (IDEAL_model.py:710-767)."""
    rho_maps = _t(rho_maps, rdtype)
    nb, _, H, W, _ = rho_maps.shape
    if te is None:
        te = _default_te(ne, nb, field)
    te = _t(te, rdtype)
    M = gen_M(te, field=field, get_Mpinv=False, rdtype=rdtype)
    rho = (torch.complex(rho_maps[:, :2, :, :, 0], rho_maps[:, :2, :, :, 1]) * rho_sc).reshape(nb, ns, -1)
    phi_var = _t(phi_tfp.variance(), rdtype) * (fm_sc ** 2)
    V = 1 - torch.exp(torch.matmul(-(2 * np.pi * te) ** 2, phi_var.reshape(nb, 1, -1)))
    if not rem_R2:
        r2_mean = _t(r2s_tfp.mean(), rdtype) * r2_sc
        r2_var = _t(r2s_tfp.variance(), rdtype) * (r2_sc ** 2)
        if r2_mean.shape[-1] > 1:
            r2_mean, r2_var = r2_mean[..., :1], r2_var[..., :1]
        V = V + torch.exp(torch.matmul(-te, r2_mean.reshape(nb, 1, -1))) * torch.matmul(te ** 2, r2_var.reshape(nb, 1, -1))
    Mr = torch.matmul(M, rho)
    var = (V * (Mr * Mr.conj()).abs()).reshape(nb, ne, H, W, 1)
    return var if only_mag else torch.cat([var, var], dim=-1)


def PDFF_uncertainty(acqs, phi_tfp, r2s_tfp, te=None, r2_sc=200.0, rem_R2=False, rdtype=torch.float32):
    """Per-voxel weighted LS with an echo-wise noise model (IDEAL_model.py:628-706); 1.5 T only."""
    cd = _cdtype(rdtype)
    acqs = _t(acqs, rdtype)
    nb, ne, H, W, _ = acqs.shape
    if te is None:
        te = gen_TEvar(ne, bs=nb, orig=True)
    te = _t(te, rdtype)
    M, P0, M_pinv = gen_M(te, get_P0=True, rdtype=rdtype)
    S = torch.complex(acqs[..., 0], acqs[..., 1]).reshape(nb, ne, -1)
    phi_mean = _t(phi_tfp.mean(), rdtype) * fm_sc
    phi_var = _t(phi_tfp.variance(), rdtype) * (fm_sc ** 2)
    if rem_R2:
        r2_mean, r2_var = torch.zeros_like(phi_mean), torch.zeros_like(phi_var)
    else:
        r2_mean = _t(r2s_tfp.mean(), rdtype) * r2_sc
        r2_var = _t(r2s_tfp.variance(), rdtype) * (r2_sc ** 2)
    xi = _xi(phi_mean.reshape(nb, -1), r2_mean.reshape(nb, -1), cd)
    Wm = torch.exp(_modulator(te, xi, -1.0, cd))
    Wp = torch.exp(_modulator(te, xi, +1.0, cd))
    V = 1 - torch.exp(torch.matmul(-(2 * np.pi * te) ** 2, phi_var.reshape(nb, 1, -1)))
    if not rem_R2:
        V = V + torch.exp(torch.matmul(te, r2_mean.reshape(nb, 1, -1))) * torch.matmul(te ** 2, r2_var.reshape(nb, 1, -1))
    g = Wp * torch.matmul(P0, Wm)                 # projector applied to the demodulator, literally (:681)
    sig = V * (g.conj() * g).abs() + V * (S.conj() * S).abs()
    w = _div_no_nan(torch.ones_like(sig), sig)    # (nb,ne,nv)
    Mh = M.transpose(1, 2).conj()                 # (nb,2,ne)
    wc = torch.complex(w, torch.zeros_like(w)).permute(2, 0, 1)                  # (nv,nb,ne)
    MtSM = torch.matmul(Mh, wc.unsqueeze(-1) * M)                               # (nv,nb,2,2)
    cov = torch.linalg.inv(MtSM)
    y = (Wm * S).permute(2, 0, 1)
    rhs = torch.matmul(Mh, (wc * y).unsqueeze(-1))                               # (nv,nb,2,1)
    rho = torch.matmul(cov, rhs).permute(1, 2, 0, 3).reshape(nb, ns, H, W, 1) / rho_sc
    res_rho = torch.cat([rho.real, rho.imag], dim=-1)
    res_var = cov.reshape(-1, nb, ns * ns).permute(1, 2, 0).abs().reshape(nb, ns * ns, H, W, 1) / (rho_sc ** 2)
    return res_rho, res_var


# ----------------------------------------------------------------------------------------------
# physics losses (consumers of the operators)
# ----------------------------------------------------------------------------------------------
def masked_mse(acqs, recon):
    """where(A != 0, S_hat, 0) per component, then the global mean of squares
    (train-IDEAL-unsup.py:218,236; train-IDEAL-single.py:155-157)."""
    recon = torch.where(acqs != 0, recon, torch.zeros_like(recon))
    return torch.mean((acqs - recon) ** 2)


def var_mse(y_true, y_pred_and_var):
    """VarMeanSquaredError (tf2gan/loss.py:130-140): squared error over sigma (not sigma^2) + log sigma."""
    idx = y_pred_and_var.shape[-1] // 2
    var = y_pred_and_var[..., idx:]
    pred = y_pred_and_var[..., :idx]
    var = torch.where(var >= 1e-5, var, torch.full_like(var, 1e-5))
    std = torch.sqrt(var)
    return torch.mean(_div_no_nan((y_true - pred) ** 2, std) + torch.log(std))


def var_mse_r2(y_true, y_pred_and_var):
    """VarMeanSquaredErrorR2 (tf2gan/loss.py:143-162): Rician negative log-likelihood."""
    if y_pred_and_var.shape[-1] > 1:
        idx = y_pred_and_var.shape[-1] // 2
        var = y_pred_and_var[..., idx:]
    else:
        idx = 1
        var = torch.ones_like(y_pred_and_var[..., :1]) * 1e-2
    pred = y_pred_and_var[..., :idx]
    var = torch.where(var >= 1e-5, var, torch.full_like(var, 1e-5))
    ll = torch.where(y_true > 1e-5, torch.log(torch.where(y_true > 1e-5, y_true, torch.ones_like(y_true))),
                     torch.zeros_like(y_true))
    ll = ll - torch.log(var)
    ll = ll - _div_no_nan(y_true ** 2 + pred ** 2, 2 * var)
    z = _div_no_nan(y_true * pred, var)
    i0e = torch.special.i0e(z)
    ll = ll + torch.where(i0e > 0, torch.log(torch.where(i0e > 0, i0e, torch.ones_like(i0e))), torch.zeros_like(i0e))
    ll = ll + z
    return torch.mean(-ll)


def physics_loss_a2a(acqs, param_maps, te=None, field=1.5, r2_sc=200.0, rdtype=torch.float32):
    """Config-2 training objective: acq_to_acq -> mask -> MSE (train-IDEAL-unsup.py:214-236).
    Returns (loss, rho_hat/rho_sc, S_hat unmasked)."""
    acqs = _t(acqs, rdtype)
    rho, recon = acq_to_acq(acqs, param_maps, te=te, field=field, r2_sc=r2_sc, rdtype=rdtype)
    return masked_mse(acqs, recon), rho, recon


def physics_loss_a2a_uq(acqs, param_maps, phi_var, r2_mean=None, r2_var=None, te=None, field=1.5, r2_sc=200.0,
                        rdtype=torch.float32):
    """The published AI-DEAL objective (train-IDEAL-unsup.py:214-231): acq_to_acq -> mask, signal variance from
    acq_uncertainty on the STOP-GRADIENT water/fat estimate, VarMeanSquaredError on [recon, var].  `phi_var`, `r2_mean`,
    `r2_var` are (nb,1,H,W,1) maps in network units (what `.variance()` / `.mean()` of the tfp outputs hold);
    r2_mean = r2_var = None is rem_R2=True.  Returns (loss, rho_hat/rho_sc, S_hat unmasked, var)."""
    acqs = _t(acqs, rdtype)
    rho, recon = acq_to_acq(acqs, param_maps, te=te, field=field, r2_sc=r2_sc, rdtype=rdtype)
    masked = torch.where(acqs != 0, recon, torch.zeros_like(recon))
    rem = r2_mean is None
    var = acq_uncertainty(rho.detach(), Moments(None, phi_var), None if rem else Moments(r2_mean, r2_var), ne=acqs.shape[1], te=te,
                          r2_sc=r2_sc, field=field, rem_R2=rem, rdtype=rdtype)
    return var_mse(acqs, torch.cat([masked, var], dim=-1)), rho, recon, var


def physics_loss_a2a_rician(acqs, param_maps, phi_var, r2_mean=None, r2_var=None, te=None, field=1.5, r2_sc=200.0,
                            rdtype=torch.float32):
    """The R2* stage of AI-DEAL (train-IDEAL-unsup.py:267-292): magnitudes of the resynthesised echoes, masked where the REAL
    channel of A is zero (:281), magnitude variance from acq_uncertainty(only_mag=True) on the stop-gradient estimate,
    Rician negative log-likelihood VarMeanSquaredErrorR2 (tf2gan/loss.py:143-162) against |A|.
    Returns (loss, rho_hat/rho_sc, |S_hat| unmasked, var)."""
    acqs = _t(acqs, rdtype)
    rho, mag = acq_to_acq(acqs, param_maps, te=te, field=field, r2_sc=r2_sc, only_mag=True, rdtype=rdtype)
    masked = torch.where(acqs[..., :1] != 0, mag, torch.zeros_like(mag))
    rem = r2_mean is None
    var = acq_uncertainty(rho.detach(), Moments(None, phi_var), None if rem else Moments(r2_mean, r2_var), ne=acqs.shape[1], te=te,
                          r2_sc=r2_sc, field=field, rem_R2=rem, only_mag=True, rdtype=rdtype)
    y = torch.sqrt(torch.sum(acqs * acqs, dim=-1, keepdim=True))
    return var_mse_r2(y, torch.cat([masked, var], dim=-1)), rho, mag, var


def physics_loss_fwd(acqs, out_maps, te, field=1.5, r2_sc=200.0, model="wfpm", rdtype=torch.float32):
    """Forward-model -> mask -> MSE objective (train-IDEAL-single.py:154-157 for model='magpha')."""
    fn = {"wfpm": IDEAL_model, "ffpd": IDEAL_mag, "magpha": IDEAL_mag_phase}[model]
    acqs = _t(acqs, rdtype)
    recon = fn(out_maps, [field, te], r2_sc=r2_sc, rdtype=rdtype)
    return masked_mse(acqs, recon), recon


# ----------------------------------------------------------------------------------------------
# PDFF / R2* extraction (ROI-analysis.py:301-306,344-354; gen_LDM_dataset.py:217-227)
# ----------------------------------------------------------------------------------------------
def pdff_extract(rho, mode="complex_sum"):
    """rho (nb,2,H,W,2) -> PDFF (nb,H,W).  'complex_sum': |F|/|W+F|; 'mag_sum': |F|/(|W|+|F|);
    'mag_disc': f >= w ? f/|W+F| : 1 - w/|W+F|.  NaN (0/0) -> 0."""
    w = torch.complex(rho[:, 0, ..., 0], rho[:, 0, ..., 1])
    f = torch.complex(rho[:, 1, ..., 0], rho[:, 1, ..., 1])
    wa, fa = w.abs(), f.abs()
    if mode == "mag_sum":
        out = fa / (wa + fa)
    else:
        wf = (w + f).abs()
        out = fa / wf if mode == "complex_sum" else torch.where(fa >= wa, fa / wf, 1 - wa / wf)
    return torch.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0)


# ----------------------------------------------------------------------------------------------
# script-level reductions: train-IDEAL-mag.py:288-289,308-316 and ROI-analysis.py:301-322
# ----------------------------------------------------------------------------------------------
def total_variation(x):
    """tf.image.total_variation summed over the batch: x (n, H, W, C) -> sum |x[1:] - x[:-1]| over rows and columns."""
    return (x[:, 1:] - x[:, :-1]).abs().sum() + (x[:, :, 1:] - x[:, :, :-1]).abs().sum()


def mag_regularisers(ls=None, demod=None, r2=None):
    """The five sums train-IDEAL-mag.py logs (:288-289 R2_TV, :308-314 Ad_TV, LS_NZ, WF_NZ, LS_cond), unweighted.
    ls (nb,3,H,W,1), demod (nb,ne,H,W,1), r2 (nb,1,H,W,1).  Differentiable through autograd.
    Quirks kept: `[..., ::2]` / `[..., :1]` / `[..., -1:]` index the LAST axis (length 1), so LS_NZ runs over all
    three coefficients and WF_NZ compares every element with itself (identically zero)."""
    ref = next(t for t in (ls, demod, r2) if t is not None)
    zero = torch.zeros((), dtype=ref.dtype)
    out = {"Ad_TV": zero, "LS_NZ": zero, "WF_NZ": zero, "LS_cond": zero, "R2_TV": zero}
    if demod is not None:
        out["Ad_TV"] = total_variation(demod.reshape(-1, *demod.shape[2:]))
    if r2 is not None:
        out["R2_TV"] = total_variation(r2.reshape(r2.shape[0], *r2.shape[-3:]))
    if ls is not None:
        out["LS_NZ"] = torch.where(ls < 0, ls * ls, torch.zeros_like(ls)).sum()
        q = ls[:, 1:2] ** 2 - 4.0 * ls[:, 0:1] * ls[:, 2:3]
        out["LS_cond"] = torch.where(q > 0, q * q, torch.zeros_like(q)).sum()
    return out


def roi_maps(maps, var=None, mode=None):
    """ROI-analysis.py:301-322.  maps (nb,3,H,W,2), var (nb,5,H,W,2) -> (nb,H,W,4|5): |W|, |F|, |W+F|, R2* [, PDFF variance].
    mode None: no variance; 'PDFF-var': the propagated variance; 'PDFF-var-Mag': the W-F row (model_sel == 'Mag')."""
    wf_abs = torch.sqrt((maps[:, :2] ** 2).sum(-1)).permute(0, 2, 3, 1)                    # (nb,H,W,2)
    sum_abs = torch.sqrt((maps[:, :2].sum(1, keepdim=True) ** 2).sum(-1)).permute(0, 2, 3, 1)
    out = torch.cat([wf_abs, sum_abs, maps[:, 2, :, :, 1:]], dim=-1)
    if mode is None:
        return out
    w_var = torch.complex(var[:, 0, ..., :1], var[:, 0, ..., 1:]).abs()
    wf_var = torch.complex(var[:, 1, ..., :1], var[:, 1, ..., 1:]).abs()
    f_var = torch.complex(var[:, 3, ..., :1], var[:, 3, ..., 1:]).abs()
    if mode == "PDFF-var-Mag":
        pv = wf_var
    else:
        wa = wf_abs[..., :1]
        pv = w_var / wa ** 2
        pv = pv - 2 * wf_var / (wa * sum_abs)
        pv = pv + (w_var + f_var + 2 * wf_var) / wa
        pv = pv * (wa ** 2 / sum_abs ** 2)
    return torch.cat([out, pv], dim=-1)


def ldm_images(out_maps, te=None, field=1.5, rdtype=torch.float32):
    """gen_LDM_dataset.py:156-158,216-218,225-227,234-237: the decoded mag/phase maps (nb,2,H,W,3|4) -> the forward signals and the
    three images the script writes per slice, each clip_by_value(., 0, 1) (NaN from 0/0 on background stays NaN):
    (S_hat (nb,ne,H,W,2), |S_hat| (nb,ne,H,W), PDFF = |F| / (|W| + |F|) on the raw magnitude channels (nb,H,W), R2* map (nb,H,W))."""
    sig = IDEAL_mag_Layer(field=field)(out_maps, te=te, rdtype=rdtype)
    m = out_maps.to(rdtype)
    pdff = torch.clamp(m[:, 0, :, :, 1] / (m[:, 0, :, :, 0] + m[:, 0, :, :, 1]), 0.0, 1.0)
    r2s = torch.clamp(m[:, 0, :, :, 2], 0.0, 1.0)
    mag = torch.clamp(torch.sqrt((sig ** 2).sum(-1)), 0.0, 1.0)
    return sig, mag, pdff, r2s


# ----------------------------------------------------------------------------------------------
# layout adapters (data.py:262-329)
# ----------------------------------------------------------------------------------------------
def A_from_MEBCRN(A):
    """(nb, ne, H, W, 2) -> (nb, H, W, 2 ne), Re/Im interleaved per echo (data.py:262-276)."""
    nb, ne, H, W, _ = A.shape
    return A.permute(0, 2, 3, 1, 4).reshape(nb, H, W, 2 * ne)


def A_to_MEBCRN(F):
    nb, H, W, c = F.shape
    return F.reshape(nb, H, W, c // 2, 2).permute(0, 3, 1, 2, 4)


def B_from_MEBCRN(B, mag_and_phase=False, c_pha=3):
    """data.py:279-299.  The mag/phase branch rotates BOTH species by c_pha * pi * B[:, 1, ..., 1]."""
    if mag_and_phase:
        ang = c_pha * B[:, 1, :, :, 1:2] * np.pi
        c, s = torch.cos(ang), torch.sin(ang)
        return torch.cat([B[:, 0, :, :, :1] * c, B[:, 0, :, :, :1] * s, B[:, 0, :, :, 1:2] * c, B[:, 0, :, :, 1:2] * s,
                          B[:, 0, :, :, 2:], B[:, 1, :, :, 2:]], dim=-1)
    return torch.cat([B[:, 0], B[:, 1], B[:, 2, :, :, 1:], B[:, 2, :, :, :1]], dim=-1)


def B_to_MEBCRN(B, mode="All"):
    """data.py:302-329."""
    z = torch.zeros_like(B[..., :1])
    if mode == "WF":
        return torch.stack([torch.cat([B[..., :1], z], -1), torch.cat([B[..., 1:], z], -1)], dim=1)
    if mode == "PM":
        return torch.cat([B[..., 1:], B[..., :1]], -1).unsqueeze(1)
    if mode == "WF-PM":
        return torch.stack([torch.cat([B[..., :1], z], -1), torch.cat([B[..., 1:2], z], -1), torch.cat([B[..., 3:], B[..., 2:3]], -1)], dim=1)
    if mode == "All":
        return torch.stack([B[..., :2], B[..., 2:4], torch.cat([B[..., 5:], B[..., 4:5]], -1)], dim=1)
    raise ValueError(mode)


def num_threads():
    return torch.get_num_threads()

