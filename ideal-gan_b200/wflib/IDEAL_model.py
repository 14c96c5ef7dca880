"""Drop-in for the reference's `wflib/IDEAL_model.py`: same names, same call signatures, same tensor layouts,
with every operator executed by the sm_100a kernels of libidealgan.so (no TensorFlow op chains, no CPU
fallback).  Put `ideal-gan_b200/` on sys.path ahead of the reference and `import wflib as wf` keeps working:

    A2B_WF, A2B2A = wf.acq_to_acq(A, A2B_PM, field=args.field)        # train-IDEAL-unsup.py:216
    B2A = wf.IDEAL_Layer(field=1.5)(B, te=te, training=False)          # train-IDEAL-TEaug.py:192,217
    B2A2B_WF = wf.get_rho(B2A, B2A2B_PM, field=args.field, te=te)      # train-IDEAL-TEaug.py:304

Tensors may be torch CUDA tensors (tested path) or TensorFlow GPU tensors (exchanged zero-copy through DLPack,
gradients registered with tf.custom_gradient, see idealgan/tf_ops.py).  Results have the type of the inputs.
Line numbers below refer to /root/reference/wflib/IDEAL_model.py.

Deliberate supersets of the library, needed by the reference's own callers:
  * acq_to_acq returns the 2-tuple (rho_hat / rho_sc, S_hat) its callers unpack and accepts only_mag=True;
    `legacy_single=True` restores the library's single S_hat tensor.
  * IDEAL_mag_Layer accepts the 2-row x 3|4-channel mag/phase tensor as well (gen_LDM_dataset.py:156-158).
  * IDEAL_mag_phase treats a missing bipolar channel as zero; get_rho(MEBCRN=False) has no bipolar term.
"""
import numpy as np

from idealgan import frontend as _fe

# Multipeak fat model (:4-19)
species = ["water", "fat"]
ns = len(species)
f_p = (np.array([0., -3.80, -3.40, -2.60, -1.94, -0.39, 0.60]) * 1E-6 * 42.58E6).astype(np.complex64)[None, :]
A_p = np.array([[1.0, 0.0], [0.0, 0.087], [0.0, 0.693], [0.0, 0.128], [0.0, 0.004], [0.0, 0.039], [0.0, 0.048]]).astype(np.complex64)
fm_sc = 300.0
rho_sc = 1.4

__all__ = ["species", "ns", "f_p", "A_p", "fm_sc", "rho_sc", "gen_TEvar", "gen_M", "gen_A", "eigenvals", "acq_to_acq",
           "CSE_to_CSE_Layer", "IDEAL_model", "IDEAL_Layer", "CSE_mag", "IDEAL_mag", "IDEAL_mag_phase", "IDEAL_mag_Layer",
           "get_rho", "PDFF_uncertainty", "acq_uncertainty", "np"]


def gen_TEvar(n_ech, bs=1, orig=False, TE_ini_min=1.0e-3, TE_ini_d=1.4e-3, d_TE_min=1.6e-3, d_TE_d=1.0e-3):
    """Echo-time vector, one train per call tiled over the batch -> (bs, n_ech, 1) float32 (:21-45).
    Host-side numpy with the global np.random state, exactly like the reference."""
    if orig:
        TE_ini_var = 1.3 * 1e-3
        d_TE_var = 2.1 * 1e-3
        te = np.arange(start=TE_ini_var, stop=TE_ini_var + d_TE_var * (n_ech - 1) + 1e-4, step=d_TE_var)
    elif not TE_ini_d and not d_TE_d:
        te = np.arange(start=TE_ini_min, stop=TE_ini_min + d_TE_min * (n_ech - 1) + 1e-4, step=d_TE_min)
    else:
        TE_ini_var = TE_ini_min + np.random.uniform(0, TE_ini_d)
        d_TE_c = d_TE_min + np.random.uniform(0, d_TE_d)
        d_TE_var = np.random.normal(d_TE_c, 1e-4, size=(n_ech - 1,))
        te = np.cumsum(np.concatenate((np.array([0.0]), d_TE_var), axis=0)) + TE_ini_var
    te = np.tile(te.astype(np.float32)[None, :, None], (bs, 1, 1))
    return _fe.new_tensor(te)


def gen_M(te, field=1.5, get_Mpinv=True, get_P0=False, get_H=False):
    """Model matrix M (bs, ne, 2), pseudo-inverse (bs, 2, ne), projector complement P0 (bs, ne, ne), H^+ (bs, 2, 2),
    with the reference's return-arity rules, including None for the combinations it does not cover (:48-77)."""
    t = _fe.tables(te, field)
    M = t.M()
    if get_P0 and get_Mpinv:
        return M, t.P0(), t.Mpinv()
    elif get_Mpinv and not get_P0 and not get_H:
        return M, t.Mpinv()
    elif get_Mpinv and not get_P0:
        return M, t.Mpinv(), t.Hpinv()
    elif not get_Mpinv and not get_P0 and not get_H:
        return M
    return None


def gen_A(M, gen_AtA_pinv=False):
    """Magnitude design matrix A = [|M_w|, Re M_f, |M_f|^2] and its pseudo-inverse (:80-97)."""
    return _fe.gen_A(M, gen_AtA_pinv)


def eigenvals(X):
    """Closed-form principal eigenpair of [[a, b/2], [b/2, c]] per voxel: (sqrt(l_max) v_max, l_min / l_max) (:100-138)."""
    return _fe.eigenvals(X)


def acq_to_acq(acqs, param_maps, te=None, field=1.5, r2_sc=200.0, only_mag=False, legacy_single=False):
    """Project the measured echoes on the water/fat model subspace and resynthesise them (:142-200):
    S_hat = Wp M M^+ Wm S.  Returns (rho_hat / rho_sc (nb,2,H,W,2), S_hat (nb,ne,H,W,2 | 1 with only_mag))."""
    n_batch, ne = acqs.shape[0], acqs.shape[1]
    if te is None:
        if field == 1.5:
            te = gen_TEvar(ne, bs=n_batch, orig=True)
        elif field == 3.0:
            te = gen_TEvar(ne, bs=n_batch, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None)
        else:
            raise ValueError("acq_to_acq: default echo times exist only for field 1.5 and 3.0 (:145-149); pass te")
    rho, s_hat = _fe.acq_to_acq(acqs, param_maps, te, field, r2_sc, only_mag)
    return s_hat if legacy_single else (rho, s_hat)


class _Layer:
    """Stand-in for tf.keras.layers.Layer: the reference's layers hold no weights, `__call__` forwards to `call`."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


_LayerBase = _fe.keras_layer_base() or _Layer


class CSE_to_CSE_Layer(_LayerBase):
    """(:203-216) Keras-style wrapper of acq_to_acq returning S_hat; a third input is replaced by the `orig` TEs."""

    def __init__(self, mag_only=False, field=1.5, r2_sc=200.0):
        super(CSE_to_CSE_Layer, self).__init__()
        self.field = field
        self.r2_sc = r2_sc

    def call(self, inputs, training=None):
        if len(inputs) == 3:
            acqs, out_maps, te = inputs
            te = gen_TEvar(acqs.shape[1], out_maps.shape[0], orig=True)
        else:
            acqs, out_maps = inputs
            te = None
        return acq_to_acq(acqs, out_maps, te=te, field=self.field, r2_sc=self.r2_sc, legacy_single=True)


def IDEAL_model(out_maps, params, r2_sc=200.0):
    """Forward model, WF-PM maps (nb, 3|4, H, W, 2) -> (nb, ne, H, W, 2); params = [field, te] (:220-299)."""
    return _fe.ideal_forward("wfpm", out_maps, params[1], params[0], r2_sc)


class IDEAL_Layer(_LayerBase):
    """(:302-311)"""

    def __init__(self, field=1.5, r2_sc=200.0):
        super(IDEAL_Layer, self).__init__()
        self.field = field
        self.r2_sc = r2_sc

    def call(self, out_maps, te=None, ne=6, training=None):
        if te is None:
            te = gen_TEvar(ne, out_maps.shape[0], orig=True)
        return IDEAL_model(out_maps, [self.field, te], r2_sc=self.r2_sc)


def CSE_mag(acqs, out_maps, params, r2_sc=200.0, demod_signal=False, R2_prob=False, uncertainty=False):
    """Magnitude-only fit: y = (e^{te R2*} |S|)^2, abc = A^+ y, eigen-decomposition -> |W|, |F| (:314-401).
    Return arity follows the reference: 2 results, or 4 depending on demod_signal / uncertainty."""
    res_rho, res_gt, res_demod, res_ls, res_unc = _fe.cse_mag(acqs, out_maps, params[1], params[0], r2_sc, R2_prob)
    if uncertainty and demod_signal:
        return (res_rho, res_gt, res_demod, res_unc)
    elif uncertainty:
        return (res_rho, res_gt, res_unc, res_ls)
    elif demod_signal:
        return (res_rho, res_gt, res_demod, res_ls)
    return (res_rho, res_gt)


def IDEAL_mag(out_maps, params, r2_sc=200.0):
    """Forward model, PDFF / PD / common-phase maps (nb, 3, H, W, 2) (:404-453)."""
    return _fe.ideal_forward("ffpd", out_maps, params[1], params[0], r2_sc)


def IDEAL_mag_phase(out_maps, params, r2_sc=200.0):
    """Forward model, per-species magnitude / phase maps (nb, 2, H, W, 3|4) with the bipolar term x 4 pi (:456-509)."""
    return _fe.ideal_forward("magpha", out_maps, params[1], params[0], r2_sc)


class IDEAL_mag_Layer(_LayerBase):
    """(:512-524) -- r2_sc is always 200 here, as in the reference."""

    def __init__(self, field=1.5, sep_phase=False):
        super(IDEAL_mag_Layer, self).__init__()
        self.field = field
        self.sep_phase = sep_phase

    def call(self, out_maps, te=None, ne=6, training=None):
        if te is None:
            te = gen_TEvar(ne, out_maps.shape[0], orig=True)
        if self.sep_phase or out_maps.shape[1] == 2:
            return IDEAL_mag_phase(out_maps, [self.field, te])
        return IDEAL_mag(out_maps, [self.field, te])


def get_rho(acqs, param_maps, field=1.5, te=None, r2_sc=200.0, phase_constraint=False, MEBCRN=True, acq_demod=False):
    """Least-squares water/fat solve rho_hat = M^+ (Wm S) / rho_sc (:527-624).  MEBCRN: acqs (nb,ne,H,W,2), param_maps
    (nb, rows>=1, H, W, 2) with (phi, R2*) in row 0 and, for rows > 3, the bipolar phase in the last row; flat:
    acqs (nb,H,W,2ne), param_maps (nb,H,W,2) ordered (R2*, phi), result (nb,H,W,4)."""
    if MEBCRN:
        n_batch, ne = acqs.shape[0], acqs.shape[1]
    else:
        n_batch, ne = acqs.shape[0], acqs.shape[3] // 2
    if te is None:
        te = gen_TEvar(ne, bs=n_batch, orig=True)
    return _fe.get_rho(acqs, param_maps, te, field, r2_sc, phase_constraint, MEBCRN, acq_demod)


def PDFF_uncertainty(acqs, phi_tfp, r2s_tfp, te=None, r2_sc=200.0, rem_R2=False):
    """Per-voxel weighted LS with an echo-wise noise model -> (rho_hat / rho_sc, |cov| / rho_sc^2) (:628-706); 1.5 T."""
    n_batch, ne = acqs.shape[0], acqs.shape[1]
    if te is None:
        te = gen_TEvar(ne, bs=n_batch, orig=True)
    return _fe.pdff_uncertainty(acqs, phi_tfp, r2s_tfp, te, r2_sc, rem_R2)


def acq_uncertainty(rho_maps, phi_tfp, r2s_tfp, ne=6, te=None, r2_sc=200.0, field=1.5, rem_R2=False, only_mag=False):
    """Signal-domain variance Var_e = V_e |M rho|_e^2 -> (nb, ne, H, W, 2 | 1) (:710-767)."""
    n_batch = rho_maps.shape[0]
    if te is None:
        if field == 1.5:
            te = gen_TEvar(ne, bs=n_batch, orig=True)
        elif field == 3.0:
            te = gen_TEvar(ne, bs=n_batch, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None)
        else:
            raise ValueError("acq_uncertainty: default echo times exist only for field 1.5 and 3.0 (:713-717); pass te")
    return _fe.acq_uncertainty(rho_maps, phi_tfp, r2s_tfp, te, r2_sc, field, rem_R2, only_mag)
