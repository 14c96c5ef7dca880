"""Framework-facing layer of the `wflib` drop-in: picks the tensor framework from the arguments, validates shapes the
way the reference's shape inference would, and routes to the differentiable torch operators (idealgan.torch_ops),
directly for torch tensors and through the DLPack bridge (idealgan.tf_ops) for TensorFlow tensors.
"""
import os

import numpy as np
import torch

from . import _lib as L
from . import tf_ops
from . import torch_ops as TO

MODELS = {"wfpm": L.MODEL_WFPM, "ffpd": L.MODEL_FFPD, "magpha": L.MODEL_MAGPHA}


def _framework():
    fw = os.environ.get("IDEALGAN_FRAMEWORK")
    if fw:
        return fw
    return "tf" if tf_ops.available() else "torch"


def keras_layer_base():
    """tf.keras.layers.Layer when the drop-in is used from TensorFlow, else None (plain callable classes)."""
    if _framework() == "tf" and tf_ops.available():
        return tf_ops.tf.keras.layers.Layer
    return None


def new_tensor(arr):
    """Host numpy array -> tensor of the session's framework (gen_TEvar's return type)."""
    if _framework() == "tf" and tf_ops.available():
        return tf_ops.tf.convert_to_tensor(arr)
    return torch.from_numpy(np.ascontiguousarray(arr))


def _is_tf(*xs):
    return any(tf_ops.is_tf_tensor(x) for x in xs)


def _dev_tensor(x, name):
    """torch CUDA float32 view of a torch / numpy input (numpy and CPU tensors are copied to the current device)."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32:
        return x                                            # the common case first: nothing to convert, nothing to probe
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        raise ValueError(f"{name}: expected a torch.Tensor, tf.Tensor or numpy array, got {type(x).__name__}")
    if x.dtype != torch.float32:
        raise ValueError(f"{name}: dtype must be float32, got {x.dtype}")
    if x.is_cuda:
        return x
    if not torch.cuda.is_available():
        raise L.IdealGanError("the IDEAL physics path runs on a CUDA device only (sm_100a kernels, no CPU fallback)")
    return x.cuda()


def _te_tensor(te):
    """Echo times as a torch tensor, wherever they live: device tensors stay on the device (no host round trip), numpy
    arrays and CPU tensors stay on the host (idealgan.torch_ops keys its table cache on their bytes)."""
    if isinstance(te, np.ndarray):
        te = torch.from_numpy(np.ascontiguousarray(te, dtype=np.float32))
    if not isinstance(te, torch.Tensor):
        raise ValueError(f"te: expected a torch.Tensor, tf.Tensor or numpy array, got {type(te).__name__}")
    if te.dim() not in (2, 3) or (te.dim() == 3 and int(te.shape[2]) != 1):
        raise ValueError(f"te must be (nb, ne, 1), got {tuple(te.shape)}")
    return te


def _dispatch(fn, tensors, names, te=None, out_shapes=None):
    """Run `fn(*tensors[, te])` (an idealgan.torch_ops function of torch tensors) on torch / numpy inputs, or lift it to
    TensorFlow when any argument -- the echo times included, which are symbolic inside @tf.function -- is a tf tensor.
    `out_shapes`: static output shapes, needed by the graph-mode hop (tf.py_function loses them)."""
    if te is not None and tf_ops.tf is not None and tf_ops.is_tf_tensor(te) and not _is_tf(*tensors) and hasattr(te, "numpy"):
        te = te.numpy()      # torch / numpy data with the session's default echo times (gen_TEvar made a tf constant): stay on the torch route
    consts = [] if te is None else [te]
    if tf_ops.tf is not None and _is_tf(*tensors, *consts):
        return tf_ops.bridge(fn, out_shapes, n_const=len(consts))(*tensors, *consts)
    all_numpy = all(isinstance(t, np.ndarray) for t in tensors)
    out = fn(*[_dev_tensor(t, n) for t, n in zip(tensors, names)], *[_te_tensor(c) for c in consts])
    if all_numpy:
        conv = lambda o: o.detach().cpu().numpy()          # noqa: E731
        return tuple(conv(o) for o in out) if isinstance(out, tuple) else conv(out)
    return out


def _te_host(te):
    """Host copy of the echo times (gen_M's stand-alone tables only; the operators never call this)."""
    if tf_ops.is_tf_tensor(te):
        if not hasattr(te, "numpy"):
            raise ValueError("gen_M / gen_A on a symbolic tensor: call them eagerly (their results are per-sample constants); "
                             "the operators take `te` itself and need no gen_M call")
        te = te.numpy()
    if isinstance(te, torch.Tensor):
        te = te.detach().cpu().numpy()
    te = np.ascontiguousarray(te, dtype=np.float32)
    if te.ndim == 3:
        te = te[:, :, 0]
    if te.ndim != 2:
        raise ValueError(f"te must be (nb, ne, 1), got {te.shape}")
    return np.ascontiguousarray(te)


def _ne_of(te):
    return int(te.shape[1])


# ---------------------------------------------------------------------------------------------------------------
# gen_M / gen_A / eigenvals
# ---------------------------------------------------------------------------------------------------------------
class Tables:
    """Host copy of the per-sample table (ig_gen_tables_host): source of gen_M's return values."""

    def __init__(self, te, field):
        self.like = te
        te_h = _te_host(te)
        self.nb, self.ne = te_h.shape
        tab = np.zeros((self.nb, L.TAB_FLOATS), np.float32)
        L.check(L.load().ig_gen_tables_host(te_h.ctypes.data, self.nb, self.ne, float(field), tab.ctypes.data), "ig_gen_tables_host")
        self.u = L.unpack_table(tab, self.ne)

    def _out(self, arr):
        arr = np.ascontiguousarray(arr.astype(np.complex64))
        if tf_ops.is_tf_tensor(self.like):       
            return tf_ops.tf.convert_to_tensor(arr)
        t = torch.from_numpy(arr)
        return t.to(self.like.device) if isinstance(self.like, torch.Tensor) else t

    def _c(self):
        return self.u["c"][0] + 1j * self.u["c"][1]

    def _pinv(self):
        return np.stack([self.u["pw"][0] + 1j * self.u["pw"][1], self.u["pf"][0] + 1j * self.u["pf"][1]], axis=1)

    def M(self):
        return self._out(np.stack([np.ones_like(self._c()), self._c()], axis=-1))

    def Mpinv(self):
        return self._out(self._pinv())

    def P0(self):
        M = np.stack([np.ones_like(self._c()), self._c()], axis=-1).astype(np.complex128)
        P = np.eye(self.ne)[None] - M @ self._pinv().astype(np.complex128)
        return self._out(0.5 * (P + np.conj(np.transpose(P, (0, 2, 1)))))

    def Hpinv(self):
        """pinv(Re(M^+ M)) (:64-68).  M^+ M is the 2x2 identity in exact arithmetic; what comes out is the identity plus the
        rounding of the stored fp32 table (~1e-7), like the reference's own fp32 value."""
        M = np.stack([np.ones_like(self._c()), self._c()], axis=-1).astype(np.complex128)
        H = np.real(self._pinv().astype(np.complex128) @ M)
        return self._out(np.linalg.pinv(H))


def tables(te, field):
    return Tables(te, field)


def gen_A(M, gen_AtA_pinv=False):
    """Small per-sample algebra on the (nb, ne, 2) model matrix; runs with the tensor's own framework ops."""
    if tf_ops.is_tf_tensor(M):                   
        Mt = torch.from_numpy(M.numpy())
    else:
        Mt = M if isinstance(M, torch.Tensor) else torch.from_numpy(np.asarray(M))
    A = torch.cat([Mt.abs()[..., :1], Mt.real[..., 1:], Mt.abs()[..., 1:] ** 2], dim=-1)
    A_pinv = torch.linalg.pinv(A.double()).to(A.dtype)
    outs = [A, A_pinv]
    if gen_AtA_pinv:
        outs.append(torch.linalg.pinv((A.transpose(1, 2) @ A).double()).to(A.dtype))
    if tf_ops.is_tf_tensor(M):                   
        outs = [tf_ops.tf.convert_to_tensor(o.cpu().numpy()) for o in outs]
    return tuple(outs)


# ---------------------------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------------------------
def ideal_forward(kind, out_maps, te, field, r2_sc):
    model = MODELS[kind]
    if len(out_maps.shape) != 5:
        raise ValueError(f"out_maps must have 5 dimensions, got shape {tuple(out_maps.shape)}")
    fn = lambda m, t: TO.ideal_forward(model, m, t, float(field), float(r2_sc))       # noqa: E731
    nb, _, H, W, _ = [int(s) for s in out_maps.shape]
    return _dispatch(fn, [out_maps], ["out_maps"], te, [(nb, _ne_of(te), H, W, 2)])


def acq_to_acq(acqs, param_maps, te, field, r2_sc, only_mag):
    if len(acqs.shape) != 5 or int(acqs.shape[-1]) != 2:
        raise ValueError("acq_to_acq: acqs must be (nb, ne, H, W, 2); magnitude-only input (n_ch == 1) is served by CSE_mag")
    fn = lambda a, p, t: TO.acq_to_acq(a, p, t, float(field), float(r2_sc), bool(only_mag))      # noqa: E731
    nb, ne, H, W, _ = [int(s) for s in acqs.shape]
    return _dispatch(fn, [acqs, param_maps], ["acqs", "param_maps"], te, [(nb, 2, H, W, 2), (nb, ne, H, W, 1 if only_mag else 2)])


def get_rho(acqs, param_maps, te, field, r2_sc, phase_constraint, MEBCRN, acq_demod):
    flags = (L.F_PHASE_CONSTRAINT if phase_constraint else 0) | (0 if MEBCRN else L.F_FLAT)
    if acq_demod and not MEBCRN:
        raise ValueError("get_rho: acq_demod is only defined for the MEBCRN layout")
    fn = lambda a, p, t: TO.get_rho(a, p, t, float(field), float(r2_sc), flags, bool(acq_demod))  # noqa: E731
    shp = [int(s) for s in acqs.shape]
    if MEBCRN:
        outs = [(shp[0], 2, shp[2], shp[3], 2)] + ([tuple(shp)] if acq_demod else [])
    else:
        outs = [(shp[0], shp[1], shp[2], 4)]
    return _dispatch(fn, [acqs, param_maps], ["acqs", "param_maps"], te, outs)


def eigenvals(X):
    lead = tuple(int(s) for s in X.shape[:-1])
    return _dispatch(TO.eigenvals, [X], ["X"], None, [lead + (2,), lead + (1,)])


def cse_mag(acqs, out_maps, te, field, r2_sc, R2_prob):
    nb, ne, H, W = [int(s) for s in acqs.shape[:4]]
    outs = [(nb, 2, H, W, 1), (nb, ne, H, W, 1), (nb, ne, H, W, 1), (nb, 3, H, W, 1), (nb, 1, H, W, 1)]
    if R2_prob:
        # the reference reads `.nu` of a tfp distribution and indexes the object itself as a tensor (:335-338)
        nu, r2 = out_maps.nu, (out_maps.tensor if hasattr(out_maps, "tensor") else out_maps.mean())
        fn = lambda a, r, n, t: TO.cse_mag(a, r, t, float(field), float(r2_sc), n)                 # noqa: E731
        return _dispatch(fn, [acqs, r2, nu], ["acqs", "out_maps", "out_maps.nu"], te, outs)
    fn = lambda a, r, t: TO.cse_mag(a, r, t, float(field), float(r2_sc))                           # noqa: E731
    return _dispatch(fn, [acqs, out_maps], ["acqs", "out_maps"], te, outs)


def acq_uncertainty(rho_maps, phi_tfp, r2s_tfp, te, r2_sc, field, rem_R2, only_mag):
    nb, _, H, W, _ = [int(s) for s in rho_maps.shape]
    outs = [(nb, _ne_of(te), H, W, 1 if only_mag else 2)]
    phi_var = phi_tfp.variance()
    if rem_R2:
        fn = lambda r, pv, t: TO.acq_uncertainty(r, pv, None, None, t, float(field), float(r2_sc), only_mag)   # noqa: E731
        return _dispatch(fn, [rho_maps, phi_var], ["rho_maps", "phi variance"], te, outs)
    r2_mean, r2_var = r2s_tfp.mean(), r2s_tfp.variance()
    if int(r2_mean.shape[-1]) > 1:                          # (:741-743)
        r2_mean, r2_var = r2_mean[..., :1], r2_var[..., :1]
    fn = lambda r, pv, rm, rv, t: TO.acq_uncertainty(r, pv, rm, rv, t, float(field), float(r2_sc), only_mag)   # noqa: E731
    return _dispatch(fn, [rho_maps, phi_var, r2_mean, r2_var], ["rho_maps", "phi variance", "R2* mean", "R2* variance"], te, outs)


def pdff_uncertainty(acqs, phi_tfp, r2s_tfp, te, r2_sc, rem_R2):
    nb, _, H, W, _ = [int(s) for s in acqs.shape]
    outs = [(nb, 2, H, W, 2), (nb, 4, H, W, 1)]
    phi_mean, phi_var = phi_tfp.mean(), phi_tfp.variance()
    if rem_R2:
        fn = lambda a, pm, pv, t: TO.pdff_uncertainty(a, pm, pv, None, None, t, float(r2_sc))       # noqa: E731
        return _dispatch(fn, [acqs, phi_mean, phi_var], ["acqs", "phi mean", "phi variance"], te, outs)
    fn = lambda a, pm, pv, rm, rv, t: TO.pdff_uncertainty(a, pm, pv, rm, rv, t, float(r2_sc))       # noqa: E731
    return _dispatch(fn, [acqs, phi_mean, phi_var, r2s_tfp.mean(), r2s_tfp.variance()],
                     ["acqs", "phi mean", "phi variance", "R2* mean", "R2* variance"], te, outs)
