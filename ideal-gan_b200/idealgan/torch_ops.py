"""torch.autograd bindings over the C ABI (idealgan.ops): the differentiable operators the `wflib` drop-in is
built from.  TF's autodiff through ~25 ops per operator in the reference becomes one forward
and one adjoint kernel launch here.  The TensorFlow binding (tf_ops.py) wraps these same functions.
"""
import torch

from . import _lib as L
from . import ops


class _TableCache:
    """Per-sample tables (gen_M / gen_A constants, ig_gen_tables) keyed on the echo times, so that the operators of one
    training step -- acq_to_acq, get_rho, acq_uncertainty on the same `te` -- build the table once and no call moves `te`
    to the host.  Host-resident echo times (numpy / CPU tensors, what gen_TEvar returns) are keyed on their bytes; device
    tensors on (address, shape, strides, torch's in-place version counter -- shared by every alias of the storage), with a
    strong reference held while cached so the address cannot be handed to another allocation under the entry."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self.entries = {}            # key -> (table, te reference or None)
        self.hits = self.misses = 0

    def get(self, te, field, device):
        if torch.cuda.is_current_stream_capturing():
            # Inside a CUDA-graph capture the table is part of the captured step: it is rebuilt by every replay from whatever the echo-time
            # buffer holds then.  A cached table would be baked into the graph (stale after the buffer is refilled), and one built here
            # lives in the graph's private pool, so nothing is looked up and nothing is stored.
            return ops.gen_tables(te.to(device).contiguous(), field)
        stream = torch._C._cuda_getCurrentRawStream(device.index)     # a table is ordered on the stream that built it
        if te.is_cuda:
            key = ("dev", te.data_ptr(), te._version, tuple(te.shape), te.stride(), float(field), device.index, stream)
            hold = te
        else:
            key = ("host", te.numpy().tobytes(), tuple(te.shape), float(field), device.index, stream)
            hold = None
        hit = self.entries.get(key)
        if hit is not None:
            self.hits += 1
            return hit[0]
        self.misses += 1
        tab = ops.gen_tables(te.to(device).contiguous(), field)
        if len(self.entries) >= self.capacity:
            self.entries.pop(next(iter(self.entries)))
        self.entries[key] = (tab, hold)
        return tab

    def clear(self):
        self.entries.clear()


table_cache = _TableCache()


def _tables(te, field, device):
    """te: (nb, ne[, 1]) tensor/array on any device -> (table on `device`, ne).  No gradient: echo times are data.
    Device-resident echo times never visit the host (one 5 us ig_gen_tables launch on a cache miss)."""
    if not isinstance(te, torch.Tensor):
        te = torch.as_tensor(te, dtype=torch.float32)
    te = te.detach()
    if te.dtype != torch.float32:
        te = te.to(torch.float32)
    if te.dim() not in (2, 3) or (te.dim() == 3 and te.shape[2] != 1):
        raise ValueError(f"te must be (nb, ne, 1) or (nb, ne), got {tuple(te.shape)}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return table_cache.get(te, field, device), te.shape[1]


def _no_grad(*tensors):
    """True when nothing asks for a gradient: the operator then skips torch.autograd.Function (15-20 us of host time per
    call, more than a single-slice kernel runs)."""
    return not torch.is_grad_enabled() or not any(t is not None and t.requires_grad for t in tensors)


def _check_batch(te_nb, nb):
    if te_nb != nb:
        raise ValueError(f"te has {te_nb} rows for a batch of {nb} (one echo train per sample)")


class _IdealForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, tab, model, ne, r2_sc, flags):
        maps = maps.contiguous()
        ctx.save_for_backward(maps, tab)
        ctx.cfg = (model, ne, r2_sc, flags)
        return ops.ideal_fwd(model, maps, tab, ne, r2_sc, flags)

    @staticmethod
    def backward(ctx, gout):
        maps, tab = ctx.saved_tensors
        model, ne, r2_sc, flags = ctx.cfg
        gout = gout.contiguous()
        if flags & L.F_FLAT:                     # interleaved upstream -> planar (the adapter's adjoint), then the usual adjoint kernel
            from . import layout
            gout, flags = layout._acq_from_flat(gout), flags & ~L.F_FLAT
        return ops.ideal_bwd(model, maps, tab, ne, gout, r2_sc, flags), None, None, None, None, None


def ideal_forward(model, maps, te, field=1.5, r2_sc=200.0, flags=0):
    """Forward signal model (IDEAL_model / IDEAL_mag / IDEAL_mag_phase) with its adjoint registered.  flags & F_FLAT: the
    result is channel-interleaved, (nb, H, W, 2 ne) = data.A_from_MEBCRN(IDEAL_op(maps)) in one kernel (train-sup.py:242-244)."""
    tab, ne = _tables(te, field, maps.device)
    _check_batch(tab.shape[0], maps.shape[0])
    if _no_grad(maps):
        return ops.ideal_fwd(model, maps, tab, ne, float(r2_sc), int(flags))
    if flags & L.F_ONLY_MAG:
        raise ValueError("ideal_forward: the magnitude-only output (F_ONLY_MAG) is an inference option; no adjoint kernel takes a magnitude upstream")
    return _IdealForward.apply(maps, tab, model, ne, float(r2_sc), int(flags))


def ldm_decode(maps, te, field=1.5, r2_sc=200.0, model=None, want_shat=True, clip=True):
    """Inference-only physics decoding of gen_LDM_dataset.py:156-158,216-237 in one kernel: decoded maps -> (S_hat | None,
    |S_hat| (nb,ne,H,W), PDFF (nb,H,W), R2* map (nb,H,W)), the three images clipped to [0, 1].  `maps`: the script's
    (nb,2,H,W,3|4) mag/phase tensor by default; model = L.MODEL_WFPM / L.MODEL_FFPD for the complex-row parameterisations."""
    tab, ne = _tables(te, field, maps.device)
    _check_batch(tab.shape[0], maps.shape[0])
    with torch.no_grad():
        return ops.ideal_decode(L.MODEL_MAGPHA if model is None else model, maps.detach(), tab, ne, r2_sc, want_shat=want_shat, clip=clip)


class _GetRho(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acqs, pm, tab, r2_sc, flags, want_demod):
        acqs, pm = acqs.contiguous(), pm.contiguous()
        ctx.save_for_backward(acqs, pm, tab)
        ctx.cfg = (r2_sc, flags)
        rho, demod = ops.get_rho_fwd(acqs, pm, tab, r2_sc, flags, want_demod)
        if demod is None:
            demod = acqs.new_empty(0)
            ctx.mark_non_differentiable(demod)
        return rho, demod

    @staticmethod
    def backward(ctx, g_rho, g_demod):
        acqs, pm, tab = ctx.saved_tensors
        r2_sc, flags = ctx.cfg
        g_demod = None if g_demod is None or g_demod.numel() == 0 else g_demod.contiguous()
        g_rho = None if g_rho is None else g_rho.contiguous()
        g_acqs, g_pm = ops.get_rho_bwd(acqs, pm, tab, g_rho, g_demod, r2_sc, flags, need_acqs=ctx.needs_input_grad[0])
        return g_acqs, g_pm, None, None, None, None


def get_rho(acqs, pm, te, field=1.5, r2_sc=200.0, flags=0, want_demod=False):
    flat = bool(flags & L.F_FLAT)
    tab, ne = _tables(te, field, acqs.device)
    _check_batch(tab.shape[0], acqs.shape[0])
    if ne != (acqs.shape[-1] // 2 if flat else acqs.shape[1]):
        raise ValueError(f"te has {ne} echoes, acquisitions have {acqs.shape[-1] // 2 if flat else acqs.shape[1]}")
    if _no_grad(acqs, pm):
        rho, demod = ops.get_rho_fwd(acqs, pm, tab, float(r2_sc), int(flags), bool(want_demod))
    else:
        rho, demod = _GetRho.apply(acqs, pm, tab, float(r2_sc), int(flags), bool(want_demod))
    return (rho, demod) if want_demod else rho


class _AcqToAcq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acqs, pm, tab, r2_sc, flags):
        acqs, pm = acqs.contiguous(), pm.contiguous()
        ctx.save_for_backward(acqs, pm, tab)
        ctx.cfg = (r2_sc, flags)
        return ops.a2a_fwd(acqs, pm, tab, r2_sc, flags, want_rho=True)

    @staticmethod
    def backward(ctx, g_rho, g_shat):
        acqs, pm, tab = ctx.saved_tensors
        r2_sc, flags = ctx.cfg
        g_rho = None if g_rho is None else g_rho.contiguous()
        g_shat = None if g_shat is None else g_shat.contiguous()
        g_acqs, g_pm = ops.a2a_bwd(acqs, pm, tab, g_rho, g_shat, r2_sc, flags, need_acqs=ctx.needs_input_grad[0])
        return g_acqs, g_pm, None, None, None


def acq_to_acq(acqs, pm, te, field=1.5, r2_sc=200.0, only_mag=False):
    """(rho_hat / rho_sc, S_hat) -- or |S_hat| with only_mag -- differentiable in acqs and pm."""
    tab, ne = _tables(te, field, acqs.device)
    _check_batch(tab.shape[0], acqs.shape[0])
    if ne != acqs.shape[1]:
        raise ValueError(f"te has {ne} echoes, acquisitions have {acqs.shape[1]}")
    if _no_grad(acqs, pm):
        return ops.a2a_fwd(acqs, pm, tab, float(r2_sc), L.F_ONLY_MAG if only_mag else 0, want_rho=True)
    return _AcqToAcq.apply(acqs, pm, tab, float(r2_sc), L.F_ONLY_MAG if only_mag else 0)


class _A2ALoss(torch.autograd.Function):
    """Fused config-2 objective: the forward pass already produces d loss / d pm, backward only scales it."""

    @staticmethod
    def forward(ctx, acqs, pm, tab, r2_sc, inv_n):
        loss, g_pm, _, _ = ops.a2a_loss(acqs.contiguous(), pm.contiguous(), tab, r2_sc, inv_n)
        if pm.shape[1] != 1:
            full = torch.zeros_like(pm)
            full[:, :1] = g_pm
            g_pm = full
        ctx.save_for_backward(g_pm)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (g_pm,) = ctx.saved_tensors
        return None, g_pm * g, None, None, None


def physics_loss_a2a(acqs, pm, te, field=1.5, r2_sc=200.0, inv_n=None):
    """mean((A - where(A != 0, acq_to_acq(A, pm), 0))^2) as one kernel (train-IDEAL-unsup.py:214-218,236,255).
    Differentiable in pm only (A is data).  inv_n: 1 / (elements of the global batch) for sharded batches."""
    tab, ne = _tables(te, field, acqs.device)
    _check_batch(tab.shape[0], acqs.shape[0])
    inv_n = 1.0 / acqs.numel() if inv_n is None else float(inv_n)
    return _A2ALoss.apply(acqs, pm, tab, float(r2_sc), inv_n)


class _A2AUqLoss(torch.autograd.Function):
    """Fused uncertainty-aware objective: forward produces every gradient, backward scales them."""

    @staticmethod
    def forward(ctx, acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc, inv_n, rician=False):
        fused = ops.a2a_rician_loss if rician else ops.a2a_uq_loss
        loss, g_pm, g_pv, g_rm, g_rv, rho = fused(acqs.contiguous(), pm.contiguous(), phi_var.contiguous(),
                                                  None if r2_mean is None else r2_mean.contiguous(),
                                                  None if r2_var is None else r2_var.contiguous(), tab, r2_sc, inv_n, want_rho=True)
        if pm.shape[1] != 1:
            full = torch.zeros_like(pm)
            full[:, :1] = g_pm
            g_pm = full
        ctx.has_r2 = r2_mean is not None
        ctx.save_for_backward(g_pm, g_pv, *([g_rm, g_rv] if ctx.has_r2 else []))
        ctx.mark_non_differentiable(rho)
        return loss.reshape(()), rho

    @staticmethod
    def backward(ctx, g, _g_rho):
        saved = ctx.saved_tensors
        g_rm, g_rv = (saved[2] * g, saved[3] * g) if ctx.has_r2 else (None, None)
        return None, saved[0] * g, saved[1] * g, g_rm, g_rv, None, None, None, None


def physics_loss_a2a_rician(acqs, pm, phi_var, r2_mean, r2_var, te, field=1.5, r2_sc=200.0, inv_n=None):
    """The R2* stage objective of AI-DEAL as one kernel (train-IDEAL-unsup.py:267-292): acq_to_acq(only_mag=True) -> mask on the
    real channel -> acq_uncertainty(only_mag=True) on the stop-gradient estimate -> VarMeanSquaredErrorR2 (Rician negative
    log-likelihood) against |A|.  Same arguments and results as physics_loss_a2a_uq; inv_n: 1 / (nb ne H W of the global batch)."""
    tab, ne = _tables(te, field, acqs.device)
    _check_batch(tab.shape[0], acqs.shape[0])
    if ne != acqs.shape[1]:
        raise ValueError(f"te has {ne} echoes, acquisitions have {acqs.shape[1]}")
    inv_n = 2.0 / acqs.numel() if inv_n is None else float(inv_n)
    return _A2AUqLoss.apply(acqs, pm, phi_var, r2_mean, r2_var, tab, float(r2_sc), inv_n, True)


def physics_loss_a2a_uq(acqs, pm, phi_var, r2_mean, r2_var, te, field=1.5, r2_sc=200.0, inv_n=None):
    """The AI-DEAL training objective as one kernel (train-IDEAL-unsup.py:214-231): acq_to_acq -> mask ->
    acq_uncertainty(stop_gradient(A2B_WF), FM, R2) -> VarMeanSquaredError.  `pm` is the (phi, R2*) sample fed to acq_to_acq,
    `phi_var` / `r2_mean` / `r2_var` the `.variance()` / `.mean()` maps of the network's output distributions
    ((nb,1,H,W,1); r2_mean = r2_var = None is rem_R2=True).  Differentiable in pm and the moment maps.
    Returns (loss, A2B_WF) with A2B_WF = rho_hat / rho_sc detached, as the training loop logs it."""
    tab, ne = _tables(te, field, acqs.device)
    _check_batch(tab.shape[0], acqs.shape[0])
    if ne != acqs.shape[1]:
        raise ValueError(f"te has {ne} echoes, acquisitions have {acqs.shape[1]}")
    inv_n = 1.0 / acqs.numel() if inv_n is None else float(inv_n)
    return _A2AUqLoss.apply(acqs, pm, phi_var, r2_mean, r2_var, tab, float(r2_sc), inv_n)


class _IdealLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, maps, acqs, tab, model, r2_sc, flags, inv_n):
        loss, gmaps, _ = ops.ideal_loss(model, maps.contiguous(), acqs.contiguous(), tab, r2_sc, flags, inv_n)
        ctx.save_for_backward(gmaps)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gmaps,) = ctx.saved_tensors
        return gmaps * g, None, None, None, None, None, None


def physics_loss_forward(model, maps, acqs, te, field=1.5, r2_sc=200.0, flags=0, inv_n=None):
    """mean((A - where(A != 0, forward_model(maps), 0))^2) as one kernel (train-IDEAL-single.py:154-157,175)."""
    tab, ne = _tables(te, field, maps.device)
    _check_batch(tab.shape[0], maps.shape[0])
    if ne != acqs.shape[1]:
        raise ValueError(f"te has {ne} echoes, acquisitions have {acqs.shape[1]}")
    inv_n = 1.0 / acqs.numel() if inv_n is None else float(inv_n)
    return _IdealLoss.apply(maps, acqs, tab, model, float(r2_sc), int(flags), inv_n)


# ---------------------------------------------------------------------------------------------------------------
# second tier
# ---------------------------------------------------------------------------------------------------------------
class _Eigenvals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = X.contiguous()
        ctx.save_for_backward(X)
        return ops.eigenvals_fwd(X)

    @staticmethod
    def backward(ctx, g_xy, g_ratio):
        (X,) = ctx.saved_tensors
        return ops.eigenvals_bwd(X, g_xy, g_ratio)


def eigenvals(X):
    return _Eigenvals.apply(X)


class _CseMag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag, r2, r2nu, tab, r2_sc):
        mag, r2 = mag.contiguous(), r2.contiguous()
        has_nu = r2nu is not None
        ctx.save_for_backward(mag, r2, tab, r2nu.contiguous() if has_nu else tab)
        ctx.cfg = (r2_sc, has_nu)
        return ops.cse_mag_fwd(mag, r2, tab, r2_sc, r2nu)

    @staticmethod
    def backward(ctx, *grads):
        mag, r2, tab, nu = ctx.saved_tensors
        r2_sc, has_nu = ctx.cfg
        g_mag, g_r2, g_nu = ops.cse_mag_bwd(mag, r2, tab, grads, r2_sc, nu if has_nu else None)
        return g_mag, g_r2, g_nu, None, None


def cse_mag(mag, r2, te, field=1.5, r2_sc=200.0, r2nu=None):
    """(rho/rho_sc, S_hat, demodulated y, abc/rho_sc^2, rank-1 ratio), differentiable in mag, r2 (and r2nu)."""
    tab, ne = _tables(te, field, mag.device)
    _check_batch(tab.shape[0], mag.shape[0])
    if ne != mag.shape[1]:
        raise ValueError(f"te has {ne} echoes, acquisitions have {mag.shape[1]}")
    return _CseMag.apply(mag, r2, r2nu, tab, float(r2_sc))


class _AcqUnc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rho, phi_var, r2_mean, r2_var, tab, ne, r2_sc, only_mag):
        rem = r2_mean is None
        ctx.save_for_backward(rho, phi_var, tab, *( () if rem else (r2_mean, r2_var)))
        ctx.cfg = (ne, r2_sc, only_mag, rem)
        return ops.acq_unc_fwd(rho, phi_var, r2_mean, r2_var, tab, ne, r2_sc, only_mag)

    @staticmethod
    def backward(ctx, g):
        ne, r2_sc, only_mag, rem = ctx.cfg
        saved = ctx.saved_tensors
        rho, phi_var, tab = saved[:3]
        r2_mean, r2_var = (None, None) if rem else saved[3:]
        g_pv, g_rm, g_rv = ops.acq_unc_bwd(rho, phi_var, r2_mean, r2_var, tab, ne, g.contiguous(), r2_sc, only_mag)
        return None, g_pv, g_rm, g_rv, None, None, None, None


def acq_uncertainty(rho, phi_var, r2_mean, r2_var, te, field=1.5, r2_sc=200.0, only_mag=False):
    """Var_e = V_e |M rho|^2; differentiable in the three moment maps (rho is a constant, as in its callers:
    train-IDEAL-unsup.py:222 passes tf.stop_gradient(A2B_WF))."""
    tab, ne = _tables(te, field, rho.device)
    _check_batch(tab.shape[0], rho.shape[0])
    return _AcqUnc.apply(rho.detach(), phi_var, r2_mean, r2_var, tab, ne, float(r2_sc), bool(only_mag))


def pdff_uncertainty(acqs, phi_mean, phi_var, r2_mean, r2_var, te, r2_sc=200.0):
    """Inference-only weighted LS (no adjoint: its only callers are evaluation scripts, ROI-analysis.py:240-244)."""
    tab, ne = _tables(te, 1.5, acqs.device)          # the reference fixes 1.5 T here (IDEAL_model.py:634)
    _check_batch(tab.shape[0], acqs.shape[0])
    with torch.no_grad():
        return ops.pdff_unc(acqs.detach(), phi_mean, phi_var, r2_mean, r2_var, tab, r2_sc)


class _MagRegs(torch.autograd.Function):
    """One kernel gives the four sums and the gradient of their weighted sum; backward scales by the upstream on `total`."""

    @staticmethod
    def forward(ctx, ls, demod, r2, weights):
        sums, g_ls, g_demod, g_r2 = ops.mag_regs(None if ls is None else ls.detach(), None if demod is None else demod.detach(),
                                                 None if r2 is None else r2.detach(), weights)
        w = torch.tensor([weights[0], weights[1], 0.0, weights[2], weights[3]], dtype=torch.float32, device=sums.device)
        ctx.save_for_backward(*[t for t in (g_ls, g_demod, g_r2) if t is not None])
        ctx.present = [t is not None for t in (g_ls, g_demod, g_r2)]
        ctx.shapes = [None if t is None else t.shape for t in (ls, demod, r2)]
        ctx.mark_non_differentiable(sums)
        return (sums * w).sum(), sums

    @staticmethod
    def backward(ctx, g_total, _g_sums):
        saved = list(ctx.saved_tensors)
        grads = []
        for present, shape in zip(ctx.present, ctx.shapes):
            grads.append((saved.pop(0) * g_total).reshape(shape) if present else None)
        return grads[0], grads[1], grads[2], None


def mag_regularisers(ls=None, demod=None, r2=None, A_demod_TV_weight=0.0, LS_NZ_weight=0.0, LS_cond_weight=0.0, R2_TV_weight=0.0):
    """The terms train-IDEAL-mag.py:308-316 (+ R2_TV :288-289) adds to G_loss, from the CSE_mag outputs.

    Returns (total, logs): total = Ad_TV*w + LS_NZ*w + LS_cond*w + R2_TV*w (differentiable in ls, demod, r2) and
    logs = dict of the unweighted sums under the names the script logs."""
    total, sums = _MagRegs.apply(ls, demod, r2, (float(A_demod_TV_weight), float(LS_NZ_weight), float(LS_cond_weight), float(R2_TV_weight)))
    names = ("Ad_TV", "LS_NZ", "WF_NZ", "LS_cond", "R2_TV")
    return total, {n: sums[i] for i, n in enumerate(names)}


def roi_maps(maps, var=None, mode=None):
    """Inference-only map assembly / PDFF variance propagation of ROI-analysis.py:301-322."""
    with torch.no_grad():
        return ops.roi_maps(maps.detach(), None if var is None else var.detach(), mode)
