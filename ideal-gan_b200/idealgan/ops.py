"""Pointer-level operators on torch CUDA tensors: thin, allocation + launch only.

Every function maps 1:1 onto an entry point of include/idealgan.h.  torch is used for device memory and
the current stream; all arithmetic happens in libidealgan.so.  Tensors must be float32, contiguous and
on the current CUDA device; shapes follow the reference.
"""
import torch

from . import _lib as L


PDFF_MODES = {"complex_sum": 0, "mag_sum": 1, "mag_disc": 2}


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw-handle call: torch.cuda.current_stream() builds a
    Stream object and costs ~5 us, a fifth of a small launch)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _chk(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name}: tensor must live on a CUDA device (this path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: dtype must be float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def gen_tables(te, field):
    """te (nb, ne[, 1]) seconds on the GPU -> (nb, TAB_FLOATS) table (gen_M / gen_A, IDEAL_model.py:48-97)."""
    te = _chk(te, "te")
    if te.dim() == 3:
        te = te[:, :, 0]
    te = te.contiguous()
    nb, ne = te.shape
    tab = torch.empty((nb, L.TAB_FLOATS), dtype=torch.float32, device=te.device)
    L.check(L.load().ig_gen_tables(te.data_ptr(), nb, ne, float(field), tab.data_ptr(), _stream()), "ig_gen_tables")
    return tab


def gen_tables_ahead(te, field, out):
    """The table of the NEXT batch, launched in front of the objective that still uses the previous one (ig_gen_tables_ahead: it overlaps
    the kernel before it in the stream and completes after it).  `te` must be complete already and `out`, a (nb, TAB_FLOATS) buffer, must
    not be in use by that kernel: rotate three buffers (step i: gen_tables_ahead(te[i + 1], field, tabs[(i + 1) % 3]); objective(tabs[i % 3]))."""
    te = _chk(te, "te")
    if te.dim() == 3:
        te = te[:, :, 0]
    if not te.is_contiguous():
        raise ValueError("gen_tables_ahead: te must be contiguous (a copy kernel in front of the table would defeat the look-ahead)")
    nb, ne = te.shape
    if tuple(out.shape) != (nb, L.TAB_FLOATS) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != te.device:
        raise ValueError(f"gen_tables_ahead: out must be a contiguous float32 ({nb}, {L.TAB_FLOATS}) tensor on {te.device}")
    L.check(L.load().ig_gen_tables_ahead(te.data_ptr(), nb, ne, float(field), out.data_ptr(), _stream()), "ig_gen_tables_ahead")
    return out


_scratch = {}


def loss_scratch(device, nb, nv):
    """Zero-initialised scratch for the *_loss kernels, cached per (device, stream); kernels re-zero it."""
    need = L.load().ig_loss_scratch_bytes(nb, nv)
    key = (device.index, _stream())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _model_dims(model, maps):
    maps = _chk(maps, "out_maps", 5)
    if model == L.MODEL_MAGPHA:
        nb, rows, H, W, ch = maps.shape
        if rows != 2 or ch not in (3, 4):
            raise ValueError(f"mag/phase maps must be (nb, 2, H, W, 3|4), got {tuple(maps.shape)}")
        return maps, nb, H, W, ch
    nb, rows, H, W, ch = maps.shape
    ok = rows >= 3 if model == L.MODEL_WFPM else rows == 3          # IDEAL_model.py:246: any row count above 3, the LAST row is the bipolar one
    if ch != 2 or not ok:
        raise ValueError(f"maps for model {model} must be (nb, 3{'+' if model == L.MODEL_WFPM else ''}, H, W, 2), got {tuple(maps.shape)}")
    return maps, nb, H, W, rows


def ideal_fwd(model, maps, tab, ne, r2_sc=200.0, flags=0):
    """flags & F_FLAT: the result comes out channel-interleaved, (nb, H, W, 2 ne) = data.A_from_MEBCRN(IDEAL_op(maps));
    flags & F_ONLY_MAG: the magnitudes |S_hat| (nb, ne, H, W, 1)."""
    maps, nb, H, W, roc = _model_dims(model, maps)
    shape = (nb, H, W, 2 * ne) if flags & L.F_FLAT else ((nb, ne, H, W, 1) if flags & L.F_ONLY_MAG else (nb, ne, H, W, 2))
    out = torch.empty(shape, dtype=torch.float32, device=maps.device)
    L.check(L.load().ig_ideal_fwd(model, maps.data_ptr(), roc, tab.data_ptr(), nb, ne, H * W, float(r2_sc), flags,
                                  out.data_ptr(), _stream()), "ig_ideal_fwd")
    return out


def ideal_decode(model, maps, tab, ne, r2_sc=200.0, want_shat=False, want_mag=True, want_pdff=True, want_r2s=True, clip=True):
    """Forward model + the images gen_LDM_dataset.py:217-235 writes per slice, one pass: (S_hat (nb,ne,H,W,2) | None,
    |S_hat| (nb,ne,H,W) | None, PDFF (nb,H,W) | None, R2* map (nb,H,W) | None), the images clipped to [0, 1] unless clip=False."""
    maps, nb, H, W, roc = _model_dims(model, maps)
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=maps.device)      # noqa: E731
    shat = new(nb, ne, H, W, 2) if want_shat else None
    mag = new(nb, ne, H, W) if want_mag else None
    pdff = new(nb, H, W) if want_pdff else None
    r2s = new(nb, H, W) if want_r2s else None
    L.check(L.load().ig_ideal_decode(model, maps.data_ptr(), roc, tab.data_ptr(), nb, ne, H * W, float(r2_sc), 0 if clip else L.F_NO_CLIP,
                                     _ptr(shat), _ptr(mag), _ptr(pdff), _ptr(r2s), _stream()), "ig_ideal_decode")
    return shat, mag, pdff, r2s


def ideal_bwd(model, maps, tab, ne, gout, r2_sc=200.0, flags=0):
    maps, nb, H, W, roc = _model_dims(model, maps)
    gout = _chk(gout, "grad_output", 5)
    gmaps = torch.empty_like(maps)
    L.check(L.load().ig_ideal_bwd(model, maps.data_ptr(), roc, tab.data_ptr(), nb, ne, H * W, float(r2_sc), flags,
                                  gout.data_ptr(), gmaps.data_ptr(), _stream()), "ig_ideal_bwd")
    return gmaps


def ideal_loss(model, maps, acqs, tab, r2_sc=200.0, flags=0, inv_n=None, want_shat=False):
    maps, nb, H, W, roc = _model_dims(model, maps)
    acqs = _chk(acqs, "acqs", 5)
    ne = acqs.shape[1]
    if acqs.shape != (nb, ne, H, W, 2):
        raise ValueError(f"acqs {tuple(acqs.shape)} does not match maps {tuple(maps.shape)}")
    inv_n = 1.0 / acqs.numel() if inv_n is None else inv_n
    gmaps = torch.empty_like(maps)
    shat = torch.empty_like(acqs) if want_shat else None
    loss = torch.empty(1, dtype=torch.float32, device=maps.device)
    scr = loss_scratch(maps.device, nb, H * W)
    L.check(L.load().ig_ideal_loss(model, maps.data_ptr(), roc, acqs.data_ptr(), tab.data_ptr(), nb, ne, H * W, float(r2_sc),
                                   flags, float(inv_n), gmaps.data_ptr(), _ptr(shat), loss.data_ptr(), scr.data_ptr(),
                                   scr.numel(), _stream()), "ig_ideal_loss")
    return loss, gmaps, shat


def _pm_view(pm, nb, H, W, flat):
    """(pointer tensor, batch stride in floats) of the (phi, R2*) row of a PM / WF-PM tensor."""
    if flat:
        pm = _chk(pm, "param_maps", 4)
        if pm.shape != (nb, H, W, 2):
            raise ValueError(f"flat param_maps must be (nb, H, W, 2), got {tuple(pm.shape)}")
        return pm, H * W * 2
    pm = _chk(pm, "param_maps", 5)
    if pm.shape[0] != nb or pm.shape[2:] != (H, W, 2):
        raise ValueError(f"param_maps {tuple(pm.shape)} does not match acquisitions (nb={nb}, H={H}, W={W})")
    return pm, pm.shape[1] * H * W * 2


def _acq_dims(acqs, flat):
    if flat:
        acqs = _chk(acqs, "acqs", 4)
        nb, H, W, c = acqs.shape
        if c % 2:
            raise ValueError("flat acquisitions need an even channel count (Re/Im interleaved)")
        return acqs, nb, c // 2, H, W
    acqs = _chk(acqs, "acqs", 5)
    nb, ne, H, W, c = acqs.shape
    if c != 2:
        raise ValueError(f"acqs must be (nb, ne, H, W, 2), got {tuple(acqs.shape)}")
    return acqs, nb, ne, H, W


def get_rho_fwd(acqs, pm, tab, r2_sc=200.0, flags=0, want_demod=False):
    flat = bool(flags & L.F_FLAT)
    acqs, nb, ne, H, W = _acq_dims(acqs, flat)
    pm, stride = _pm_view(pm, nb, H, W, flat)
    bip_ptr, bip_stride = 0, 0
    if not flat and pm.shape[1] > 3:                 # literal reference rule (IDEAL_model.py:567-568)
        bip_ptr, bip_stride = pm[:, -1].data_ptr(), stride
    rho = torch.empty((nb, H, W, 4) if flat else (nb, 2, H, W, 2), dtype=torch.float32, device=acqs.device)
    demod = torch.empty_like(acqs) if want_demod else None
    L.check(L.load().ig_get_rho_fwd(acqs.data_ptr(), pm.data_ptr(), stride, bip_ptr, bip_stride, tab.data_ptr(), nb, ne, H * W,
                                    float(r2_sc), flags, rho.data_ptr(), _ptr(demod), _stream()), "ig_get_rho_fwd")
    return rho, demod


def get_rho_maps(acqs, pm, tab, r2_sc=200.0, flags=0, pdff_mode="complex_sum"):
    """LS solve with the PDFF and R2* maps computed in the same pass: (rho, pdff (nb,H,W), r2s (nb,H,W))."""
    flat = bool(flags & L.F_FLAT)
    acqs, nb, ne, H, W = _acq_dims(acqs, flat)
    pm, stride = _pm_view(pm, nb, H, W, flat)
    bip_ptr, bip_stride = 0, 0
    if not flat and pm.shape[1] > 3:
        bip_ptr, bip_stride = pm[:, -1].data_ptr(), stride
    rho = torch.empty((nb, H, W, 4) if flat else (nb, 2, H, W, 2), dtype=torch.float32, device=acqs.device)
    pdff = torch.empty((nb, H, W), dtype=torch.float32, device=acqs.device)
    r2s = torch.empty((nb, H, W), dtype=torch.float32, device=acqs.device)
    L.check(L.load().ig_get_rho_maps(acqs.data_ptr(), pm.data_ptr(), stride, bip_ptr, bip_stride, tab.data_ptr(), nb, ne, H * W,
                                     float(r2_sc), flags, PDFF_MODES[pdff_mode], rho.data_ptr(), pdff.data_ptr(), r2s.data_ptr(), _stream()),
            "ig_get_rho_maps")
    return rho, pdff, r2s


def get_rho_bwd(acqs, pm, tab, g_rho, g_demod, r2_sc=200.0, flags=0, need_acqs=True):
    flat = bool(flags & L.F_FLAT)
    acqs, nb, ne, H, W = _acq_dims(acqs, flat)
    pm, stride = _pm_view(pm, nb, H, W, flat)
    g_rho = None if g_rho is None else _chk(g_rho, "grad rho")
    g_demod = None if g_demod is None else _chk(g_demod, "grad demod")
    bip_ptr = bip_stride = 0
    g_bip = None
    row = torch.empty((nb, H, W, 2), dtype=torch.float32, device=acqs.device)
    if not flat and pm.shape[1] > 3:
        bip_ptr, bip_stride = pm[:, -1].data_ptr(), stride
        g_bip = torch.empty((nb, H, W, 2), dtype=torch.float32, device=acqs.device)
    g_acqs = torch.empty_like(acqs) if need_acqs else None
    L.check(L.load().ig_get_rho_bwd(acqs.data_ptr(), pm.data_ptr(), stride, bip_ptr, bip_stride, tab.data_ptr(), nb, ne, H * W,
                                    float(r2_sc), flags, _ptr(g_rho), _ptr(g_demod), _ptr(g_acqs), row.data_ptr(), _ptr(g_bip),
                                    _stream()), "ig_get_rho_bwd")
    if flat:
        g_pm = row
    elif pm.shape[1] == 1:
        g_pm = row.unsqueeze(1)                 # the kernel's dense row IS the (nb, 1, H, W, 2) gradient: no fill, no copy
    else:
        g_pm = torch.zeros_like(pm)
        g_pm[:, 0] = row
        if g_bip is not None:
            g_pm[:, -1] = g_bip
    return g_acqs, g_pm


def a2a_fwd(acqs, pm, tab, r2_sc=200.0, flags=0, want_rho=True):
    acqs, nb, ne, H, W = _acq_dims(acqs, False)
    pm, stride = _pm_view(pm, nb, H, W, False)
    rho = torch.empty((nb, 2, H, W, 2), dtype=torch.float32, device=acqs.device) if want_rho else None
    shat = torch.empty((nb, ne, H, W, 1 if flags & L.F_ONLY_MAG else 2), dtype=torch.float32, device=acqs.device)
    L.check(L.load().ig_a2a_fwd(acqs.data_ptr(), pm.data_ptr(), stride, tab.data_ptr(), nb, ne, H * W, float(r2_sc), flags,
                                _ptr(rho), shat.data_ptr(), _stream()), "ig_a2a_fwd")
    return rho, shat


def a2a_bwd(acqs, pm, tab, g_rho, g_shat, r2_sc=200.0, flags=0, need_acqs=True):
    acqs, nb, ne, H, W = _acq_dims(acqs, False)
    pm, stride = _pm_view(pm, nb, H, W, False)
    g_rho = None if g_rho is None else _chk(g_rho, "grad rho")
    g_shat = None if g_shat is None else _chk(g_shat, "grad S_hat")
    row = torch.empty((nb, H, W, 2), dtype=torch.float32, device=acqs.device)
    g_acqs = torch.empty_like(acqs) if need_acqs else None
    L.check(L.load().ig_a2a_bwd(acqs.data_ptr(), pm.data_ptr(), stride, tab.data_ptr(), nb, ne, H * W, float(r2_sc), flags,
                                _ptr(g_rho), _ptr(g_shat), _ptr(g_acqs), row.data_ptr(), _stream()), "ig_a2a_bwd")
    if pm.shape[1] == 1:
        g_pm = row.unsqueeze(1)
    else:
        g_pm = torch.zeros_like(pm)
        g_pm[:, 0] = row
    return g_acqs, g_pm


def a2a_loss(acqs, pm, tab, r2_sc=200.0, inv_n=None, want_rho=False, want_shat=False):
    """Fused config-2 objective.  Returns (loss[1], g_pm (nb,1,H,W,2), rho | None, shat | None)."""
    acqs, nb, ne, H, W = _acq_dims(acqs, False)
    pm, stride = _pm_view(pm, nb, H, W, False)
    inv_n = 1.0 / acqs.numel() if inv_n is None else inv_n
    g_pm = torch.empty((nb, 1, H, W, 2), dtype=torch.float32, device=acqs.device)
    rho = torch.empty((nb, 2, H, W, 2), dtype=torch.float32, device=acqs.device) if want_rho else None
    shat = torch.empty_like(acqs) if want_shat else None
    loss = torch.empty(1, dtype=torch.float32, device=acqs.device)
    scr = loss_scratch(acqs.device, nb, H * W)
    L.check(L.load().ig_a2a_loss(acqs.data_ptr(), pm.data_ptr(), stride, tab.data_ptr(), nb, ne, H * W, float(r2_sc), float(inv_n),
                                 g_pm.data_ptr(), _ptr(rho), _ptr(shat), loss.data_ptr(), scr.data_ptr(), scr.numel(), _stream()),
            "ig_a2a_loss")
    return loss, g_pm, rho, shat


def _plane(t, name, nb, nv):
    """(nb, 1, H, W, 1)-like tensor -> contiguous (nb, nv) view."""
    t = _chk(t, name)
    if t.numel() != nb * nv:
        raise ValueError(f"{name}: expected {nb} x {nv} values, got shape {tuple(t.shape)}")
    return t.reshape(nb, nv)


def a2a_rician_loss(acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc=200.0, inv_n=None, want_rho=False):
    """Fused Rician (magnitude) objective of the R2* stage (ig_uq.cu); arguments and results as a2a_uq_loss, one loss element
    per (echo, voxel): inv_n defaults to 1 / (nb ne H W)."""
    return _uq_call("ig_a2a_rician_loss", acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc, inv_n, want_rho, per_component=False)


def a2a_uq_loss(acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc=200.0, inv_n=None, want_rho=False):
    """Fused uncertainty-aware objective (ig_uq.cu).  phi_var / r2_mean / r2_var: (nb,1,H,W,1) moment maps in network units
    (r2_mean = r2_var = None: rem_R2).  Returns (loss[1], g_pm (nb,1,H,W,2), g_phi_var, g_r2_mean | None, g_r2_var | None,
    rho | None), the moment gradients shaped like their inputs."""
    return _uq_call("ig_a2a_uq_loss", acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc, inv_n, want_rho, per_component=True)


def _uq_call(symbol, acqs, pm, phi_var, r2_mean, r2_var, tab, r2_sc, inv_n, want_rho, per_component):
    acqs, nb, ne, H, W = _acq_dims(acqs, False)
    pm, stride = _pm_view(pm, nb, H, W, False)
    if (r2_mean is None) != (r2_var is None):
        raise ValueError("r2_mean and r2_var go together (both None = rem_R2)")
    nv = H * W
    pv = _plane(phi_var, "phi_var", nb, nv)
    rm = None if r2_mean is None else _plane(r2_mean, "r2_mean", nb, nv)
    rv = None if r2_var is None else _plane(r2_var, "r2_var", nb, nv)
    inv_n = (1.0 / acqs.numel() if per_component else 2.0 / acqs.numel()) if inv_n is None else inv_n
    dev = acqs.device
    g_pm = torch.empty((nb, 1, H, W, 2), dtype=torch.float32, device=dev)
    g_pv = torch.empty(phi_var.shape, dtype=torch.float32, device=dev)
    g_rm = None if rm is None else torch.empty(r2_mean.shape, dtype=torch.float32, device=dev)
    g_rv = None if rv is None else torch.empty(r2_var.shape, dtype=torch.float32, device=dev)
    rho = torch.empty((nb, 2, H, W, 2), dtype=torch.float32, device=dev) if want_rho else None
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    scr = loss_scratch(dev, nb, nv)
    L.check(getattr(L.load(), symbol)(acqs.data_ptr(), pm.data_ptr(), stride, pv.data_ptr(), _ptr(rm), _ptr(rv), tab.data_ptr(), nb, ne, nv,
                                      float(r2_sc), float(inv_n), g_pm.data_ptr(), g_pv.data_ptr(), _ptr(g_rm), _ptr(g_rv), _ptr(rho),
                                      loss.data_ptr(), scr.data_ptr(), scr.numel(), _stream()), symbol)
    return loss, g_pm, g_pv, g_rm, g_rv, rho


# ---------------------------------------------------------------------------------------------------------------
# second tier (ig_tier2.cu)
# ---------------------------------------------------------------------------------------------------------------


def eigenvals_fwd(X):
    X = _chk(X, "X")
    if X.shape[-1] != 3:
        raise ValueError(f"eigenvals: last axis must hold (a, b, c), got {tuple(X.shape)}")
    n = X.numel() // 3
    xy = torch.empty(X.shape[:-1] + (2,), dtype=torch.float32, device=X.device)
    ratio = torch.empty(X.shape[:-1] + (1,), dtype=torch.float32, device=X.device)
    L.check(L.load().ig_eigenvals(X.data_ptr(), n, xy.data_ptr(), ratio.data_ptr(), _stream()), "ig_eigenvals")
    return xy, ratio


def eigenvals_bwd(X, g_xy, g_ratio):
    X = _chk(X, "X")
    gX = torch.empty_like(X)
    g_xy = None if g_xy is None else _chk(g_xy, "grad xy")
    g_ratio = None if g_ratio is None else _chk(g_ratio, "grad ratio")
    L.check(L.load().ig_eigenvals_bwd(X.data_ptr(), X.numel() // 3, _ptr(g_xy), _ptr(g_ratio), gX.data_ptr(), _stream()), "ig_eigenvals_bwd")
    return gX


def cse_mag_fwd(mag, r2, tab, r2_sc=200.0, r2nu=None):
    mag = _chk(mag, "acqs", 5)
    nb, ne, H, W, ch = mag.shape
    if ch != 1:
        raise ValueError(f"CSE_mag takes magnitudes (nb, ne, H, W, 1), got {tuple(mag.shape)}")
    nv = H * W
    r2 = _plane(r2, "out_maps", nb, nv)
    r2nu = None if r2nu is None else _plane(r2nu, "out_maps.nu", nb, nv)
    new = lambda c: torch.empty((nb, c, H, W, 1), dtype=torch.float32, device=mag.device)   # noqa: E731
    rho, fit, demod, ls, unc = new(2), new(ne), new(ne), new(3), new(1)
    L.check(L.load().ig_cse_mag_fwd(mag.data_ptr(), r2.data_ptr(), _ptr(r2nu), tab.data_ptr(), nb, ne, nv, float(r2_sc), rho.data_ptr(),
                                    fit.data_ptr(), demod.data_ptr(), ls.data_ptr(), unc.data_ptr(), _stream()), "ig_cse_mag_fwd")
    return rho, fit, demod, ls, unc


def cse_mag_bwd(mag, r2, tab, grads, r2_sc=200.0, r2nu=None):
    mag = _chk(mag, "acqs", 5)
    nb, ne, H, W, _ = mag.shape
    nv = H * W
    r2v = _plane(r2, "out_maps", nb, nv)
    r2nuv = None if r2nu is None else _plane(r2nu, "out_maps.nu", nb, nv)
    gs = [None if g is None else _chk(g, "upstream") for g in grads]
    g_mag, g_r2 = torch.empty_like(mag), torch.empty_like(r2v)
    g_nu = None if r2nu is None else torch.empty_like(r2nuv)
    L.check(L.load().ig_cse_mag_bwd(mag.data_ptr(), r2v.data_ptr(), _ptr(r2nuv), tab.data_ptr(), nb, ne, nv, float(r2_sc), *[_ptr(g) for g in gs],
                                    g_mag.data_ptr(), g_r2.data_ptr(), _ptr(g_nu), _stream()), "ig_cse_mag_bwd")
    return g_mag, g_r2.reshape(r2.shape), None if g_nu is None else g_nu.reshape(r2nu.shape)


def acq_unc_fwd(rho, phi_var, r2_mean, r2_var, tab, ne, r2_sc=200.0, only_mag=False):
    rho = _chk(rho, "rho_maps", 5)
    nb, rows, H, W, ch = rho.shape
    if rows < 2 or ch != 2:
        raise ValueError(f"rho_maps must be (nb, >=2, H, W, 2), got {tuple(rho.shape)}")
    rho = rho[:, :2].contiguous()
    nv = H * W
    pv = _plane(phi_var, "phi variance", nb, nv)
    rm = None if r2_mean is None else _plane(r2_mean, "R2* mean", nb, nv)
    rv = None if r2_var is None else _plane(r2_var, "R2* variance", nb, nv)
    out = torch.empty((nb, ne, H, W, 1 if only_mag else 2), dtype=torch.float32, device=rho.device)
    L.check(L.load().ig_acq_unc_fwd(rho.data_ptr(), pv.data_ptr(), _ptr(rm), _ptr(rv), tab.data_ptr(), nb, ne, nv, float(r2_sc), int(only_mag),
                                    out.data_ptr(), _stream()), "ig_acq_unc_fwd")
    return out


def acq_unc_bwd(rho, phi_var, r2_mean, r2_var, tab, ne, g_out, r2_sc=200.0, only_mag=False):
    rho = _chk(rho, "rho_maps", 5)[:, :2].contiguous()
    nb, _, H, W, _ = rho.shape
    nv = H * W
    pv = _plane(phi_var, "phi variance", nb, nv)
    rm = None if r2_mean is None else _plane(r2_mean, "R2* mean", nb, nv)
    rv = None if r2_var is None else _plane(r2_var, "R2* variance", nb, nv)
    g_out = _chk(g_out, "upstream")
    g_pv = torch.empty_like(pv)
    g_rm = None if rm is None else torch.empty_like(rm)
    g_rv = None if rv is None else torch.empty_like(rv)
    L.check(L.load().ig_acq_unc_bwd(rho.data_ptr(), pv.data_ptr(), _ptr(rm), _ptr(rv), tab.data_ptr(), nb, ne, nv, float(r2_sc), int(only_mag),
                                    g_out.data_ptr(), g_pv.data_ptr(), _ptr(g_rm), _ptr(g_rv), _stream()), "ig_acq_unc_bwd")
    return (g_pv.reshape(phi_var.shape), None if g_rm is None else g_rm.reshape(r2_mean.shape),
            None if g_rv is None else g_rv.reshape(r2_var.shape))


def pdff_unc(acqs, phi_mean, phi_var, r2_mean, r2_var, tab, r2_sc=200.0):
    acqs, nb, ne, H, W = _acq_dims(acqs, False)
    nv = H * W
    pm, pv = _plane(phi_mean, "phi mean", nb, nv), _plane(phi_var, "phi variance", nb, nv)
    rm = None if r2_mean is None else _plane(r2_mean, "R2* mean", nb, nv)
    rv = None if r2_var is None else _plane(r2_var, "R2* variance", nb, nv)
    rho = torch.empty((nb, 2, H, W, 2), dtype=torch.float32, device=acqs.device)
    cov = torch.empty((nb, 4, H, W, 1), dtype=torch.float32, device=acqs.device)
    L.check(L.load().ig_pdff_unc(acqs.data_ptr(), pm.data_ptr(), pv.data_ptr(), _ptr(rm), _ptr(rv), tab.data_ptr(), nb, ne, nv, float(r2_sc),
                                 rho.data_ptr(), cov.data_ptr(), _stream()), "ig_pdff_unc")
    return rho, cov




def pdff_extract(rho, mode="complex_sum"):
    rho = _chk(rho, "rho", 5)
    nb, rows, H, W, ch = rho.shape
    if rows < 2 or ch != 2:
        raise ValueError(f"rho must be (nb, >=2, H, W, 2), got {tuple(rho.shape)}")
    rho = rho[:, :2].contiguous()
    out = torch.empty((nb, H, W), dtype=torch.float32, device=rho.device)
    L.check(L.load().ig_pdff_extract(rho.data_ptr(), nb, H * W, PDFF_MODES[mode], out.data_ptr(), _stream()), "ig_pdff_extract")
    return out


_reg_scratch = {}


def mag_regs(ls=None, demod=None, r2=None, weights=(0.0, 0.0, 0.0, 0.0), want_grads=True):
    """Regularisers of train-IDEAL-mag.py:288-289,308-316 in one pass (ig_mag_regs).

    ls (nb,3,H,W,1) fit coefficients, demod (nb,ne,H,W,1) demodulated echoes, r2 (nb,1,H,W,1) or (nb,H,W,1) R2* map
    (any may be None).  weights = (A_demod_TV_weight, LS_NZ_weight, LS_cond_weight, R2_TV_weight).
    Returns (sums[5] = Ad_TV, LS_NZ, WF_NZ, LS_cond, R2_TV, g_ls, g_demod, g_r2): gradients of the weighted sum."""
    ref = next((t for t in (ls, demod, r2) if t is not None), None)
    if ref is None:
        raise ValueError("mag_regs: at least one of ls, demod, r2 is needed")
    ls = None if ls is None else _chk(ls, "ls", 5)
    demod = None if demod is None else _chk(demod, "demod", 5)
    r2 = None if r2 is None else _chk(r2, "r2")
    nb = ref.shape[0]
    H, W = (ls if ls is not None else demod).shape[2:4] if (ls is not None or demod is not None) else r2.shape[-3:-1]
    ne = 0
    if ls is not None and tuple(ls.shape) != (nb, 3, H, W, 1):
        raise ValueError(f"ls must be (nb, 3, H, W, 1), got {tuple(ls.shape)}")
    if demod is not None:
        ne = demod.shape[1]
        if tuple(demod.shape) != (nb, ne, H, W, 1):
            raise ValueError(f"demod must be (nb, ne, {H}, {W}, 1), got {tuple(demod.shape)}")
    if r2 is not None and (r2.numel() != nb * H * W or r2.shape[0] != nb or r2.shape[-1] != 1):
        raise ValueError(f"r2 must hold one (H, W, 1) map per sample, got {tuple(r2.shape)}")
    if len(weights) != 4:
        raise ValueError("weights = (A_demod_TV_weight, LS_NZ_weight, LS_cond_weight, R2_TV_weight)")
    lib = L.load()
    need = lib.ig_mag_regs_scratch_bytes(nb, H, W)
    key = (ref.device.index, _stream())
    buf = _reg_scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=ref.device)
        _reg_scratch[key] = buf
    sums = torch.empty(5, dtype=torch.float32, device=ref.device)
    g = [torch.empty_like(t) if (t is not None and want_grads) else None for t in (ls, demod, r2)]
    L.check(lib.ig_mag_regs(_ptr(ls), _ptr(demod), _ptr(r2), nb, ne, H, W, *(float(w) for w in weights), sums.data_ptr(),
                            _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), buf.data_ptr(), buf.numel(), _stream()), "ig_mag_regs")
    return sums, g[0], g[1], g[2]


ROI_MODES = {None: 0, "PDFF-var": 1, "PDFF-var-Mag": 2}


def roi_maps(maps, var=None, mode=None):
    """ROI-analysis.py:301-322: maps (nb,3,H,W,2) [, var (nb,5,H,W,2)] -> (nb,H,W,4|5) = |W|, |F|, |W+F|, R2* [, PDFF variance]."""
    maps = _chk(maps, "maps", 5)
    nb, rows, H, W, ch = maps.shape
    if rows != 3 or ch != 2:
        raise ValueError(f"maps must be (nb, 3, H, W, 2), got {tuple(maps.shape)}")
    m = ROI_MODES[mode]
    if m:
        var = _chk(var, "var", 5)
        if tuple(var.shape) != (nb, 5, H, W, 2):
            raise ValueError(f"var must be (nb, 5, {H}, {W}, 2), got {tuple(var.shape)}")
    out = torch.empty((nb, H, W, 5 if m else 4), dtype=torch.float32, device=maps.device)
    L.check(L.load().ig_roi_maps(maps.data_ptr(), _ptr(var) if m else 0, nb, H * W, m, out.data_ptr(), _stream()), "ig_roi_maps")
    return out
