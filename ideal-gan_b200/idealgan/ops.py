"""Pointer-level operators on torch CUDA tensors: thin, allocation + launch only.

Every function maps 1:1 onto an entry point of include/idealgan.h.  torch is used for device memory and
the current stream; all arithmetic happens in libidealgan.so.  Tensors must be float32, contiguous and
on the current CUDA device; shapes follow the reference.
"""
import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


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
    ok = rows in (3, 4) if model == L.MODEL_WFPM else rows == 3
    if ch != 2 or not ok:
        raise ValueError(f"maps for model {model} must be (nb, 3{'|4' if model == L.MODEL_WFPM else ''}, H, W, 2), got {tuple(maps.shape)}")
    return maps, nb, H, W, rows


def ideal_fwd(model, maps, tab, ne, r2_sc=200.0, flags=0):
    maps, nb, H, W, roc = _model_dims(model, maps)
    out = torch.empty((nb, ne, H, W, 2), dtype=torch.float32, device=maps.device)
    L.check(L.load().ig_ideal_fwd(model, maps.data_ptr(), roc, tab.data_ptr(), nb, ne, H * W, float(r2_sc), flags,
                                  out.data_ptr(), _stream()), "ig_ideal_fwd")
    return out


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


def get_rho_bwd(acqs, pm, tab, g_rho, g_demod, r2_sc=200.0, flags=0, need_acqs=True):
    flat = bool(flags & L.F_FLAT)
    acqs, nb, ne, H, W = _acq_dims(acqs, flat)
    pm, stride = _pm_view(pm, nb, H, W, flat)
    g_rho = None if g_rho is None else _chk(g_rho, "grad rho")
    g_demod = None if g_demod is None else _chk(g_demod, "grad demod")
    g_pm = torch.zeros_like(pm)
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
    else:
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
