"""Layout adapters of the reference's data.py (A_from_MEBCRN / B_from_MEBCRN / B_to_MEBCRN, data.py:262-329) on the GPU.

Same names, arguments and results as the reference functions, for torch CUDA tensors; each call is one read and one
write of the tensor (ig_layout.cu) instead of the reference's transpose / stack / reshape / concat chain.  All three
are linear, so their autograd backward is the matching adapter in the other direction.
"""
import torch

from . import _lib as L
from .ops import _chk, _stream

_TO_MODE = {"All": 0, "WF-PM": 1, "WF": 2, "PM": 3}


def _acq_to_flat(A):
    A = _chk(A, "A")
    if A.dim() != 5 or A.shape[-1] != 2:
        raise ValueError(f"A_from_MEBCRN: expected (nb, ne, H, W, 2), got {tuple(A.shape)}")
    nb, ne, H, W, _ = A.shape
    out = torch.empty((nb, H, W, 2 * ne), dtype=torch.float32, device=A.device)
    L.check(L.load().ig_acq_to_flat(A.data_ptr(), nb, ne, H * W, out.data_ptr(), _stream()), "ig_acq_to_flat")
    return out


def _acq_from_flat(F):
    F = _chk(F, "A")
    if F.dim() != 4 or F.shape[-1] % 2:
        raise ValueError(f"A_to_MEBCRN: expected (nb, H, W, 2 ne), got {tuple(F.shape)}")
    nb, H, W, c = F.shape
    out = torch.empty((nb, c // 2, H, W, 2), dtype=torch.float32, device=F.device)
    L.check(L.load().ig_acq_from_flat(F.data_ptr(), nb, c // 2, H * W, out.data_ptr(), _stream()), "ig_acq_from_flat")
    return out


class _AFromMEBCRN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        return _acq_to_flat(A)

    @staticmethod
    def backward(ctx, g):
        return _acq_from_flat(g.contiguous())


class _AToMEBCRN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, F):
        return _acq_from_flat(F)

    @staticmethod
    def backward(ctx, g):
        return _acq_to_flat(g.contiguous())


def A_from_MEBCRN(A):
    """(nb, ne, H, W, 2) -> (nb, H, W, 2 ne), Re/Im interleaved per echo (data.py:262-276)."""
    return _AFromMEBCRN.apply(A.contiguous())


def A_to_MEBCRN(A):
    """Inverse of A_from_MEBCRN (the reference has no such function; it is A_from_MEBCRN's adjoint)."""
    return _AToMEBCRN.apply(A.contiguous())


def B_from_MEBCRN(B, mag_and_phase=False, c_pha=3):
    """(nb, 3, H, W, 2) -> (nb, H, W, 6) = (W_re, W_im, F_re, F_im, R2*, phi); with mag_and_phase, (nb, 2, H, W, ch) ->
    (nb, H, W, 4 + 2 (ch - 2)) as data.py:279-299 (both species rotated by c_pha * pi * B[:, 1, ..., 1]).  Forward only."""
    B = _chk(B, "B")
    if B.dim() != 5:
        raise ValueError(f"B_from_MEBCRN: expected a 5-D tensor, got {tuple(B.shape)}")
    nb, rows, H, W, ch = B.shape
    if mag_and_phase:
        if rows != 2 or ch < 3:
            raise ValueError(f"B_from_MEBCRN(mag_and_phase=True): expected (nb, 2, H, W, >=3), got {tuple(B.shape)}")
        out = torch.empty((nb, H, W, 4 + 2 * (ch - 2)), dtype=torch.float32, device=B.device)
        L.check(L.load().ig_maps_to_flat(B.data_ptr(), nb, H * W, 1, ch, float(c_pha), out.data_ptr(), _stream()), "ig_maps_to_flat")
        return out
    if rows != 3 or ch != 2:
        raise ValueError(f"B_from_MEBCRN: expected (nb, 3, H, W, 2), got {tuple(B.shape)}")
    out = torch.empty((nb, H, W, 6), dtype=torch.float32, device=B.device)
    L.check(L.load().ig_maps_to_flat(B.data_ptr(), nb, H * W, 0, 2, 0.0, out.data_ptr(), _stream()), "ig_maps_to_flat")
    return out


def B_to_MEBCRN(B, mode="All"):
    """data.py:302-329: 'All' (nb,H,W,6) -> (nb,3,H,W,2); 'WF-PM' (nb,H,W,4) -> (nb,3,H,W,2); 'WF' (nb,H,W,2) -> (nb,2,H,W,2);
    'PM' (nb,H,W,2) -> (nb,1,H,W,2).  Forward only."""
    if mode not in _TO_MODE:
        raise ValueError(f"B_to_MEBCRN: unknown mode {mode!r}")
    B = _chk(B, "B")
    want = {"All": 6, "WF-PM": 4, "WF": 2, "PM": 2}[mode]
    if B.dim() != 4 or B.shape[-1] != want:
        raise ValueError(f"B_to_MEBCRN(mode={mode!r}): expected (nb, H, W, {want}), got {tuple(B.shape)}")
    nb, H, W, _ = B.shape
    rows = {"All": 3, "WF-PM": 3, "WF": 2, "PM": 1}[mode]
    out = torch.empty((nb, rows, H, W, 2), dtype=torch.float32, device=B.device)
    L.check(L.load().ig_maps_from_flat(B.data_ptr(), nb, H * W, _TO_MODE[mode], out.data_ptr(), _stream()), "ig_maps_from_flat")
    return out
