"""Out-of-bounds WRITES of every kernel family, caught with guard bands: `compute-sanitizer` is closed on the GPU pool, so each
output tensor an operator allocates is carved out of a larger buffer filled with a canary value, and the bands on both sides must
be intact after the launch.  Shapes cover the dispatch rules that decide which kernel runs (tensor-map ring, bulk-copy ring, plain
kernels on packed and scalar lanes, ragged last tiles, one-row images)."""
import math

import numpy as np
import pytest
import torch

from idealgan import _lib as L
from idealgan import ops, synth

pytestmark = pytest.mark.gpu
CANARY = 31337.0
PAD = 256          # floats on each side: keeps the carved tensor 16-byte aligned like a fresh allocation


class GuardedTorch:
    """Stands in for the `torch` module inside idealgan.ops: allocations get guard bands, everything else is torch."""

    def __init__(self):
        self.guards = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def empty(self, *shape, dtype=torch.float32, device=None):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        n = math.prod(shape)
        assert dtype == torch.float32
        buf = torch.full((n + 2 * PAD,), CANARY, dtype=dtype, device=device)
        self.guards.append((buf, n, tuple(shape)))
        return buf[PAD:PAD + n].view(shape)

    def empty_like(self, t):
        return self.empty(tuple(t.shape), dtype=t.dtype, device=t.device)

    def check(self, what):
        torch.cuda.synchronize()
        assert self.guards, f"{what}: no output went through the guarded allocator"
        for buf, n, shape in self.guards:
            assert bool((buf[:PAD] == CANARY).all()), f"{what}: write BEFORE an output of shape {shape}"
            assert bool((buf[PAD + n:] == CANARY).all()), f"{what}: write PAST an output of shape {shape}"
            assert not bool((buf[PAD:PAD + n] == CANARY).all()) or n == 0, f"{what}: an output of shape {shape} was never written"
        self.guards.clear()


SHAPES = [  # (nb, H, W, ne)
    (2, 7, 9, 6),         # odd voxel count: scalar lanes everywhere
    (1, 8, 16, 6),        # one 128-voxel row: a mostly out-of-range TMA box
    (3, 30, 34, 5),       # even, not a multiple of 128: bulk-copy ring / plain packed kernels, ne below its bucket
    (2, 40, 48, 6),       # ragged last tile on the tensor-map ring
    (1, 64, 64, 8),
    (2, 16, 24, 12),      # more than 8 echoes: parked-y ring, plain kernels for the generic-ring operators
]


@pytest.mark.parametrize("shape", SHAPES)
def test_no_kernel_writes_outside_its_outputs(shape, monkeypatch):
    nb, H, W, ne = shape
    rng = np.random.default_rng(sum(shape))
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()      # noqa: E731
    te = dev(synth.te_random(nb, ne, rng, d_te_min=0.9e-3 if ne > 8 else 1.6e-3, d_te_d=0.3e-3 if ne > 8 else 1.0e-3))
    tab = ops.gen_tables(te, 1.5)
    wfpm = dev(synth.wfpm_maps(nb, H, W, rng, bipolar=True))
    ffpd = dev(synth.ffpd_maps(nb, H, W, rng))
    mp4, mp3 = dev(synth.magpha_maps(nb, H, W, rng, bipolar=True)), dev(synth.magpha_maps(nb, H, W, rng, bipolar=False))
    sig = ops.ideal_fwd(L.MODEL_WFPM, wfpm, tab, ne)
    acqs = torch.where(sig != 0, sig + 0.02 * torch.randn_like(sig), torch.zeros_like(sig)).contiguous()
    acqs[0, 0, H // 2, W // 2, 1] = 0.0                                    # one ragged voxel: the per-component paths run too
    pm = wfpm[:, 2:3].contiguous()
    up, up_rho = torch.randn_like(acqs), torch.randn((nb, 2, H, W, 2), device="cuda")
    plane = lambda lo, hi: (lo + (hi - lo) * torch.rand((nb, 1, H, W, 1), device="cuda")).contiguous()      # noqa: E731
    pv, rv, rm, phm = plane(1e-4, 4e-3), plane(1e-4, 3e-3), pm[..., 1:2].contiguous(), pm[..., 0:1].contiguous()
    mag = acqs.pow(2).sum(-1, keepdim=True).sqrt().contiguous()
    ups5 = [torch.randn((nb, c, H, W, 1), device="cuda") for c in (2, ne, ne, 3, 1)]
    guarded = GuardedTorch()
    monkeypatch.setattr(ops, "torch", guarded)
    calls = {
        "ideal_fwd wfpm": lambda: ops.ideal_fwd(L.MODEL_WFPM, wfpm, tab, ne),
        "ideal_fwd wfpm flat": lambda: ops.ideal_fwd(L.MODEL_WFPM, wfpm, tab, ne, flags=L.F_FLAT),
        "ideal_fwd ffpd": lambda: ops.ideal_fwd(L.MODEL_FFPD, ffpd, tab, ne),
        "ideal_fwd magpha 4ch": lambda: ops.ideal_fwd(L.MODEL_MAGPHA, mp4, tab, ne),
        "ideal_fwd magpha 3ch": lambda: ops.ideal_fwd(L.MODEL_MAGPHA, mp3, tab, ne),
        "ideal_decode": lambda: ops.ideal_decode(L.MODEL_MAGPHA, mp3, tab, ne, want_shat=True),
        "ideal_bwd wfpm": lambda: ops.ideal_bwd(L.MODEL_WFPM, wfpm, tab, ne, up),
        "ideal_bwd magpha 3ch": lambda: ops.ideal_bwd(L.MODEL_MAGPHA, mp3, tab, ne, up),
        "ideal_loss wfpm": lambda: ops.ideal_loss(L.MODEL_WFPM, wfpm, acqs, tab, want_shat=True),
        "ideal_loss magpha 4ch": lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp4, acqs, tab),
        "ideal_loss magpha 3ch": lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp3, acqs, tab),
        "get_rho_fwd": lambda: ops.get_rho_fwd(acqs, pm, tab, want_demod=True),
        "get_rho_maps": lambda: ops.get_rho_maps(acqs, pm, tab),
        "get_rho_bwd": lambda: ops.get_rho_bwd(acqs, pm, tab, up_rho, up),
        "get_rho_bwd phase-constrained": lambda: ops.get_rho_bwd(acqs, pm, tab, up_rho, None, flags=L.F_PHASE_CONSTRAINT),
        "a2a_fwd": lambda: ops.a2a_fwd(acqs, pm, tab),
        "a2a_fwd only_mag": lambda: ops.a2a_fwd(acqs, pm, tab, flags=L.F_ONLY_MAG),
        "a2a_bwd dPM": lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=False),
        "a2a_bwd dPM + dS": lambda: ops.a2a_bwd(acqs, pm, tab, up_rho, up, need_acqs=True),
        "a2a_loss": lambda: ops.a2a_loss(acqs, pm, tab),
        "a2a_loss + outputs": lambda: ops.a2a_loss(acqs, pm, tab, want_rho=True, want_shat=True),
        "a2a_uq_loss": lambda: ops.a2a_uq_loss(acqs, pm, pv, rm, rv, tab, want_rho=True),
        "a2a_uq_loss rem_R2": lambda: ops.a2a_uq_loss(acqs, pm, pv, None, None, tab),
        "a2a_rician_loss": lambda: ops.a2a_rician_loss(acqs, pm, pv, rm, rv, tab, want_rho=True),
        "cse_mag_fwd": lambda: ops.cse_mag_fwd(mag, rm, tab),
        "cse_mag_bwd": lambda: ops.cse_mag_bwd(mag, rm, tab, ups5),
        "acq_unc_fwd": lambda: ops.acq_unc_fwd(up_rho, pv, rm, rv, tab, ne),
        "acq_unc_bwd": lambda: ops.acq_unc_bwd(up_rho, pv, rm, rv, tab, ne, up),
        "pdff_unc": lambda: ops.pdff_unc(acqs, phm, pv, rm, rv, tab),
        "pdff_unc rem_R2": lambda: ops.pdff_unc(acqs, phm, pv, None, None, tab),
        "pdff_extract": lambda: ops.pdff_extract(up_rho),
        "roi_maps": lambda: ops.roi_maps(wfpm[:, :3].contiguous(), torch.rand((nb, 5, H, W, 2), device="cuda"), "PDFF-var"),
    }
    for what, call in calls.items():
        call()
        guarded.check(f"{what} at {shape}")
