"""Seeded synthetic inputs for the IDEAL physics path (numpy only, no framework).

Follows the input recipe fixed in SURVEY.md §8d: water/fat maps in a centred disc (zeros outside, so
the `where(A != 0)` mask of the physics loss is exercised), field map in the tanh range, R2* in
[0, 1] with a share of negatives for the relu gate of `IDEAL_model`
(/root/reference/wflib/IDEAL_model.py:238), bipolar phase, and echo times drawn the way
`gen_TEvar` draws them (/root/reference/wflib/IDEAL_model.py:21-45) but independently per sample.

Layouts produced (all float32):
  maps "WF-PM"     (nb, 3|4, H, W, 2)   IDEAL_Layer / get_rho / acq_to_acq   (data.py:117-122)
  maps "ff/pd/pha" (nb, 3,   H, W, 2)   IDEAL_mag                            (data.py:99-115)
  maps "mag/pha"   (nb, 2,   H, W, 3|4) IDEAL_mag_phase                      (train-IDEAL-single.py:143-152)
  te               (nb, ne, 1) seconds
"""
import numpy as np

TE_ORIG_FIRST = 1.3e-3
TE_ORIG_STEP = 2.1e-3


def disc_mask(H, W, radius=0.45):
    yy, xx = np.mgrid[0:H, 0:W]
    cy, cx = (H - 1) / 2.0, (W - 1) / 2.0
    return (((yy - cy) / H) ** 2 + ((xx - cx) / W) ** 2) <= radius ** 2


def te_orig(nb, ne):
    """The `orig=True` echo train, one row per sample: 1.3 ms + 2.1 ms * k."""
    te = TE_ORIG_FIRST + TE_ORIG_STEP * np.arange(ne, dtype=np.float64)
    return np.tile(te.astype(np.float32)[None, :, None], (nb, 1, 1))


def te_random(nb, ne, rng, te_ini_min=1.0e-3, te_ini_d=1.4e-3, d_te_min=1.6e-3, d_te_d=1.0e-3):
    """Per-sample random echo trains with the distribution of `gen_TEvar`'s random branch."""
    te = np.empty((nb, ne, 1), dtype=np.float32)
    for b in range(nb):
        te1 = te_ini_min + rng.uniform(0.0, te_ini_d)
        dc = d_te_min + rng.uniform(0.0, d_te_d)
        d = np.concatenate(([0.0], rng.normal(dc, 1e-4, size=ne - 1)))
        te[b, :, 0] = (np.cumsum(d) + te1).astype(np.float32)
    return te


def wfpm_maps(nb, H, W, rng, bipolar=False, neg_r2_frac=0.1, masked=True):
    rows = 4 if bipolar else 3
    m = np.zeros((nb, rows, H, W, 2), dtype=np.float32)
    m[:, 0:2] = rng.uniform(-0.5, 0.5, size=(nb, 2, H, W, 2))
    m[:, 2, :, :, 0] = rng.uniform(-1.0, 1.0, size=(nb, H, W))
    r2 = rng.uniform(0.0, 1.0, size=(nb, H, W))
    if neg_r2_frac > 0:
        neg = rng.uniform(size=(nb, H, W)) < neg_r2_frac
        r2 = np.where(neg, -0.2 * r2, r2)
    m[:, 2, :, :, 1] = r2
    if bipolar:
        m[:, 3, :, :, 0] = rng.uniform(-0.25, 0.25, size=(nb, H, W))
    if masked:
        m *= disc_mask(H, W)[None, None, :, :, None].astype(np.float32)
    return m


def ffpd_maps(nb, H, W, rng, masked=True):
    m = np.zeros((nb, 3, H, W, 2), dtype=np.float32)
    m[:, 0, :, :, 0] = rng.uniform(0.0, 1.0, size=(nb, H, W))        # PDFF
    m[:, 1, :, :, 0] = rng.uniform(0.0, 1.0, size=(nb, H, W))        # PD
    m[:, 1, :, :, 1] = rng.uniform(0.0, 1.0, size=(nb, H, W))        # R2*/200
    m[:, 2, :, :, 0] = rng.uniform(-0.25, 0.25, size=(nb, H, W))     # common phase / 4pi
    m[:, 2, :, :, 1] = rng.uniform(-1.0, 1.0, size=(nb, H, W))       # field map / 300
    if masked:
        m *= disc_mask(H, W)[None, None, :, :, None].astype(np.float32)
    return m


def magpha_maps(nb, H, W, rng, bipolar=True, masked=True):
    ch = 4 if bipolar else 3
    m = np.zeros((nb, 2, H, W, ch), dtype=np.float32)
    m[:, 0, :, :, 0:2] = rng.uniform(0.0, 0.7, size=(nb, H, W, 2))   # |W|, |F|
    m[:, 0, :, :, 2] = rng.uniform(0.0, 1.0, size=(nb, H, W))        # R2*/200
    m[:, 1, :, :, 0:2] = rng.uniform(-0.25, 0.25, size=(nb, H, W, 2))  # phases / 4pi
    m[:, 1, :, :, 2] = rng.uniform(-1.0, 1.0, size=(nb, H, W))       # field map / 300
    if bipolar:
        m[:, 1, :, :, 3] = rng.uniform(-0.06, 0.06, size=(nb, H, W))  # bipolar / 4pi
    if masked:
        m *= disc_mask(H, W)[None, None, :, :, None].astype(np.float32)
    return m


def add_noise(acqs, rng, sigma=0.02, keep_zeros=True):
    """acqs + N(0, sigma^2); voxels that were exactly zero stay zero (background)."""
    noisy = acqs + rng.normal(0.0, sigma, size=acqs.shape).astype(np.float32)
    if keep_zeros:
        noisy = np.where(acqs != 0, noisy, 0.0)
    return noisy.astype(np.float32)
