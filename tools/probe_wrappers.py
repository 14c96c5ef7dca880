#!/usr/bin/env python
"""Where the host time of a drop-in call goes (C1: one slice) and where the streamed C5 shard spends its wall time.
    python tools/probe_wrappers.py            (one GPU)"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ideal-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import wflib as wf  # noqa: E402
from idealgan import _lib as L  # noqa: E402
from idealgan import dist as igdist  # noqa: E402
from idealgan import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
H = W = 384
rng = np.random.default_rng(0)
maps = torch.from_numpy(synth.wfpm_maps(1, H, W, rng)).to(dev)
te = torch.from_numpy(synth.te_orig(1, 6)).to(dev)
layer = wf.IDEAL_Layer(field=1.5)
pm = maps[:, 2:3].contiguous()


def c1():
    sig = layer(maps, te=te, training=False)
    return wf.get_rho(sig, pm, field=1.5, te=te)


for _ in range(20):
    c1()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500):
    c1()
host = (time.perf_counter() - t0) / 500
torch.cuda.synchronize()
print(f"C1 host time per (IDEAL_Layer + get_rho) pair: {host * 1e6:.1f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    c1()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

# ---- C5 stream breakdown ---------------------------------------------------------------------------------------------------
nb, chunk, ne = 256, 32, 6
maps_h = igdist.pinned_empty((nb, 2, H, W, 3), dev)
maps_h.copy_(torch.from_numpy(synth.magpha_maps(8, H, W, rng, bipolar=False)).repeat(nb // 8, 1, 1, 1, 1))
te5 = synth.te_orig(nb, ne)
img = {"mag": igdist.pinned_empty((nb, ne, H, W), dev), "pdff": igdist.pinned_empty((nb, H, W), dev), "r2s": igdist.pinned_empty((nb, H, W), dev)}
print("pinned:", maps_h.is_pinned(), img["mag"].is_pinned())


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def h2d_only():
    for s in range(0, nb, chunk):
        maps_h[s:s + chunk].to(dev, non_blocking=True)


gb_in, gb_out = maps_h.numel() * 4 / 1e9, sum(t.numel() for t in img.values()) * 4 / 1e9
t = timed(h2d_only)
print(f"H2D only: {t * 1e3:.1f} ms  {gb_in / t:.1f} GB/s")
d_mag = torch.empty((chunk, ne, H, W), device=dev)


def d2h_only():
    for s in range(0, nb, chunk):
        img["mag"][s:s + chunk].copy_(d_mag, non_blocking=True)


t = timed(d2h_only)
print(f"D2H only (mag): {t * 1e3:.1f} ms  {img['mag'].numel() * 4 / 1e9 / t:.1f} GB/s")
for ch in (8, 32, 64):
    t = timed(lambda: igdist.synthesize_to_host(L.MODEL_MAGPHA, maps_h, te5, out_host=False, chunk_nb=ch, device=dev, images=img))
    print(f"synthesize_to_host chunk {ch}: {t * 1e3:.1f} ms  in {gb_in / t:.1f} GB/s out {gb_out / t:.1f} GB/s")
torch_pinned = torch.empty((nb, ne, H, W)).pin_memory()
t = timed(lambda: [torch_pinned[s:s + chunk].copy_(d_mag, non_blocking=True) for s in range(0, nb, chunk)])
print(f"D2H only into torch-pinned memory: {t * 1e3:.1f} ms  {torch_pinned.numel() * 4 / 1e9 / t:.1f} GB/s")
