#!/usr/bin/env python
"""A/B harness for kernel experiments: times a set of entry points at 64 x 384 x 384 x 6 on disc-masked and on unmasked
data and reports the worst error of the fused C2 objective against the fp64 oracle on two unmasked slices.

    IDEALGAN_LIB=ideal-gan_b200/idealgan/libidealgan_x.so python tools/ab_kernels.py [--which loss,bwd,...]

(IDEALGAN_LIB selects a variant build made with `make -C ideal-gan_b200/csrc SUFFIX=_x EXTRA=-D...`.)  One JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ideal-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from idealgan import _lib as L  # noqa: E402
from idealgan import ops, synth  # noqa: E402


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def device_batch(nb, H, W, ne, masked, seed, te=None):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    maps = torch.empty((nb, 3, H, W, 2), device=dev)
    maps[:, :2] = torch.rand((nb, 2, H, W, 2), device=dev, generator=g) - 0.5
    maps[:, 2, :, :, 0] = 2 * torch.rand((nb, H, W), device=dev, generator=g) - 1
    maps[:, 2, :, :, 1] = torch.rand((nb, H, W), device=dev, generator=g)
    if masked:
        maps *= torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).to(dev)[None, None, :, :, None]
    te = torch.from_numpy(synth.te_orig(nb, ne) if te is None else te).to(dev)
    tab = ops.gen_tables(te, 1.5)
    sig = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne)
    noise = 0.02 * torch.randn(sig.shape, device=dev, generator=g)
    acqs = torch.where(sig != 0, sig + noise, torch.zeros_like(sig)).contiguous()
    pm = (maps[:, 2:3] * 0.95).contiguous()
    return maps, tab, acqs, pm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="loss,bwd,bwd_ds,c4,uq,rician")
    ap.add_argument("--ne", type=int, default=6)
    args = ap.parse_args()
    which = set(args.which.split(","))
    out = {"lib": os.path.basename(L.LIB_PATH), "ne": args.ne}
    nb, H, W, ne = 64, 384, 384, args.ne
    nv = H * W
    dev = torch.device("cuda", 0)
    te_full = None if ne <= 6 else synth.te_random(nb, ne, np.random.default_rng(5), d_te_min=0.9e-3, d_te_d=0.3e-3)
    for masked in (True, False):
        tag = "masked" if masked else "unmasked"
        maps, tab, acqs, pm = device_batch(nb, H, W, ne, masked, 1234, te_full)
        g = torch.Generator(device=dev)
        g.manual_seed(7)

        def rec(name, bytes_per_voxel, fn):
            ms = timeit(fn)
            out[f"{name}_{tag}"] = {"ms": round(ms, 5), "GBps": round(bytes_per_voxel * nb * nv / ms / 1e6, 1)}

        if "loss" in which:
            rec("a2a_loss", 8 * ne + 16, lambda: ops.a2a_loss(acqs, pm, tab))
        if "loss_out" in which:
            rec("a2a_loss_out", 16 * ne + 32, lambda: ops.a2a_loss(acqs, pm, tab, want_rho=True, want_shat=True))
        if "plain" in which:
            # the operators an unmodified train-IDEAL-TEaug.py step reaches (IDEAL_Layer, get_rho and their adjoints, 2..12 echoes) and acq_to_acq forward
            up_r = torch.randn((nb, 2, H, W, 2), device=dev, generator=g)
            rec("ideal_fwd", 8 * ne + 24, lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne))
            rec("ideal_bwd", 8 * ne + 48, lambda: ops.ideal_bwd(L.MODEL_WFPM, maps, tab, ne, acqs))
            rec("ideal_loss_wfpm", 8 * ne + 48, lambda: ops.ideal_loss(L.MODEL_WFPM, maps, acqs, tab))
            rec("get_rho_fwd", 8 * ne + 24, lambda: ops.get_rho_fwd(acqs, pm, tab))
            rec("get_rho_bwd", 16 * ne + 32, lambda: ops.get_rho_bwd(acqs, pm, tab, up_r, None))
            rec("a2a_fwd", 16 * ne + 24, lambda: ops.a2a_fwd(acqs, pm, tab))
            del up_r
        if "bwd" in which or "bwd_ds" in which:
            up = torch.randn(acqs.shape, device=dev, generator=g)
            if "bwd" in which:
                rec("a2a_bwd", 16 * ne + 16, lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=False))
            if "bwd_ds" in which:
                rec("a2a_bwd_ds", 24 * ne + 16, lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=True))
            del up
        if "uq" in which or "rician" in which or "pdff" in which:
            pv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 4e-3
            rv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 3e-3
            rm = pm[..., 1:2].contiguous()
            if "uq" in which:
                rec("a2a_uq_loss", 8 * ne + 40, lambda: ops.a2a_uq_loss(acqs, pm, pv, rm, rv, tab))
            if "pdff" in which:
                phm = pm[..., 0:1].contiguous()
                rec("pdff_unc", 8 * ne + 48, lambda: ops.pdff_unc(acqs, phm, pv, rm, rv, tab))
            if "rician" in which:
                rec("a2a_rician_loss", 8 * ne + 40, lambda: ops.a2a_rician_loss(acqs, pm, pv, rm, rv, tab))
        if "regs" in which:
            ls_ = torch.randn((nb, 3, H, W, 1), device=dev, generator=g)
            dm_ = torch.rand((nb, ne, H, W, 1), device=dev, generator=g)
            r2_ = torch.rand((nb, 1, H, W, 1), device=dev, generator=g)
            rec("mag_regs", 2 * (4 * ne + 12 + 4), lambda: ops.mag_regs(ls_, dm_, r2_, (0.1, 0.2, 0.3, 0.4)))
            del ls_, dm_, r2_
        if "c4" in which:
            mp = torch.rand((nb, 2, H, W, 4), device=dev, generator=g) * 0.5
            mp[:, 1] -= 0.25
            if masked:
                mp *= torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).to(dev)[None, None, :, :, None]
            a4 = ops.ideal_fwd(L.MODEL_MAGPHA, mp, tab, ne)
            a4 = torch.where(a4 != 0, a4 + 0.02 * torch.randn(a4.shape, device=dev, generator=g), torch.zeros_like(a4)).contiguous()
            rec("ideal_loss_magpha", 8 * ne + 64, lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp, a4, tab))
            del mp, a4
        del maps, tab, acqs, pm
        torch.cuda.empty_cache()
    # accuracy of the fused objective against the fp64 oracle (unmasked: every voxel carries phase up to the full range)
    if "loss" in which and ne == 6:
        from oracle import ideal_oracle as orc
        rng = np.random.default_rng(42)
        nb2 = 2
        te = synth.te_orig(nb2, ne)
        m = synth.wfpm_maps(nb2, H, W, rng, neg_r2_frac=0.0, masked=False)
        with torch.no_grad():
            sig = orc.IDEAL_model(torch.from_numpy(m), [1.5, torch.from_numpy(te)]).numpy()
        a = synth.add_noise(sig, rng)
        pm_h = np.ascontiguousarray(m[:, 2:3]) * np.float32(0.95)
        p = torch.from_numpy(pm_h).double().requires_grad_(True)
        lref, _, _ = orc.physics_loss_a2a(torch.from_numpy(a), p, te=torch.from_numpy(te), rdtype=torch.float64)
        (gref,) = torch.autograd.grad(lref, [p])
        tab = ops.gen_tables(torch.from_numpy(te).cuda(), 1.5)
        loss, gk, _, _ = ops.a2a_loss(torch.from_numpy(a).cuda(), torch.from_numpy(pm_h).cuda(), tab)
        gk = gk.cpu().double()
        out["loss_rel_err_vs_fp64"] = abs(loss.item() - lref.item()) / lref.item()
        out["grad_rel_to_max_err_vs_fp64"] = ((gk - gref).abs().max() / gref.abs().max()).item()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
