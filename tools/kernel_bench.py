#!/usr/bin/env python
"""Per-kernel timing of every entry point at the BASELINE configs, with the algorithmic bytes per voxel
and the resulting fraction of the measured HBM roofline.  Writes one JSON document to stdout.

    python tools/kernel_bench.py [--reps 20] > profiles/kernels_rNN.json
"""
import argparse
import json
import os
import time
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ideal-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from idealgan import _lib as L  # noqa: E402
from idealgan import ops, synth  # noqa: E402


def timeit(fn, reps):
    # every row starts from an idle board: after a few hundred milliseconds of back-to-back launches the power cap engages (sw_power_cap,
    # SM clock ~1.82 GHz) and the issue-bound kernels (C2 / UQ / Rician objectives) measure 4-7 % slower than alone, the HBM-bound ones do not
    time.sleep(0.5)
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
    time.sleep(0.5)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # SURVEY 8d's method beside it: one event pair around 2 x reps back-to-back launches (keeps launch gaps and PDL overlap as a loop has them)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2 * reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return ts[len(ts) // 2], ts[0], a.elapsed_time(b) / (2 * reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    rows = []

    def add(name, cfg, nb, nv, ne, bytes_per_voxel, fn):
        med, best, b2b = timeit(fn, args.reps)
        gbs = bytes_per_voxel * nb * nv / (med * 1e-3) / 1e9
        rows.append({"kernel": name, "config": cfg, "nb": nb, "nv": nv, "ne": ne, "bytes_per_voxel": bytes_per_voxel,
                     "ms_median": med, "ms_best": best, "GBps": gbs, "frac_of_measured_hbm": gbs / peak,
                     "voxel_echoes_per_s": nb * nv * ne / (med * 1e-3),
                     "ms_back_to_back": b2b, "frac_back_to_back": bytes_per_voxel * nb * nv / (b2b * 1e-3) / 1e9 / peak})

    def make(nb, H, W, ne, bip=False, te_random=False):
        rng = np.random.default_rng(1234)
        mask = torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).to(dev)[None, None, :, :, None]
        maps = torch.empty((nb, 4 if bip else 3, H, W, 2), device=dev)
        maps[:, :2] = torch.rand((nb, 2, H, W, 2), device=dev, generator=g) - 0.5
        maps[:, 2, :, :, 0] = 2 * torch.rand((nb, H, W), device=dev, generator=g) - 1
        maps[:, 2, :, :, 1] = torch.rand((nb, H, W), device=dev, generator=g)
        if bip:
            maps[:, 3] = 0.5 * torch.rand((nb, H, W, 2), device=dev, generator=g) - 0.25
        maps *= mask
        te = torch.from_numpy(synth.te_random(nb, ne, rng) if te_random else synth.te_orig(nb, ne)).to(dev)
        tab = ops.gen_tables(te, 1.5)
        sig = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne)
        acqs = torch.where(sig != 0, sig + 0.02 * torch.randn(sig.shape, device=dev, generator=g), torch.zeros_like(sig)).contiguous()
        return maps, te, tab, acqs

    # C2-sized tensors: 64 x 384 x 384 x 6
    nb, H, W, ne = 64, 384, 384, 6
    nv = H * W
    maps, te, tab, acqs = make(nb, H, W, ne)
    pm = (maps[:, 2:3] * 0.95).contiguous()
    up = torch.randn_like(acqs)
    up_rho = torch.randn((nb, 2, H, W, 2), device=dev)
    add("ig_gen_tables", "C2", nb, nv, ne, 0, lambda: ops.gen_tables(te, 1.5))
    add("ig_ideal_fwd[wfpm]", "C1/C3 forward", nb, nv, ne, 24 + 8 * ne, lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne))
    add("ig_ideal_fwd[wfpm,flat]", "forward with interleaved output", nb, nv, ne, 24 + 8 * ne,
        lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne, flags=L.F_FLAT))
    add("ig_ideal_bwd[wfpm]", "adjoint", nb, nv, ne, 8 * ne + 24 + 24, lambda: ops.ideal_bwd(L.MODEL_WFPM, maps, tab, ne, up))
    add("ig_ideal_loss[wfpm]", "fused fwd+mask+MSE+bwd", nb, nv, ne, 8 * ne + 24 + 24, lambda: ops.ideal_loss(L.MODEL_WFPM, maps, acqs, tab))
    add("ig_get_rho_fwd", "C1 LS solve", nb, nv, ne, 8 * ne + 8 + 16, lambda: ops.get_rho_fwd(acqs, pm, tab))
    add("ig_get_rho_bwd", "LS solve adjoint (dPM + dS)", nb, nv, ne, 8 * ne + 8 + 16 + 8 * ne + 8,
        lambda: ops.get_rho_bwd(acqs, pm, tab, up_rho, None))
    add("ig_a2a_fwd", "acq_to_acq forward (rho + S_hat)", nb, nv, ne, 8 * ne + 8 + 16 + 8 * ne, lambda: ops.a2a_fwd(acqs, pm, tab))
    add("ig_a2a_bwd", "acq_to_acq adjoint (dPM only)", nb, nv, ne, 8 * ne + 8 + 8 * ne + 8,
        lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=False))
    add("ig_a2a_bwd[+dS]", "acq_to_acq adjoint (dPM + dS)", nb, nv, ne, 8 * ne + 8 + 8 * ne + 8 + 8 * ne,
        lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=True))
    add("ig_a2a_loss", "C2 fused objective (headline)", nb, nv, ne, 8 * ne + 8 + 8, lambda: ops.a2a_loss(acqs, pm, tab))
    add("ig_a2a_loss[+rho,S_hat]", "C2 fused + materialised outputs", nb, nv, ne, 8 * ne + 8 + 8 + 16 + 8 * ne,
        lambda: ops.a2a_loss(acqs, pm, tab, want_rho=True, want_shat=True))
    pv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 4e-3
    rv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 3e-3
    rm = pm[..., 1:2].contiguous()
    add("ig_a2a_uq_loss", "AI-DEAL UQ objective (fused, all gradients)", nb, nv, ne, 8 * ne + 8 + 12 + 8 + 12,
        lambda: ops.a2a_uq_loss(acqs, pm, pv, rm, rv, tab))
    add("ig_a2a_rician_loss", "AI-DEAL R2* stage objective (Rician, fused)", nb, nv, ne, 8 * ne + 8 + 12 + 8 + 12,
        lambda: ops.a2a_rician_loss(acqs, pm, pv, rm, rv, tab))
    # second tier: magnitude fit, uncertainty propagation
    mag = torch.sqrt((acqs ** 2).sum(-1, keepdim=True)).contiguous()
    r2map = pm[..., 1:2].contiguous().reshape(nb, 1, H, W, 1)
    add("ig_cse_mag_fwd", "CSE_mag (rho, fit, demod, ls, unc)", nb, nv, ne, 12 * ne + 28,
        lambda: ops.cse_mag_fwd(mag, r2map, tab))
    ups_cse = [torch.randn((nb, c_, H, W, 1), device=dev, generator=g) for c_ in (2, ne, ne, 3, 1)]
    add("ig_cse_mag_bwd", "CSE_mag adjoint (all five upstreams -> d mag, d R2*)", nb, nv, ne, 4 * ne + 4 + 4 * (2 + ne + ne + 3 + 1) + 4 * ne + 4,
        lambda: ops.cse_mag_bwd(mag, r2map, tab, ups_cse))
    del ups_cse
    rho_hat, _ = ops.a2a_fwd(acqs, pm, tab)
    add("ig_acq_unc_fwd", "acq_uncertainty", nb, nv, ne, 16 + 12 + 8 * ne, lambda: ops.acq_unc_fwd(rho_hat, pv, rm, rv, tab, ne))
    up_var = torch.randn((nb, ne, H, W, 2), device=dev, generator=g)
    add("ig_acq_unc_bwd", "acq_uncertainty adjoint (d phi_var, d R2* mean, d R2* var)", nb, nv, ne, 16 + 12 + 8 * ne + 12,
        lambda: ops.acq_unc_bwd(rho_hat, pv, rm, rv, tab, ne, up_var))
    del up_var
    phm = pm[..., 0:1].contiguous()
    add("ig_pdff_unc", "PDFF_uncertainty (weighted LS per voxel)", nb, nv, ne, 8 * ne + 16 + 16 + 16,
        lambda: ops.pdff_unc(acqs, phm, pv, rm, rv, tab))
    add("ig_pdff_extract", "PDFF map", nb, nv, ne, 16 + 4, lambda: ops.pdff_extract(rho_hat))
    _, _, demod_s, ls_s, _ = ops.cse_mag_fwd(mag, r2map, tab)
    add("ig_mag_regs", "train-IDEAL-mag regularisers (4 sums + gradients)", nb, nv, ne, 2 * (4 * ne + 12 + 4),
        lambda: ops.mag_regs(ls_s.reshape(nb, 3, H, W, 1), demod_s.reshape(nb, ne, H, W, 1), r2map, (0.1, 0.2, 0.3, 0.4)))
    var5 = torch.rand((nb, 5, H, W, 2), device=dev, generator=g) * 1e-3
    add("ig_roi_maps", "ROI-analysis map assembly + PDFF variance", nb, nv, ne, 24 + 40 + 20,
        lambda: ops.roi_maps(maps, var5, "PDFF-var"))
    del mag, rho_hat, demod_s, ls_s, var5
    flat = torch.empty((nb, H, W, 2 * ne), device=dev)
    lib = L.load()
    add("ig_acq_to_flat", "A_from_MEBCRN (planar -> interleaved)", nb, nv, ne, 16 * ne,
        lambda: L.check(lib.ig_acq_to_flat(acqs.data_ptr(), nb, ne, nv, flat.data_ptr(), torch.cuda.current_stream().cuda_stream), "to_flat"))
    back = torch.empty_like(acqs)
    add("ig_acq_from_flat", "interleaved -> planar", nb, nv, ne, 16 * ne,
        lambda: L.check(lib.ig_acq_from_flat(flat.data_ptr(), nb, ne, nv, back.data_ptr(), torch.cuda.current_stream().cuda_stream), "from_flat"))
    pm_flat = torch.stack([pm[:, 0, :, :, 1], pm[:, 0, :, :, 0]], dim=-1).contiguous()        # flat layout: (R2*, phi)
    add("ig_get_rho_fwd[flat]", "LS solve on interleaved acquisitions", nb, nv, ne, 8 * ne + 8 + 16,
        lambda: ops.get_rho_fwd(flat, pm_flat, tab, flags=L.F_FLAT))
    del flat, back, pm_flat
    # C4: bipolar mag/phase fused objective
    mp = torch.rand((nb, 2, H, W, 4), device=dev, generator=g) * 0.5
    mp[:, 1] -= 0.25
    mp *= torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).to(dev)[None, None, :, :, None]
    add("ig_ideal_fwd[magpha]", "C4 forward", nb, nv, ne, 32 + 8 * ne, lambda: ops.ideal_fwd(L.MODEL_MAGPHA, mp, tab, ne))
    add("ig_ideal_loss[magpha]", "C4 fused objective", nb, nv, ne, 8 * ne + 32 + 32, lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp, acqs, tab))
    ff = torch.rand((nb, 3, H, W, 2), device=dev, generator=g)
    add("ig_ideal_fwd[ffpd]", "C5 forward", nb, nv, ne, 24 + 8 * ne, lambda: ops.ideal_fwd(L.MODEL_FFPD, ff, tab, ne))
    del maps, acqs, up, mp, ff
    torch.cuda.empty_cache()
    # C3: 256 x 192 x 192 x 6, per-sample random echo times
    nb, H, W = 256, 192, 192
    nv = H * W
    maps, te, tab, acqs = make(nb, H, W, ne, te_random=True)
    add("ig_ideal_fwd[wfpm]", "C3 per-sample TEs 256x192x192", nb, nv, ne, 24 + 8 * ne, lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne))
    pm3 = maps[:, 2:3].contiguous()
    add("ig_get_rho_fwd", "C3", nb, nv, ne, 8 * ne + 8 + 16, lambda: ops.get_rho_fwd(acqs, pm3, tab))
    up3 = torch.randn((nb, 2, H, W, 2), device=dev)
    add("ig_get_rho_bwd", "C3 (dPM + dS)", nb, nv, ne, 8 * ne + 8 + 16 + 8 * ne + 8, lambda: ops.get_rho_bwd(acqs, pm3, tab, up3, None))
    del maps, acqs
    torch.cuda.empty_cache()
    # C1: single slice latency
    maps, te, tab, acqs = make(1, 384, 384, ne)
    pm1 = maps[:, 2:3].contiguous()
    add("ig_ideal_fwd[wfpm]", "C1 single slice (latency)", 1, 384 * 384, ne, 24 + 8 * ne, lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne))
    add("ig_get_rho_fwd", "C1 single slice (latency)", 1, 384 * 384, ne, 8 * ne + 8 + 16, lambda: ops.get_rho_fwd(acqs, pm1, tab))
    # C1 as a latency figure: tables + forward + LS solve of ONE slice replayed from a CUDA graph (no Python between the launches)
    te1 = torch.from_numpy(synth.te_orig(1, ne)).to(dev)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(3):
            t1 = ops.gen_tables(te1, 1.5)
            s1 = ops.ideal_fwd(L.MODEL_WFPM, maps, t1, ne)
            ops.get_rho_fwd(s1, pm1, t1)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            t1 = ops.gen_tables(te1, 1.5)
            s1 = ops.ideal_fwd(L.MODEL_WFPM, maps, t1, ne)
            r1, _ = ops.get_rho_fwd(s1, pm1, t1)
    add("graph[tables+fwd+solve]", "C1 single slice, CUDA-graph replay (latency)", 1, 384 * 384, ne, 144, graph.replay)
    # C4 at the script's own batch size (train-IDEAL-single.py: 3 slices): tables + fused mag/phase objective from a CUDA graph
    nb4 = 3
    rng4 = np.random.default_rng(4)
    mp4 = torch.from_numpy(synth.magpha_maps(nb4, 384, 384, rng4)).to(dev)
    te4 = torch.from_numpy(np.ascontiguousarray(synth.te_random(nb4, ne, rng4, te_ini_d=0.4e-3, d_te_min=0.9e-3, d_te_d=0.3e-3)[:, :, 0])).to(dev)
    a4 = ops.ideal_fwd(L.MODEL_MAGPHA, mp4, ops.gen_tables(te4, 1.5), ne)
    a4 = torch.where(a4 != 0, a4 + 0.02 * torch.randn(a4.shape, device=dev, generator=g), torch.zeros_like(a4)).contiguous()
    with torch.cuda.stream(side):
        for _ in range(3):
            ops.ideal_loss(L.MODEL_MAGPHA, mp4, a4, ops.gen_tables(te4, 1.5))
        side.synchronize()
        graph4 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph4, stream=side):
            l4, g4, _ = ops.ideal_loss(L.MODEL_MAGPHA, mp4, a4, ops.gen_tables(te4, 1.5))
    add("graph[tables+ideal_loss magpha]", "C4 at nb = 3 (train-IDEAL-single), CUDA-graph replay (latency)", nb4, 384 * 384, ne, 8 * ne + 32 + 32, graph4.replay)
    add("ig_ideal_loss[magpha]", "C4 at nb = 3 through the Python wrappers (latency)", nb4, 384 * 384, ne, 8 * ne + 32 + 32,
        lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp4, a4, ops.gen_tables(te4, 1.5)))
    # C5 consumer: forward model + the three clipped images of gen_LDM_dataset.py:216-237 in one pass (3-channel decoder maps)
    del mp4, a4
    torch.cuda.empty_cache()
    nb, H, W = 64, 384, 384
    nv = H * W
    m3 = torch.from_numpy(synth.magpha_maps(8, H, W, np.random.default_rng(5), bipolar=False)).to(dev).repeat(nb // 8, 1, 1, 1, 1).contiguous()
    tab5 = ops.gen_tables(torch.from_numpy(synth.te_orig(nb, ne)).to(dev), 1.5)
    add("ig_ideal_decode[images]", "C5 consumer: clip |S_e|, PDFF, R2* (no complex output)", nb, nv, ne, 24 + 4 * ne + 8,
        lambda: ops.ideal_decode(L.MODEL_MAGPHA, m3, tab5, ne))
    add("ig_ideal_decode[images+signals]", "C5 consumer + the TFRecord's complex signals", nb, nv, ne, 24 + 4 * ne + 8 + 8 * ne,
        lambda: ops.ideal_decode(L.MODEL_MAGPHA, m3, tab5, ne, want_shat=True))
    del m3
    # more than 8 echoes (train-IDEAL-TEaug.py:614-618 trains with 3 .. 12): forward, LS solve and the C2 objective at ne = 12
    ne12 = 12
    torch.cuda.empty_cache()
    maps, te, tab, acqs = make(nb, H, W, ne12)
    pm = (maps[:, 2:3] * 0.95).contiguous()
    add("ig_ideal_fwd[wfpm]", "ne = 12 forward", nb, nv, ne12, 24 + 8 * ne12, lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne12))
    add("ig_get_rho_fwd", "ne = 12 LS solve", nb, nv, ne12, 8 * ne12 + 8 + 16, lambda: ops.get_rho_fwd(acqs, pm, tab))
    add("ig_a2a_loss", "ne = 12 C2 fused objective (ring, y parked in the stage)", nb, nv, ne12, 8 * ne12 + 8 + 8, lambda: ops.a2a_loss(acqs, pm, tab))
    up12 = torch.randn_like(acqs)
    add("ig_a2a_bwd", "ne = 12 acq_to_acq adjoint (dPM only; generic ring, one block per SM, two 100 KB stages)", nb, nv, ne12, 8 * ne12 + 8 + 8 * ne12 + 8,
        lambda: ops.a2a_bwd(acqs, pm, tab, None, up12, need_acqs=False))
    add("ig_ideal_bwd[wfpm]", "ne = 12 forward-model adjoint (generic ring beyond 8 echoes: RowLossOp<BWD>, one block per SM, three stages)", nb, nv, ne12,
        8 * ne12 + 24 + 24, lambda: ops.ideal_bwd(L.MODEL_WFPM, maps, tab, ne12, up12))
    add("ig_ideal_loss[wfpm]", "ne = 12 fused forward objective (generic ring)", nb, nv, ne12, 8 * ne12 + 24 + 24, lambda: ops.ideal_loss(L.MODEL_WFPM, maps, acqs, tab))
    up_r12 = torch.randn((nb, 2, H, W, 2), device=dev, generator=g)
    add("ig_get_rho_bwd", "ne = 12 LS solve adjoint (dPM + dS; one voxel per thread beyond 8 echoes)", nb, nv, ne12, 16 * ne12 + 8 + 16 + 8,
        lambda: ops.get_rho_bwd(acqs, pm, tab, up_r12, None))
    # an echo count that is not a bucket of the plain kernels: 7 (ring operators are instantiated per echo count)
    del maps, te, tab, acqs, pm, up12, up_r12
    torch.cuda.empty_cache()
    ne7 = 7
    maps, te, tab, acqs = make(nb, H, W, ne7)
    pm = (maps[:, 2:3] * 0.95).contiguous()
    up7 = torch.randn_like(acqs)
    add("ig_a2a_bwd", "ne = 7 acq_to_acq adjoint (dPM only)", nb, nv, ne7, 16 * ne7 + 16, lambda: ops.a2a_bwd(acqs, pm, tab, None, up7, need_acqs=False))
    add("ig_a2a_loss", "ne = 7 C2 fused objective (ring kernels are instantiated per echo count)", nb, nv, ne7, 8 * ne7 + 16, lambda: ops.a2a_loss(acqs, pm, tab))
    print(json.dumps({"hbm_peak_gbs": peak, "device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
