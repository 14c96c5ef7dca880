#!/usr/bin/env python
"""Headline benchmark: voxel-echoes/s of the unsupervised physics loss, forward + backward.

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): acq_to_acq -> where(A != 0) mask -> MSE -> d/dPM on a
batch of 64 slices x 384 x 384 x 6 echoes per GPU, fp32, orig echo times, 1.5 T; synthetic inputs built per
SURVEY §8d (disc-masked random maps -> forward model -> N(0, 0.02^2) noise).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (weak scaling: 64 slices per GPU)
    python bench.py --impl reference          (the reference's op-chain algorithm on the host cores, torch-CPU port)

One JSON line on stdout (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = the same objective through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H per step);
`roofline` = algorithmic bytes of the fused kernel / its measured duration against MEASURED_PEAKS.json.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "ideal-gan_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

NB, H, W, NE = 64, 384, 384, 6
FIELD, R2_SC = 1.5, 200.0
METRIC = "voxel-echoes/sec (fwd+bwd)"
UNIT = "voxel-echoes/s"
WORKLOAD = f"C2 unsupervised physics loss fwd+bwd: acq_to_acq+mask+MSE+dPM, {NB}x{H}x{W}x{NE} echoes per GPU, fp32"
ALGO_BYTES_PER_VOXEL = 8 * NE + 8 + 8          # read ne echoes + PM row, write the PM gradient (64 B)
HBM_FALLBACK_GBS = 6650.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so the real
# stdout is set aside for that line and file descriptor 1 is pointed at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


# ---------------------------------------------------------------------------------------------------------------
# clocks: sample NVML while the timed regions run
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._on = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self._nv = None
            self.error = repr(e)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            else:
                time.sleep(0.0005)

    def __enter__(self):
        self._on.set()
        return self

    def __exit__(self, *exc):
        self._on.clear()

    def close(self):
        self._stop.set()

    def summary(self):
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's algorithm (batched complex op chain + autograd) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_step(acqs, pm, te):
    from oracle import ideal_oracle as orc
    p = pm.clone().requires_grad_(True)
    loss, _, _ = orc.physics_loss_a2a(acqs, p, te=te, field=FIELD, r2_sc=R2_SC)
    (g,) = torch.autograd.grad(loss, [p])
    return loss, g


def cpu_sample(nb, seed=1234):
    """A bounded sample of the C2 workload: nb slices of 384 x 384 x 6 built on the CPU with the oracle."""
    from idealgan import synth
    from oracle import ideal_oracle as orc
    rng = np.random.default_rng(seed)
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = torch.from_numpy(synth.te_orig(nb, NE))
    with torch.no_grad():
        sig = orc.IDEAL_model(torch.from_numpy(maps), [FIELD, te]).numpy()
    acqs = torch.from_numpy(synth.add_noise(sig, rng))
    pm = torch.from_numpy(np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.95))
    return acqs, pm, te


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: as many slices per step (8, 4, 2 or 1) as keep the whole run within ~150 s
    nb = 8
    acqs, pm, te = cpu_sample(nb)
    cpu_port_step(acqs, pm, te)
    c0 = time.perf_counter()
    cpu_port_step(acqs, pm, te)
    per_step = time.perf_counter() - c0
    while nb > 1 and per_step * (args.steps + args.warmup) > 150.0:
        nb //= 2
        per_step /= 2
    acqs, pm, te = acqs[:nb].contiguous(), pm[:nb].contiguous(), te[:nb].contiguous()
    for _ in range(max(args.warmup, 1)):
        cpu_port_step(acqs, pm, te)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(acqs, pm, te)
    dt = time.perf_counter() - t0
    value = nb * H * W * NE * args.steps / dt
    sample = f"{nb} of the {NB} slices per step ({H}x{W}x{NE}), oracle/ideal_oracle.py physics_loss_a2a + torch autograd, complex64"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference is TensorFlow op chains; TF is not installable here, so its "
                       "algorithm is timed as the torch-CPU port held to the reference's vectors (tests/test_oracle_golden.py)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def build_device_inputs(device, seed):
    """Synthetic C2 batch generated on the device with this repo's own forward kernel (SURVEY §8d recipe)."""
    from idealgan import _lib as L
    from idealgan import ops, synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    mask = torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).to(device)[None, None, :, :, None]
    maps = torch.empty((NB, 3, H, W, 2), device=device)
    maps[:, :2] = torch.rand((NB, 2, H, W, 2), device=device, generator=g) - 0.5
    maps[:, 2, :, :, 0] = 2.0 * torch.rand((NB, H, W), device=device, generator=g) - 1.0
    maps[:, 2, :, :, 1] = torch.rand((NB, H, W), device=device, generator=g)
    maps *= mask
    te = torch.from_numpy(synth.te_orig(NB, NE)).to(device)
    tab = ops.gen_tables(te, FIELD)
    clean = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, NE, R2_SC)
    noise = 0.02 * torch.randn(clean.shape, device=device, generator=g)
    acqs = torch.where(clean != 0, clean + noise, torch.zeros_like(clean)).contiguous()
    pm = (maps[:, 2:3] * 0.95).contiguous()          # an imperfect (phi, R2*) estimate, as a network would give
    return acqs, pm, te


def run_ours(args):
    from idealgan import _lib as L
    from idealgan import ops
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    try:                                                  # keep this rank (and the pinned buffers it allocates) next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception as e:                                # not fatal: affinity only matters for the host-buffer leg on multi-socket hosts
        log(f"bench.py: CPU affinity not set ({e!r})")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    lib = L.load()
    assert lib.ig_device_ok() == 1, "libidealgan targets sm_100 (B200)"

    acqs, pm, te = build_device_inputs(device, 1234 + rank)
    nv = H * W
    inv_n = 1.0 / (acqs.numel() * world)                 # mean over the GLOBAL batch; shards sum to it
    stream = torch.cuda.current_stream()
    g_pm = torch.empty((NB, 1, H, W, 2), dtype=torch.float32, device=device)
    from idealgan import dist as igdist
    # the scalar exchange: "peer" = fused into the loss kernel (stores into every rank's mailbox over NVLink, no collective
    # kernel); "nccl" = asynchronous all-reduce of step i under the kernels of step i + 1
    peer, exchange_note = None, None
    if world > 1 and args.exchange == "peer":
        try:
            peer = igdist.PeerLossExchange(device, lag=args.lag)       # collective set-up: fails on every rank or on none
        except RuntimeError as e:
            exchange_note = f"peer exchange unavailable ({e}); NCCL all-reduce used"
            print(exchange_note, file=sys.stderr)
    reducer = igdist.AsyncLossReducer(device, depth=2)
    loss_local = torch.zeros(1, dtype=torch.float32, device=device)
    scratch = ops.loss_scratch(device, NB, nv)
    te2 = te[:, :, 0].contiguous()
    # The per-sample tables depend on the echo times only, which arrive with the batch, ahead of the maps: every step builds
    # its own table, but on a side stream and into one of two buffers, so it runs under the previous step's loss kernel.
    side = torch.cuda.Stream(device)
    tabs = [torch.empty((NB, L.TAB_FLOATS), dtype=torch.float32, device=device) for _ in range(2)]
    tab_ready = [torch.cuda.Event() for _ in range(2)]
    tab_free = [None, None]
    counter = [0]

    def step(ev=None):
        j = counter[0] % 2
        counter[0] += 1
        loss = reducer.acquire()
        if tab_free[j] is not None:
            side.wait_event(tab_free[j])                  # the loss kernel two steps back has finished with this buffer
        L.check(lib.ig_gen_tables(te2.data_ptr(), NB, NE, FIELD, tabs[j].data_ptr(), side.cuda_stream), "ig_gen_tables")
        tab_ready[j].record(side)
        stream.wait_event(tab_ready[j])
        if ev:
            ev[0].record(stream)
        if peer is not None:
            L.check(lib.ig_a2a_loss_peer(acqs.data_ptr(), pm.data_ptr(), nv * 2, tabs[j].data_ptr(), NB, NE, nv, R2_SC, inv_n, g_pm.data_ptr(), 0, 0,
                                         loss_local.data_ptr(), scratch.data_ptr(), scratch.numel(), peer.handle, peer.step, peer.lag, peer.prev.data_ptr(),
                                         stream.cuda_stream), "ig_a2a_loss_peer")
            peer.step += 1
        else:
            L.check(lib.ig_a2a_loss(acqs.data_ptr(), pm.data_ptr(), nv * 2, tabs[j].data_ptr(), NB, NE, nv, R2_SC, inv_n, g_pm.data_ptr(), 0, 0,
                                    loss.data_ptr(), scratch.data_ptr(), scratch.numel(), stream.cuda_stream), "ig_a2a_loss")
        if ev:
            ev[1].record(stream)
        tab_free[j] = torch.cuda.Event()
        tab_free[j].record(stream)
        if peer is None:
            reducer.submit()                              # scalar loss over NVLink: the only exchange on this path

    def fence():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        t0.record(stream)
        side.wait_event(t0)                               # no table of a timed step starts before the opening event
        for i in range(args.steps):
            step(kev[i])
        if peer is not None:
            peer.last(stream.cuda_stream)                 # the last step's global scalar is complete before the closing event
        reducer.drain()                                   # every reduction is ordered before the closing event
        t1.record(stream)
        fence()
    ms = t0.elapsed_time(t1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    final_loss = (peer._last if peer is not None else reducer.last()).item()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------------------
    e2e_s, e2e_steps, e2e_loss, h2d, d2h = 0.0, 0, None, 0, 0
    if not args.no_e2e:
        acqs_h = torch.empty(acqs.shape, dtype=torch.float32).pin_memory()
        pm_h = torch.empty(pm.shape, dtype=torch.float32).pin_memory()
        te_h = te2.cpu().pin_memory()
        acqs_h.copy_(acqs)
        pm_h.copy_(pm)
        g_h = torch.empty(g_pm.shape, dtype=torch.float32).pin_memory()
        l_h = torch.empty(1, dtype=torch.float32).pin_memory()
        h2d, d2h = int((acqs_h.numel() + pm_h.numel() + te_h.numel()) * 4), int((g_h.numel() + 1) * 4)
        ctx = ctypes.c_void_p()
        L.check(lib.ig_ctx_create(local, args.chunk, NE, nv, ctypes.byref(ctx)), "ig_ctx_create")

        def e2e_step():
            L.check(lib.ig_a2a_loss_host(ctx, acqs_h.data_ptr(), pm_h.data_ptr(), te_h.data_ptr(), NB, FIELD, R2_SC, inv_n, l_h.data_ptr(),
                                         g_h.data_ptr()), "ig_a2a_loss_host")

        e2e_steps = max(3, min(args.steps, args.e2e_steps))
        for _ in range(3):
            e2e_step()
        fence()
        with clocks:
            w0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - w0
        lib.ig_ctx_destroy(ctx)
        e2e_loss = l_h.item() * world if world > 1 else l_h.item()
    clocks.close()

    # ---- max over ranks ----------------------------------------------------------------------------------
    times = torch.tensor([ms, kernel_ms, e2e_s * 1e3], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, kernel_ms, e2e_ms = (float(x) for x in times.cpu())
    units_per_step = NB * nv * NE * world
    value = units_per_step * args.steps / (ms * 1e-3)
    e2e_value = units_per_step * e2e_steps / (e2e_ms * 1e-3) if e2e_steps else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
        achieved = ALGO_BYTES_PER_VOXEL * NB * nv / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json"))).get("a2a_loss_bytes_per_launch")
        except Exception:
            pass
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            nbc = 8
            a_c, p_c, t_c = cpu_sample(nbc)
            cpu_port_step(a_c, p_c, t_c)
            best = float("inf")
            reps = 0
            t_budget = time.perf_counter()
            while reps < 5 and (time.perf_counter() - t_budget) < 25.0:
                c0 = time.perf_counter()
                cpu_port_step(a_c, p_c, t_c)
                best = min(best, time.perf_counter() - c0)
                reps += 1
            cpu_baseline = {"value": nbc * nv * NE / best, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{nbc} of the {NB} slices ({H}x{W}x{NE}), best of {reps}: reference algorithm as torch-CPU "
                                      "complex64 op chain + autograd (oracle/ideal_oracle.py); TensorFlow itself is not installable here"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": NB * world, "sharding": f"batch axis, {NB} slices per GPU, "
                       + ("scalar loss exchanged by the loss kernel itself (peer-memory stores over NVLink, no collective kernel)" if peer is not None
                          else "async NCCL all-reduce of the scalar loss only (overlaps the next step)") if world > 1 else "single GPU",
                       "l2": f"inputs {(acqs.numel() + pm.numel()) * 4 / 1e6:.0f} MB per step > 126 MB L2, no flush needed",
                       "step": "ig_gen_tables (side stream, double-buffered) + ig_a2a_loss (fused loss + gradient)" + ((" with the scalar exchange fused in (ig_a2a_loss_peer)" if peer is not None else " + async all_reduce(loss)") if world > 1 else ""),
                       "loss": final_loss, "e2e_loss": e2e_loss, **({"exchange_note": exchange_note} if exchange_note else {})},
            "clocks": clocks.summary(),
            "e2e": None if not e2e_steps else {
                    "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "timer": "host perf_counter around the blocking C-ABI call, device synchronised on both sides",
                    "api": f"ig_a2a_loss_host (3-slot H2D/compute/D2H pipeline, chunks of {args.chunk} slices)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "a2a_loss_tma_kernel<NE=6, MINB=2, STAGES=3, EXACT, CH=8, MODE=1>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_VOXEL * NB * nv,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"},
            "cpu_baseline": cpu_baseline,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--chunk", type=int, default=8, help="slices per chunk of the host pipeline")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer", help="multi-GPU scalar exchange (see idealgan/dist.py)")
    ap.add_argument("--lag", type=int, default=1, help="peer exchange: the global loss a step receives is `lag` steps old")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident leg only (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200:
            args.steps, args.warmup = 5, 1
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
