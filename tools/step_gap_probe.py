"""Where the 2-3 us between consecutive loss kernels of bench.py's step go: the same step with and without the per-kernel timing events,
with the table on the side stream / on the main stream / built once.  Builder's probe, prints one line per variant."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ideal-gan_b200"))
import bench as B                                          # noqa: E402
from idealgan import _lib as L                             # noqa: E402
from idealgan import ops                                   # noqa: E402


def main():
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = L.load()
    acqs, pm, te, _ = B.build_device_inputs(device, 1234)
    NB, NE, H, W = B.NB, B.NE, B.H, B.W
    nv = H * W
    inv_n = 1.0 / acqs.numel()
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device)
    g_pm = torch.empty((NB, 1, H, W, 2), dtype=torch.float32, device=device)
    loss = torch.zeros(1, dtype=torch.float32, device=device)
    scratch = ops.loss_scratch(device, NB, nv)
    te2 = te[:, :, 0].contiguous()
    tabs = [torch.empty((NB, L.TAB_FLOATS), dtype=torch.float32, device=device) for _ in range(3)]
    ready = [torch.cuda.Event() for _ in range(3)]
    free = [None] * 3
    cnt = [0]

    def tables(j, st):
        L.check(lib.ig_gen_tables(te2.data_ptr(), NB, NE, B.FIELD, tabs[j].data_ptr(), st), "ig_gen_tables")

    def loss_k(j):
        L.check(lib.ig_a2a_loss(acqs.data_ptr(), pm.data_ptr(), nv * 2, tabs[j].data_ptr(), NB, NE, nv, B.R2_SC, inv_n, g_pm.data_ptr(), 0, 0,
                                loss.data_ptr(), scratch.data_ptr(), scratch.numel(), stream.cuda_stream), "ig_a2a_loss")

    def step_side(ev):
        j = cnt[0] % 2
        cnt[0] += 1
        if free[j] is not None:
            side.wait_event(free[j])
        tables(j, side.cuda_stream)
        ready[j].record(side)
        stream.wait_event(ready[j])
        if ev:
            ev[0].record(stream)
        loss_k(j)
        if ev:
            ev[1].record(stream)
        free[j] = torch.cuda.Event()
        free[j].record(stream)

    def step_chain(ev):
        i = cnt[0]
        cnt[0] += 1
        L.check(lib.ig_gen_tables_ahead(te2.data_ptr(), NB, NE, B.FIELD, tabs[(i + 1) % 3].data_ptr(), stream.cuda_stream), "ig_gen_tables_ahead")
        loss_k(i % 3)

    def step_inline(ev):
        tables(0, stream.cuda_stream)
        loss_k(0)

    def step_static(ev):
        loss_k(0)

    def once(step, events, steps=100):
        time.sleep(0.4)                                     # let the board drop out of its power cap: every variant starts from the same state
        for _ in range(50):
            step(None)
        torch.cuda.synchronize()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if events else None for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        side.wait_event(t0)
        for i in range(steps):
            step(kev[i])
        t1.record(stream)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / steps

    for j in range(3):
        tables(j, stream.cuda_stream)
    torch.cuda.synchronize()
    variants = [("side-stream tables, per-kernel events (bench.py today)", step_side, True),
                ("side-stream tables, no per-kernel events", step_side, False),
                ("table of step i+1 ahead of objective i, one stream (PDL chain)", step_chain, False),
                ("tables on the main stream (serial)", step_inline, False),
                ("table built once (loss kernel back to back)", step_static, False)]
    res = {v[0]: [] for v in variants}
    for _ in range(8):                                      # round-robin: drift hits every variant alike
        for name, step, events in variants:
            res[name].append(once(step, events))
    ref = loss.item()
    step_chain(None)
    torch.cuda.synchronize()
    assert loss.item() == ref or abs(loss.item() - ref) < 1e-6 * abs(ref), (loss.item(), ref)
    for name, xs in res.items():
        print(f"{name:64s} ms/step min {min(xs):.5f} median {np.median(xs):.5f} max {max(xs):.5f}", flush=True)


if __name__ == "__main__":
    main()
