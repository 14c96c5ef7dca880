#!/usr/bin/env python
"""Headline kernel on a batch small enough to stay L2-resident (8 slices = 66 MB) against the HBM-streaming batch (64):
if time per slice is the same, the kernel is bound by its math/issue, not by HBM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ideal-gan_b200"))
import torch, bench
from idealgan import ops
dev = torch.device("cuda", 0)
acqs, pm, te = bench.build_device_inputs(dev, 1234)
for nb in (64, 32, 16, 8, 4):
    a, p_, t = acqs[:nb].contiguous(), pm[:nb].contiguous(), te[:nb].contiguous()
    tab = ops.gen_tables(t, 1.5)
    for _ in range(10):
        ops.a2a_loss(a, p_, tab)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 200
    ev[0].record()
    for _ in range(n):
        ops.a2a_loss(a, p_, tab)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    print(f"nb={nb:3d} {ms*1e3:8.1f} us/launch  {ms*1e3/nb:6.2f} us/slice  ({(a.numel()+p_.numel())*4/1e6:.0f} MB in)")
