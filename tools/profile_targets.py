#!/usr/bin/env python
"""Launch each kernel of interest IG_PROFILE_REPS times (default 2: warm-up + one) at the C2 shape, for `ncu --set full -k regex:ig::` captures:
python tools/profile_targets.py   (see profiles/ncu_kernels_r01.md for the summary of such a run)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ideal-gan_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from idealgan import _lib as L  # noqa: E402
from idealgan import ops  # noqa: E402

dev = torch.device("cuda", 0)
acqs, pm, te, _ = bench.build_device_inputs(dev, 1234)
nb, ne, H, W, _ = acqs.shape
tab = ops.gen_tables(te, 1.5)
g = torch.Generator(device=dev)
g.manual_seed(1)
up = torch.randn(acqs.shape, device=dev, generator=g)
up_rho = torch.randn((nb, 2, H, W, 2), device=dev, generator=g)
pv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 4e-3
rv = torch.rand((nb, 1, H, W, 1), device=dev, generator=g) * 3e-3
rm = pm[..., 1:2].contiguous()
maps = torch.cat([up_rho * 0.1, pm], dim=1).contiguous()
mp = torch.rand((nb, 2, H, W, 4), device=dev, generator=g) * 0.5
mag = torch.sqrt((acqs ** 2).sum(-1, keepdim=True)).contiguous()
_, _, demod_s, ls_s, _ = ops.cse_mag_fwd(mag, rm.reshape(nb, 1, H, W, 1), tab)
var5 = torch.rand((nb, 5, H, W, 2), device=dev, generator=g) * 1e-3
mp3 = mp[..., :3].contiguous()
ups5 = [torch.randn((nb, c_, H, W, 1), device=dev, generator=g) for c_ in (2, ne, ne, 3, 1)]
phm = pm[..., 0:1].contiguous()
targets = [
    lambda: ops.a2a_loss(acqs, pm, tab),
    lambda: ops.a2a_loss(acqs, pm, tab, want_rho=True, want_shat=True),
    lambda: ops.a2a_uq_loss(acqs, pm, pv, rm, rv, tab),
    lambda: ops.a2a_rician_loss(acqs, pm, pv, rm, rv, tab),
    lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=True),
    lambda: ops.get_rho_bwd(acqs, pm, tab, up_rho, None),
    lambda: ops.ideal_loss(L.MODEL_WFPM, maps, acqs, tab),
    lambda: ops.ideal_loss(L.MODEL_MAGPHA, mp, acqs, tab),
    lambda: ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne),
    lambda: ops.get_rho_fwd(acqs, pm, tab),
    lambda: ops.mag_regs(ls_s, demod_s, rm.reshape(nb, 1, H, W, 1), (0.1, 0.2, 0.3, 0.4)),
    lambda: ops.roi_maps(maps, var5, "PDFF-var"),
    lambda: ops.ideal_decode(L.MODEL_MAGPHA, mp3, tab, ne),
    lambda: ops.ideal_decode(L.MODEL_MAGPHA, mp3, tab, ne, want_shat=True),
    lambda: ops.pdff_extract(up_rho),
    lambda: ops.pdff_unc(acqs, phm, pv, rm, rv, tab),
    lambda: ops.cse_mag_fwd(mag, rm.reshape(nb, 1, H, W, 1), tab),
    lambda: ops.cse_mag_bwd(mag, rm.reshape(nb, 1, H, W, 1), tab, ups5),
    lambda: ops.acq_unc_fwd(up_rho, pv, rm, rv, tab, ne),
    lambda: ops.acq_unc_bwd(up_rho, pv, rm, rv, tab, ne, up),
    lambda: ops.a2a_fwd(acqs, pm, tab),
    lambda: ops.a2a_bwd(acqs, pm, tab, None, up, need_acqs=False),
]
reps = int(os.environ.get("IG_PROFILE_REPS", "2"))
for fn in targets:
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
print("done")
