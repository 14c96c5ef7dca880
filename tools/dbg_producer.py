import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/ideal-gan_b200")
import torch, bench
from idealgan import _lib as L, ops
dev = torch.device("cuda", 0)
acqs, pm, te = bench.build_device_inputs(dev, 1234)
tab = ops.gen_tables(te, 1.5)
for i in range(3):
    os.environ.pop("IG_A2A_DEBUG", None)
    ops.a2a_loss(acqs, pm, tab)
torch.cuda.synchronize()
os.environ["IG_A2A_DEBUG"] = "1"
ops.a2a_loss(acqs, pm, tab)
torch.cuda.synchronize()
