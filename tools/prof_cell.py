"""Profiling driver for the cell-list clash screen: N poses of two 150-atom fragments, a few steps, nothing else.
    python tools/prof_cell.py [n_poses] [steps] [q7]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
from firecode_b200 import clash, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
q7 = len(sys.argv) > 3 and sys.argv[3] == "q7"
dev = torch.device("cuda", 0)
a, b = bench.make_fragments()
a_dev = torch.from_numpy(a).to(dev)[None].contiguous()
b_dev = torch.from_numpy(b).to(dev)[None].contiguous()
rng = np.random.default_rng(synthetic.SEED + 5)
qq = rng.normal(size=(n, 4))
d = rng.normal(size=(n, 3))
d /= np.linalg.norm(d, axis=1, keepdims=True)
base = synthetic.radius_of_gyration(a) + synthetic.radius_of_gyration(b)
t = d * rng.uniform(base - 2.0, base + 4.0, size=(n, 1))
pose7 = clash.pack_poses7(qq, t)
from oracle import port

xf = torch.from_numpy(port.pose7_to_xf(pose7)).to(dev)
p7 = torch.from_numpy(pose7).to(dev)
bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for s in range(steps):
    if s == steps - 1:
        ev0.record()
    prep = clash.DevicePrep(a_dev, 1.5)
    clash.screen_device_ex(prep, b_dev, p7 if q7 else xf, clash.POSE_Q7 if q7 else clash.POSE_XF64, bits_out=bits)
    prep.free()
ev1.record()
torch.cuda.synchronize()
npass = int(np.unpackbits(bits.cpu().numpy().view(np.uint8)).sum())
print(f"poses {n} pass {npass} last step {ev0.elapsed_time(ev1):.3f} ms")
