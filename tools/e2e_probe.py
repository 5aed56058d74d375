"""End-to-end probe of the compact-pose host API: time per call for several chunk sizes, with the library's trace."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
from firecode_b200 import clash, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
a, b = bench.make_fragments()
rng = np.random.default_rng(1)
p7 = clash.pinned_empty((n, 7), np.float32)
q = rng.normal(size=(n, 4)).astype(np.float32)
d = rng.normal(size=(n, 3))
d /= np.linalg.norm(d, axis=1, keepdims=True)
base = synthetic.radius_of_gyration(a) + synthetic.radius_of_gyration(b)
p7[:, :4] = q
p7[:, 4:] = d * rng.uniform(base - 2, base + 4, size=(n, 1))
bits = clash.pinned_empty(((n + 31) // 32,), np.uint32)
for chunk in sys.argv[2:] or ["1048576"]:
    os.environ["FC_CLASH_CHUNK"] = chunk
    for _ in range(3):
        clash.compenetration_check_batch_pose7(a, b, p7, thresh=1.5, bits_out=bits)
    os.environ["FC_CLASH_TRACE"] = "1"
    clash.compenetration_check_batch_pose7(a, b, p7, thresh=1.5, bits_out=bits)
    del os.environ["FC_CLASH_TRACE"]
    t0 = time.perf_counter()
    for _ in range(5):
        r = clash.compenetration_check_batch_pose7(a, b, p7, thresh=1.5, bits_out=bits)
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk}: {dt * 1e3:.2f} ms per call, {n / dt:.3e} poses/s, pass {r.n_pass}", flush=True)
