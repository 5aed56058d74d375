"""C4 (BASELINE.json configs[3]): RMSD pruning of a synthetic ensemble of a 120-atom molecule.
usage: python tools/run_c4.py [n] [n_basins]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from firecode_b200 import pruner, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else n // 100
rng = np.random.default_rng(synthetic.SEED + 4)
t0 = time.perf_counter()
atoms, structures, _ = synthetic.pruning_ensemble(rng, n, 120, nb)
print("gen %.2fs" % (time.perf_counter() - t0), flush=True)
for i in range(2):
    t0 = time.perf_counter()
    kept, mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
    dt = time.perf_counter() - t0
    rep = pruner.last_report
    print(f"n={n} kept={int(mask.sum())} passes={rep.passes} pairs={rep.pairs_tiled:.3e} solved={rep.pairs_solved:.3e} "
          f"{dt:.3f}s {rep.pairs_tiled/dt:.3e} pairs/s ties={rep.n_ties_total}", flush=True)
