"""C2 (BASELINE.json configs[1]): trimolecular cyclical embed of three 50-conformer ensembles of 60 atoms.
usage: python tools/run_c2.py [n_conf_tuples] [n_orb]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from firecode_b200 import embeds, problem, synthetic
from firecode_b200.synthetic_embedder import make_embedder

n_tuples = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_orb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.perf_counter()
emb = make_embedder("cyclical", 50, 60, seed=synthetic.SEED + 2, n_mols=3, n_reactive=2, n_orb=n_orb)
prob = problem.cyclical_problem(emb)
print("setup %.2fs" % (time.perf_counter() - t0), flush=True)
for rep_i in range(2):
    t0 = time.perf_counter()
    poses, cons, rep = embeds.cyclical3_screen(prob, conf_tuple_range=(0, n_tuples), want_status=False, want_coords=False)
    dt = time.perf_counter() - t0
    print(f"tuples={n_tuples} groups={len(rep.group_choice)} poses={rep.n_poses} pass={rep.n_clash_pass} kept={rep.n_kept} "
          f"rechecks={rep.n_fp64_rechecks} ties={rep.n_ties_total} {dt:.3f}s {rep.n_poses/dt:.3e} poses/s mingap={rep.group_gap.min():.2e}", flush=True)
