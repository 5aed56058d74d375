"""Cross-checks of the fast paths against the exact paths at sizes the CPU oracle cannot reach:
  * cell-list clash screen vs all-pairs kernel (identical status bytes) over many geometries / thresholds;
  * two-stage pruning (FP32 screen + FP64 exact) vs the FP64-only pair kernel (identical masks, all conventions);
  * trimolecular embed: conformer-triple slices vs one call.
usage: python tools/stress_crosscheck.py [n_rounds]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from firecode_b200 import clash, embeds, problem, pruner, synthetic
from firecode_b200.synthetic_embedder import make_embedder

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(99)
t0 = time.time()
n_clash = n_prune = 0
for rd in range(rounds):
    n_a, n_b = int(rng.integers(5, 254)), int(rng.integers(5, 300))
    thr = float(rng.choice([0.6, 1.0, 1.5, 2.2, 3.0]))
    n_conf_a, n_conf_b = int(rng.integers(1, 6)), int(rng.integers(1, 4))
    _, ca, _, _ = synthetic.conformer_ensemble(rng, n_conf_a, n_a, n_torsions=3)
    _, cb, _, _ = synthetic.conformer_ensemble(rng, n_conf_b, n_b, n_torsions=3)
    n = 400_000
    xf = synthetic.sweep_poses(rng, ca[0], cb[0], n)
    conf_a = np.sort(rng.integers(0, n_conf_a, size=n))
    conf_b = rng.integers(0, n_conf_b, size=n)
    order = np.lexsort((conf_b, conf_a))
    conf_a, conf_b = conf_a[order], conf_b[order]
    for mc in (0, 3):
        res = {}
        for mode in ("0", "1"):
            os.environ["FC_CLASH_MODE"] = mode
            res[mode] = clash.compenetration_check_batch(ca, cb, xf, thresh=thr, max_clashes=mc, conf_a=conf_a, conf_b=conf_b)
        same = np.array_equal(res["0"].status & 1, res["1"].status & 1)
        assert same, (rd, n_a, n_b, thr, mc, int(((res["0"].status ^ res["1"].status) & 1).sum()))
        n_clash += n
    os.environ.pop("FC_CLASH_MODE", None)
    print(f"round {rd}: clash {n_a}x{n_b} thr {thr} confs {n_conf_a}x{n_conf_b}: pass {res['1'].mask.mean():.3f} OK", flush=True)

    n_s, n_at, nb = int(rng.integers(2000, 30000)), int(rng.integers(10, 80)), int(rng.integers(20, 800))
    atoms, structures, _ = synthetic.pruning_ensemble(rng, n_s, n_at, nb, jitter=(0.02, 0.45))
    energies = rng.uniform(0, 3, size=n_s) if rd % 2 else None
    for keep, pm in (("first", "greedy"), ("last", "snapshot")):
        masks = {}
        for env in ("0", "1"):
            os.environ["FC_PRUNE_FP64"] = env
            _, masks[env] = pruner.prune_by_rmsd(structures, atoms, 0.45, energies=energies, max_dE=1.0 if energies is not None else 0.0,
                                                 keep=keep, pass_mode=pm)
        assert np.array_equal(masks["0"], masks["1"]), (rd, n_s, n_at, keep, pm)
        n_prune += pruner.last_report.pairs_tiled
    os.environ.pop("FC_PRUNE_FP64", None)
    print(f"round {rd}: prune {n_s}x{n_at} basins {nb}: kept {int(masks['0'].sum())} OK", flush=True)

emb = make_embedder("cyclical", [4, 3, 5], 30, seed=77, n_mols=3, n_reactive=2, n_orb=1)
prob = problem.cyclical_problem(emb)
poses, cons, rep = embeds.cyclical3_screen(prob)
kept, base = [], 0
for lo, hi in ((0, 7), (7, 31), (31, 60)):
    p_, c_, r_ = embeds.cyclical3_screen(prob, conf_tuple_range=(lo, hi))
    kept.append(r_.kept_indices + base)
    base += r_.n_poses
assert np.array_equal(np.concatenate(kept), rep.kept_indices)
print(f"stress OK: {n_clash} poses x2 paths, {n_prune} pruning pairs x2 paths, trimolecular slices ({rep.n_poses} poses) in {time.time() - t0:.0f} s")
