"""C5 (BASELINE.json configs[4]): 36-step torsion scan over 8 rotatable bonds of a 1 000-conformer ensemble."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, networkx as nx
from firecode_b200 import synthetic, torsion
rng = np.random.default_rng(synthetic.SEED + 5)
atoms, coords, bonds, picks = synthetic.conformer_ensemble(rng, 1000, 120, n_torsions=8)
g = nx.Graph(); g.add_nodes_from(range(120)); g.add_edges_from(bonds)
tors = []
for p, ch in picks:
    nb_p = [k for k in g.neighbors(p) if k != ch]; nb_c = [k for k in g.neighbors(ch) if k != p]
    if nb_p and nb_c:
        tors.append((nb_p[0], p, ch, nb_c[0]))
masks = [torsion.get_rotation_mask(g, t) for t in tors]
angles = np.arange(36) * 10.0
for i in range(3):
    t = time.perf_counter(); res = torsion.torsion_scan(coords, tors, masks, angles, thresh=1.5, want_coords=False); dt = time.perf_counter() - t
    n = len(coords) * len(tors) * len(angles)
    print(f"C5: {n} structures, pass {res['passed'].mean():.4f}, {dt*1e3:.2f} ms, {n/dt:.3e} structures/s", flush=True)
