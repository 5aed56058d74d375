"""GPU parity of the batched torsion rotation + clash filter (C-ABI fc_torsion_scan)."""

import numpy as np
import pytest
from networkx import Graph

from firecode_b200 import synthetic, torsion
from oracle import port

pytestmark = pytest.mark.gpu


def _setup(seed, n_conf, n_atoms, n_tors):
    rng = np.random.default_rng(seed)
    atoms, coords, bonds, picks = synthetic.conformer_ensemble(rng, n_conf, n_atoms, n_torsions=n_tors)
    g = Graph()
    g.add_nodes_from(range(n_atoms))
    g.add_edges_from(bonds)
    torsions = []
    for p, ch in picks:
        nb_p = [k for k in g.neighbors(p) if k != ch]
        nb_c = [k for k in g.neighbors(ch) if k != p]
        if nb_p and nb_c:
            torsions.append((nb_p[0], p, ch, nb_c[0]))
    masks = [torsion.get_rotation_mask(g, t) for t in torsions]
    for t, m in zip(torsions, masks):
        assert np.array_equal(m, port.rotation_mask(g, t))
    return coords, np.array(torsions), np.array(masks)


@pytest.mark.parametrize("seed,n_conf,n_atoms,n_tors,n_ang", [(3, 4, 40, 4, 36), (8, 2, 120, 8, 12), (5, 1, 12, 2, 7)])
def test_torsion_scan_matches_oracle(gpu, seed, n_conf, n_atoms, n_tors, n_ang):
    coords, torsions, masks = _setup(seed, n_conf, n_atoms, n_tors)
    assert len(torsions) > 0
    angles = np.arange(n_ang) * (360.0 / n_ang)
    res = torsion.torsion_scan(coords, torsions, masks, angles, thresh=1.5, want_min_dist=True)
    ref_xyz, ref_pass, ref_dmin = port.torsion_scan(coords, torsions, masks, angles, thresh=1.5)
    assert np.abs(res["coords"] - ref_xyz).max() < 1e-9
    ok = np.abs(ref_dmin - 1.5) > 1e-6
    assert np.array_equal(res["passed"][ok], ref_pass[ok])
    assert np.all(res["near"][~ok])
    assert np.allclose(res["min_dist"], ref_dmin, atol=1e-9)
    assert 0 < res["passed"].mean() <= 1


def test_single_call_api(gpu):
    coords, torsions, masks = _setup(11, 1, 30, 3)
    new = torsion.rotate_dihedral(coords[0], torsions[0], 37.0, mask=masks[0])
    ref_xyz, ref_pass, _ = port.torsion_scan(coords[:1], torsions[:1], masks[:1], [37.0])
    assert np.abs(new - ref_xyz[0, 0, 0]).max() < 1e-9
    assert torsion.torsion_comp_check(new, torsions[0], masks[0], thresh=1.5) == bool(ref_pass[0, 0, 0])
    assert np.array_equal(coords[0][~masks[0]], new[~masks[0]])  # static atoms untouched
