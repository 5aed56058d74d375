"""GPU parity of the conformational-search path (fc_csearch_apply behind torsion.random_csearch /
clustered_csearch) against the oracle port and against structures produced by the UNMODIFIED reference
(tests/golden/csearch_*.npz; firecode/torsion_module.py:436-571, 726-891)."""

import os

import numpy as np
import pytest

from firecode_b200 import torsion
from firecode_b200.utils import cartesian_product
from oracle import make_golden, port
from test_oracle_pinning import GOLDEN

pytestmark = pytest.mark.gpu


def _case(n_atoms, n_tors, seed):
    atoms, coords, g, tors = make_golden.make_csearch_case(n_atoms, n_tors, seed)
    return atoms, coords, g, [make_golden.DuckTorsion(t, nf) for t, nf in tors]


@pytest.mark.parametrize("n_atoms,n_tors,seed", [(30, 4, 5), (45, 5, 8), (24, 3, 21), (60, 4, 33)])
def test_csearch_apply_matches_oracle(gpu, n_atoms, n_tors, seed):
    atoms, coords, g, tors = _case(n_atoms, n_tors, seed)
    tuples = [t.torsion for t in tors]
    masks = [torsion.get_rotation_mask(g, t) for t in tuples]
    angles = cartesian_product(*[t.get_angles() for t in tors])
    starts = np.stack([coords, coords + 0.01 * np.random.default_rng(seed).normal(size=coords.shape)])
    out, rotated, near = torsion.csearch_apply(starts, tuples, masks, angles)
    assert out.shape == (2, len(angles), n_atoms, 3)
    checked = 0
    for s in range(2):
        for k in range(0, len(angles), max(1, len(angles) // 150)):
            x, rot, closest = port.csearch_apply(starts[s], tuples, masks, angles[k])
            if closest <= 1e-6:
                assert near[s, k]
                continue
            assert rot == rotated[s, k], (s, k)
            assert np.abs(x - out[s, k]).max() < 1e-9
            checked += 1
    assert checked > 30 and (rotated > 0).any() and (rotated == 0).any()


@pytest.mark.parametrize("name", sorted(make_golden.CSEARCH_CASES))
def test_csearch_drivers_match_reference_golden(gpu, name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    n_atoms, n_tors, seed = (int(v) for v in z["params"])
    atoms, coords, g, tors = _case(n_atoms, n_tors, seed)
    assert abs(coords.sum() - float(z["checksum"])) < 1e-9
    np.random.seed(seed)
    rnd = torsion.random_csearch(atoms, coords, tors, g, n_out=25, logfunction=None, interactive_print=False)
    assert rnd.shape == z["random"].shape and np.abs(rnd - z["random"]).max() < 1e-9
    np.random.seed(seed)
    clu = torsion.clustered_csearch(atoms, coords, tors, g, n=1000, n_out=100000, logfunction=None,
                                    interactive_print=False)
    assert clu.shape == z["clustered"].shape and np.abs(clu - z["clustered"]).max() < 1e-9
