"""GPU parity of the TFD ensemble pruning (fc_tfd_fingerprints / fc_tfd_first_match behind
torsion.prune_conformers_tfd) against the oracle port and against masks produced by the UNMODIFIED
reference (tests/golden/tfd_prune_*.npz, firecode/torsion_module.py:957-1043)."""

import os

import numpy as np
import pytest

from firecode_b200 import torsion
from oracle import make_golden, port
from test_oracle_pinning import GOLDEN

pytestmark = pytest.mark.gpu


def _forced(ties):
    return {("tfd", int(t["a"]), int(t["b"])): bool(t["decision"]) for t in (ties if ties is not None else [])}


@pytest.mark.parametrize("name", ["tfd_prune_a", "tfd_prune_b"])
def test_tfd_prune_matches_reference_golden(gpu, name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    structures, quads = make_golden.make_tfd_case(*[int(v) for v in z["params"]])
    assert abs(structures.sum() - float(z["checksum"])) < 1e-9 and np.array_equal(quads, z["quadruplets"])
    kept, mask = torsion.prune_conformers_tfd(structures, quads)
    assert mask.dtype == bool and kept.shape == (int(mask.sum()),) + structures.shape[1:]
    if len(torsion.last_tfd_ties) == 0:
        assert np.array_equal(mask, z["ref_mask"])
    else:  # near-threshold sums: compare through the port conditioned on the listed decisions
        ties = port.Ties(eps=1e-6, forced=_forced(torsion.last_tfd_ties))
        assert np.array_equal(mask, port.prune_conformers_tfd(structures, quads, ties=ties)[1])
    assert 1 < mask.sum() < len(mask)


@pytest.mark.parametrize("n,n_atoms,n_basins,seed,thresh", [(900, 24, 60, 11, 10), (60, 12, 3, 12, 25), (3000, 14, 40, 13, 10),
                                                          (5, 10, 1, 14, 10), (1, 10, 1, 15, 10)])
def test_tfd_prune_matches_oracle(gpu, n, n_atoms, n_basins, seed, thresh):
    structures, quads = make_golden.make_tfd_case(n, n_atoms, n_basins, seed)
    tf = torsion._get_tf_mat(structures, quads)
    assert np.abs(tf - port.tf_mat(structures, quads)).max() < 1e-9
    kept, mask = torsion.prune_conformers_tfd(structures, quads, thresh=thresh)
    ties = port.Ties(eps=1e-6, forced=_forced(torsion.last_tfd_ties))
    _, ref_mask = port.prune_conformers_tfd(structures, quads, thresh=thresh, ties=ties)
    assert not [k for k in ties.seen if k not in ties.forced]
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(kept, structures[ref_mask])


def test_tfd_similarity_and_fingerprint_helpers(gpu):
    structures, quads = make_golden.make_tfd_case(4, 12, 2, 3)
    fp0 = torsion.get_torsion_fingerprint(structures[0], quads)
    assert np.abs(fp0 - port.torsion_fingerprint(structures[0], quads)).max() < 1e-9
    for i in range(1, 4):
        fpi = port.torsion_fingerprint(structures[i], quads)
        for thr in (5, 60, 400):
            assert torsion.tfd_similarity(fp0, fpi, thresh=thr) == (port.tfd_sum(fp0, fpi) < thr)
