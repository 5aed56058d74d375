"""GPU parity of the refining steps that follow the embed (firecode_b200.refining) against the kept sets the
UNMODIFIED reference methods produced on the same duck-typed embedder state (tests/golden/refining_*.npz, made by
oracle/make_golden.py:main_refining from /root/reference/firecode/embedder.py:1410-1514, 1954-2039)."""

import os
from types import SimpleNamespace

import numpy as np
import pytest

from firecode_b200 import refining
from firecode_b200.errors import ZeroCandidatesError
from oracle import make_golden

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Duck:
    def __init__(self, name):
        n, n_frag, seed = make_golden.REFINING_CASES[name]
        atoms, structures, ids, constrained, table, dists = make_golden.make_refining_case(n, n_frag, seed)
        self.gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        assert float(self.gold["checksum"]) == float(structures.sum())
        self.n = n
        self.embed, self.ids, self.atoms = "multiembed", ids, atoms
        self.options = SimpleNamespace(clash_thresh=1.5, max_clashes=2, rmsd=0.5)
        self.structures, self.constrained_indices = structures, constrained
        self.energies, self.exit_status = np.arange(n, dtype=float), np.zeros(n, dtype=bool)
        self.pairings_table, self.objects, self._dists = table, [None] * n_frag, dists
        self.lines = []

    def log(self, *a, **k):
        self.lines.append(a[0] if a else "")

    def debuglog(self, *a, **k):
        pass

    def log_warnings(self):
        pass

    def get_pairing_dists_from_constrained_indices(self, pair):      # embedder.py:1627-1642
        for lett, p in self.pairings_table.items():
            if p[0] == pair[0] and p[1] == pair[1]:
                return self._dists[lett]
        return None


@pytest.mark.parametrize("name", sorted(make_golden.REFINING_CASES))
def test_compenetration_refining_matches_reference(gpu, name):
    d = Duck(name)
    tag = np.arange(d.n)
    d.constrained_indices = np.concatenate([d.constrained_indices, np.broadcast_to(tag[:, None, None], (d.n, 1, 2))], axis=1)
    refining.compenetration_refining(d)
    assert np.array_equal(d.constrained_indices[:, -1, 0], d.gold["kept_compenetration"])
    assert len(d.structures) == len(d.gold["kept_compenetration"]) and 0 < len(d.structures) < d.n
    assert d.energies.shape == (len(d.structures),) and np.all(d.energies == 1e10) and not d.exit_status.any()
    assert any("for compenetration" in str(x) for x in d.lines)


@pytest.mark.parametrize("name", sorted(make_golden.REFINING_CASES))
def test_fitness_refining_matches_reference(gpu, name):
    d = Duck(name)
    refining.fitness_refining(d, threshold=3.0)
    assert len(d.b200_fitness_near) == 0          # no structure within 1e-6 of the threshold in these fixtures
    assert np.array_equal(d.energies.astype(np.int64), d.gold["kept_fitness"])
    assert len(d.structures) == len(d.constrained_indices) == len(d.exit_status) == len(d.gold["kept_fitness"])


@pytest.mark.parametrize("name", sorted(make_golden.REFINING_CASES))
def test_similarity_refining_matches_reference(gpu, name):
    d = Duck(name)
    refining.similarity_refining(d, tfd=False, moi=True, rmsd=True)
    assert np.array_equal(d.energies.astype(np.int64), d.gold["kept_similarity"])
    assert len(d.structures) == len(d.constrained_indices) == len(d.gold["kept_similarity"])


def test_fitness_check_single_and_zero_candidates(gpu):
    d = Duck("refining_a")
    x = d.structures[0]
    pairs = d.constrained_indices[0]
    targets = [d.get_pairing_dists_from_constrained_indices(p) for p in pairs]
    err = sum(np.linalg.norm(x[a] - x[b]) - t for (a, b), t in zip(pairs, targets) if t is not None)
    assert refining.fitness_check(x, pairs, targets, err + 1e-3) is True
    assert refining.fitness_check(x, pairs, targets, err - 1e-3) is False
    mask, e = refining.fitness_check_batch(d.structures[:5], d.constrained_indices[:5], [targets] * 5, 3.0, return_errors=True)
    assert abs(e[0] - err) < 1e-12
    with pytest.raises(ZeroCandidatesError):
        refining.fitness_refining(d, threshold=-1e9)
