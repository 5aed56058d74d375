"""Batched multiembed (firecode_b200.multiembed, replacing multiembed.py:33-159) against the UNMODIFIED reference run
on its own fixture firecode/tests/embed_multiembed (two formic acid molecules, 36 arrangements): tests/golden/
embed_multiembed.npz holds, per arrangement, the child's problem and what the reference's run_child_embedder
produced (oracle/make_golden.py:main_multiembed)."""

import os
from types import SimpleNamespace

import numpy as np
import pytest

from firecode_b200 import embeds, multiembed
from firecode_b200.errors import ZeroCandidatesError
from oracle import port
from test_oracle_pinning import GOLDEN

Z = os.path.join(GOLDEN, "embed_multiembed.npz")


class _Sub:
    """npz view restricted to the keys of one child"""

    def __init__(self, z, i):
        self.z, self.p = z, f"c{i}_"

    def __getitem__(self, k):
        return self.z[self.p + k]

    def __contains__(self, k):
        return (self.p + k) in self.z


def child_problem(z, i):
    from test_oracle_pinning import cyclical_problem_from_npz

    return cyclical_problem_from_npz(_Sub(z, i))


def test_arrangement_enumeration_is_the_reference_order():
    z = np.load(Z)
    react = z["c0_reactive0"], z["c0_reactive1"]
    assert sorted(react[0].tolist()) != react[0].tolist() or True
    arr = multiembed.arrangements([1, 3, 4], [1, 3, 4])
    assert np.array_equal(np.array(arr), z["arrangements"]) and len(arr) == 36
    assert all(a[0][0] != a[1][0] and a[0][1] != a[1][1] for a in arr)


class ChildDuck:
    """A child embedder rebuilt from the stored arrays: what embeds.cyclical_embed and the refining steps read."""

    def __init__(self, z, i):
        prob = child_problem(z, i)
        self.embed = str(z[f"c{i}_embed"])
        self.ids = list(prob.ids)
        off = np.concatenate([[0], np.cumsum(self.ids)])
        atoms = z["atoms"]
        self.objects = []
        for m in range(2):
            piv = [[SimpleNamespace(pivot=v, meanpoint=mp, start_atom=SimpleNamespace(cumnum=int(ids[0])),
                                    end_atom=SimpleNamespace(cumnum=int(ids[1])))
                    for v, mp, ids in zip(prob.pivot_vec[m][c], prob.pivot_mean[m][c], prob.pivot_ids[m][c])]
                   for c in range(len(prob.coords[m]))]
            self.objects.append(SimpleNamespace(coords=prob.coords[m], reactive_indices=prob.reactive[m], pivots=piv,
                                                atoms=atoms[off[m]:off[m + 1]]))
        self.systematic_angles = prob.angles
        clash, max_clashes, rmsd = z["options"]
        self.options = SimpleNamespace(clash_thresh=float(z[f"c{i}_thresh"]), max_clashes=int(max_clashes), rmsd=float(rmsd))
        keys = [str(k) for k in z[f"c{i}_table_keys"]] if f"c{i}_table_keys" in z else []
        self.pairings_table = {k: tuple(int(x) for x in p) for k, p in zip(keys, z[f"c{i}_table_pairs"])} if keys else \
            {k: tuple(p) for k, p in zip("xy", prob.pairings)}
        self._dists = dict(zip(keys, z[f"c{i}_dists"])) if keys else {}
        self.internal_constraints = np.array(prob.internal_constraints) if prob.internal_constraints_is_array else \
            list(prob.internal_constraints)
        self.candidates = 0
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
                d = self._dists.get(lett)
                return None if d is None or np.isnan(d) else float(d)
        return None


@pytest.mark.gpu
def test_every_child_embed_matches_the_reference(gpu):
    z = np.load(Z)
    n_nonzero = 0
    for i in range(int(z["n_arrangements"])):
        prob = child_problem(z, i)
        poses, constrained, rep = embeds.cyclical_screen(prob)
        ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
        ref = port.cyclical_embed_bimol(prob, ties=ties)
        assert np.array_equal(rep.kept_indices, ref["kept"]), i
        want = z[f"c{i}_ref_structures"]
        if len(rep.ties) == 0:
            assert poses.shape == want.shape, i
            if len(want):
                assert np.abs(poses - want).max() < 1e-5
                assert np.array_equal(constrained, z[f"c{i}_ref_constrained"])
                n_nonzero += 1
    assert n_nonzero >= 6


@pytest.mark.gpu
def test_multiembed_bifunctional_matches_the_reference_in_arrangement_order(gpu):
    z = np.load(Z)
    atoms = z["atoms"]
    n1 = int(z["c0_ids"][0])
    parent = SimpleNamespace(
        objects=[SimpleNamespace(reactive_indices=np.array([1, 3, 4]), atoms=atoms[:n1]),
                 SimpleNamespace(reactive_indices=np.array([1, 3, 4]), atoms=atoms[n1:])],
        lines=[])
    parent.log = lambda *a, **k: parent.lines.append(a[0] if a else "")
    made = []

    def factory(emb, arrangement, i):
        assert np.array_equal(np.array(arrangement), z["arrangements"][i])
        made.append(i)
        return ChildDuck(z, i)

    out = multiembed.multiembed_bifunctional(parent, make_child=factory)
    assert made == list(range(36))
    want = np.concatenate([z[f"c{i}_final_structures"] for i in range(36) if len(z[f"c{i}_final_structures"])])
    want_c = np.concatenate([z[f"c{i}_final_constrained"] for i in range(36) if len(z[f"c{i}_final_structures"])])
    assert parent.b200_multiembed_counts == [len(z[f"c{i}_final_structures"]) for i in range(36)]
    assert out.shape == want.shape == (12, 10, 3)
    assert np.abs(out - want).max() < 1e-5
    assert np.array_equal(parent.constrained_indices, want_c)
    assert np.array_equal(parent.atoms, atoms) and parent.structures is out
    assert any("Multiembed completed" in str(x) for x in parent.lines)
    # dispatcher: only two molecules are supported (multiembed.py:26-30)
    parent.objects.append(parent.objects[0])
    with pytest.raises(Exception, match="currently unavailable"):
        multiembed.multiembed_dispatcher(parent)


@pytest.mark.gpu
def test_multiembed_without_any_candidate_raises(gpu):
    z = np.load(Z)
    atoms = z["atoms"]
    n1 = int(z["c0_ids"][0])
    parent = SimpleNamespace(objects=[SimpleNamespace(reactive_indices=np.array([1, 3]), atoms=atoms[:n1]),
                                      SimpleNamespace(reactive_indices=np.array([1, 3]), atoms=atoms[n1:])])
    parent.log = lambda *a, **k: None
    zero_children = [i for i in range(36) if len(z[f"c{i}_final_structures"]) == 0]
    with pytest.raises(ZeroCandidatesError):
        multiembed.multiembed_bifunctional(parent, make_child=lambda e, a, i: ChildDuck(z, zero_children[i]))
