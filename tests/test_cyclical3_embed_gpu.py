"""GPU parity of the trimolecular cyclical embed (C-ABI fc_cyclical3_screen) against the oracle port
and against outputs of the UNMODIFIED reference stored in tests/golden/synth_trimol_*.npz."""

import os

import numpy as np
import pytest

from firecode_b200 import embeds, problem
from firecode_b200.errors import ZeroCandidatesError
from oracle import port
from synth_embedder import make_embedder
from test_oracle_pinning import GOLDEN, cyclical_problem_from_npz

pytestmark = pytest.mark.gpu

CHOICE_EPS = 1e-9  # degrees: grid-search candidates closer than this are listed, not trusted


def _check_problem(prob, poses, constrained, rep):
    forced_choice = {int(g): int(rep.group_choice[g]) for g in np.flatnonzero(rep.group_gap <= CHOICE_EPS)}
    ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
    ref = port.cyclical_embed_trimol(prob, ties=ties, forced_choice=forced_choice, choice_eps=CHOICE_EPS)
    missing = [k for k in ref["ties"].seen if k not in ties.forced]
    assert not missing, missing[:5]
    assert len(ref["groups"]) == len(rep.group_choice)
    # the stateful direction search picked the same candidate in every group
    assert [g["choice"] for g in ref["groups"]] == rep.group_choice.tolist()
    assert np.allclose([g["gap"] for g in ref["groups"]], rep.group_gap, atol=1e-7)
    assert rep.n_poses == len(ref["clash_pass"])
    assert np.array_equal((rep.status & 1).astype(bool), ref["clash_pass"])
    assert np.array_equal(rep.kept_indices, ref["kept"])
    assert poses.shape == ref["poses"].shape
    if len(poses):
        assert np.abs(poses - ref["poses"]).max() < 1e-5
    assert np.array_equal(np.asarray(constrained).reshape(-1, 3, 2), ref["constrained"])
    return ref


@pytest.mark.parametrize("kw", [
    dict(n_conf=2, n_atoms=12, seed=3, n_reactive=2, n_orb=1),
    dict(n_conf=[2, 1, 2], n_atoms=[20, 16, 24], seed=9, n_reactive=1, n_orb=2),
    dict(n_conf=[1, 2, 1], n_atoms=[33, 30, 40], seed=31, n_reactive=2, n_orb=1),
])
def test_trimolecular_embed_matches_oracle(gpu, kw):
    emb = make_embedder("cyclical", n_mols=3, **kw)
    prob = problem.cyclical_problem(emb)
    n_tot = sum(c.shape[1] for c in prob.coords)
    try:
        emb.b200_want_status = True
        poses = embeds.cyclical_embed(emb)
        constrained = emb.constrained_indices
    except ZeroCandidatesError:
        poses, constrained = np.zeros((0, n_tot, 3)), np.zeros((0, 3, 2), dtype=int)
    rep = emb.b200_report
    assert rep.n_poses > 0
    _check_problem(prob, poses, constrained, rep)


@pytest.mark.parametrize("name", ["synth_trimol_a", "synth_trimol_b"])
def test_trimolecular_embed_matches_reference_golden(gpu, name):
    """CUDA path against what the unmodified reference produced for the same problem."""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    prob = cyclical_problem_from_npz(z)
    poses, constrained, rep = embeds.cyclical3_screen(prob)
    assert rep.group_gap.min() > CHOICE_EPS
    # near-threshold decisions would make the comparison conditional; these fixtures have none
    if len(rep.ties) == 0:
        assert poses.shape == z["ref_structures"].shape
        assert np.abs(poses - z["ref_structures"]).max() < 1e-5
        assert np.array_equal(constrained, z["ref_constrained"])
    else:
        _check_problem(prob, poses, constrained, rep)


def test_trimolecular_reference_fixture_zero_candidates(gpu):
    """firecode/tests/embed_trimolecular: every orientation fails the pairing filter (quirk N10)."""
    z = np.load(os.path.join(GOLDEN, "embed_trimolecular.npz"))
    prob = cyclical_problem_from_npz(z)
    poses, constrained, rep = embeds.cyclical3_screen(prob)
    assert rep.n_poses == 0 and len(poses) == 0


def test_trimolecular_conf_tuple_slices_concatenate(gpu):
    """Sharding by conformer-triple ranges (multi-GPU path) reproduces the single call."""
    emb = make_embedder("cyclical", n_mols=3, n_conf=[2, 2, 1], n_atoms=14, seed=5, n_reactive=2, n_orb=1)
    prob = problem.cyclical_problem(emb)
    poses, constrained, rep = embeds.cyclical3_screen(prob)
    parts, kept, base = [], [], 0
    for lo, hi in ((0, 1), (1, 3), (3, 4)):
        p, c, r = embeds.cyclical3_screen(prob, conf_tuple_range=(lo, hi))
        parts.append(p)
        kept.append(r.kept_indices + base)
        base += r.n_poses
    assert base == rep.n_poses
    assert np.array_equal(np.concatenate(kept), rep.kept_indices)
    assert np.array_equal(np.concatenate(parts), poses)


def test_trimolecular_many_survivors_per_group(gpu):
    """A loose clash threshold leaves most of the 216 poses of a group alive, so groups accumulate more
    than 32 kept poses: exercises the lane-parallel moment screen over several 32-pose batches and the
    in-order exact evaluation of the pairs it cannot rule out."""
    emb = make_embedder("cyclical", n_mols=3, n_conf=1, n_atoms=[60, 50, 55], seed=7, n_reactive=2, n_orb=1,
                        thresh=0.2)
    prob = problem.cyclical_problem(emb)
    emb.b200_want_status = True
    poses = embeds.cyclical_embed(emb)
    rep = emb.b200_report
    ref = _check_problem(prob, poses, emb.constrained_indices, rep)
    per_group = np.bincount(ref["kept"] // len(prob.angles), minlength=len(ref["groups"]))
    assert per_group.max() > 32, per_group
    assert rep.n_clash_pass > rep.n_kept > 0


@pytest.mark.parametrize("internal_as_array", [True, False])
def test_trimolecular_pairing_filter(gpu, internal_as_array):
    """User pairings restrict the orientations (embeds.py:473-476); a pairing that is an internal constraint
    only counts when internal_constraints is NOT an ndarray (quirk N10, embeds.py:820-826)."""
    emb = make_embedder("cyclical", n_mols=3, n_conf=[2, 2, 1], n_atoms=14, seed=5, n_reactive=2, n_orb=1)
    base = problem.cyclical_problem(emb)
    supers = port.cyclical_groups_trimol(base)
    assert supers
    couple = tuple(int(x) for x in supers[0]["ids"][0][0])      # a couple of orientation 0
    bogus = (0, 1)                                           # never a couple between two molecules
    emb.pairings_table = {"a": couple, "b": bogus}
    emb.internal_constraints = np.array([bogus]) if internal_as_array else [bogus]
    prob = problem.cyclical_problem(emb)
    poses, constrained, rep = embeds.cyclical3_screen(prob)
    ref = port.cyclical_embed_trimol(prob, ties=port.Ties(eps=1e-6, forced=rep.forced_decisions()))
    assert len(ref["groups"]) == len(rep.group_choice)
    if internal_as_array:
        assert rep.n_poses == 0          # the bogus pairing can never be satisfied
    else:
        assert 0 < len(rep.group_choice) < 8 * len(supers)
        assert np.array_equal(rep.kept_indices, ref["kept"])
        assert np.array_equal(constrained, ref["constrained"])
        assert all(couple in [tuple(c) for c in g["ids"]] for g in ref["groups"])


def test_trimolecular_c2_shaped_multi_conformer(gpu):
    """a C2-shaped case -- three molecules of 60 atoms with three conformers EACH (27 conformer
    triples, 216 groups, 46 656 poses of 180 atoms) -- against the oracle port: direction-search choices, clash mask,
    kept indices in order, coordinates and constrained pairs."""
    emb = make_embedder("cyclical", n_mols=3, n_conf=3, n_atoms=60, seed=11, n_reactive=2, n_orb=1, thresh=0.9)
    prob = problem.cyclical_problem(emb)
    emb.b200_want_status = True
    poses = embeds.cyclical_embed(emb)
    rep = emb.b200_report
    ref = _check_problem(prob, poses, emb.constrained_indices, rep)
    assert len(ref["groups"]) == 216 and rep.n_poses == 216 * 216
    assert len(ref["kept"]) > 500 and int(ref["clash_pass"].sum()) > 2000
    confs = {tuple(g["conf"]) for g in ref["groups"]} if "conf" in ref["groups"][0] else None
    assert confs is None or len(confs) == 27
