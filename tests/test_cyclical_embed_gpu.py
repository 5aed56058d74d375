"""GPU parity of the bimolecular cyclical embed (C-ABI fc_cyclical_screen) against the oracle."""

import numpy as np
import pytest

from firecode_b200 import embeds, problem
from firecode_b200.errors import ZeroCandidatesError
from oracle import port
from synth_embedder import make_embedder

pytestmark = pytest.mark.gpu


def _check(emb):
    prob = problem.cyclical_problem(emb)
    n_tot = sum(c.shape[1] for c in prob.coords)
    try:
        poses = embeds.cyclical_embed(emb)
    except ZeroCandidatesError:
        poses = np.zeros((0, n_tot, 3))
    rep = emb.b200_report
    ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
    ref = port.cyclical_embed_bimol(prob, ties=ties)
    missing = [k for k in ref["ties"].seen if k not in ties.forced]
    assert not missing, missing[:5]
    assert rep.n_poses == len(ref["clash_pass"])
    assert np.array_equal((rep.status & 1).astype(bool), ref["clash_pass"])
    assert np.array_equal(rep.kept_indices, ref["kept"])
    assert poses.shape == ref["poses"].shape
    assert np.abs(poses - ref["poses"]).max() < 1e-5
    assert np.array_equal(np.asarray(emb.constrained_indices).reshape(-1, 2, 2), ref["constrained"])
    return rep, ref


@pytest.mark.parametrize("n_conf,n_atoms,n_orb,n_reactive,seed", [(3, 24, 2, 2, 11), (2, 16, 1, 2, 4),
                                                                  (4, 40, 2, 2, 7), (3, 20, 2, 1, 13)])
def test_cyclical_embed_matches_oracle(gpu, n_conf, n_atoms, n_orb, n_reactive, seed):
    emb = make_embedder("cyclical", n_conf, n_atoms, seed=seed, n_reactive=n_reactive, n_orb=n_orb)
    rep, ref = _check(emb)
    assert rep.n_poses > 0


def test_cyclical_embed_bigger(gpu):
    """Two 6-conformer ensembles of 60 atoms: 36 conformer pairs x 16 pivot pairs x 2 x 36 poses."""
    emb = make_embedder("cyclical", 6, 60, seed=20261018, n_reactive=2, n_orb=2)
    rep, ref = _check(emb)
    assert rep.n_poses > 10000
