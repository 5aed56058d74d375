"""GPU parity of the string embed (C-ABI fc_string_screen) against the CPU oracle port."""

import numpy as np
import pytest

from firecode_b200 import embeds, problem
from firecode_b200.errors import ZeroCandidatesError
from oracle import port
from synth_embedder import make_embedder

pytestmark = pytest.mark.gpu


def _check(emb):
    prob = problem.string_problem(emb)
    poses = embeds.string_embed(emb)
    rep = emb.b200_report
    # the oracle takes the GPU's decision wherever a value lies within 1e-6 of its threshold and
    # verifies by itself that the value really is that close (Ties only consults `forced` then)
    ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
    ref = port.string_embed(prob, ties=ties)
    assert rep.n_poses == prob.n_poses
    # every near-threshold decision the oracle met was listed by the GPU
    missing = [k for k in ref["ties"].seen if k not in ties.forced]
    assert not missing, missing[:5]
    clash_gpu = (rep.status & 1).astype(bool)
    assert np.array_equal(clash_gpu, ref["clash_pass"])
    assert np.array_equal(rep.kept_indices, ref["kept"])
    assert poses.shape == ref["poses"].shape
    assert np.abs(poses - ref["poses"]).max() < 1e-5
    assert emb.constrained_indices.shape == (len(poses), 1, 2)
    return rep, ref


@pytest.mark.parametrize("n_conf,n_atoms,n_orb,seed", [(3, 30, 2, 5), (2, 12, 1, 9), (4, 45, 2, 21)])
def test_string_embed_matches_oracle(gpu, n_conf, n_atoms, n_orb, seed):
    emb = make_embedder("string", n_conf, n_atoms, seed=seed, n_orb=n_orb)
    rep, ref = _check(emb)
    assert rep.n_kept == len(ref["kept"]) > 0


def test_string_embed_c1_size(gpu):
    """BASELINE config C1: 2 x (10 conformers, 30 atoms), K = 2 centres, 36 angles = 14 400 tuples."""
    emb = make_embedder("string", 10, 30, seed=20261018, n_orb=2)
    rep, ref = _check(emb)
    assert rep.n_poses == 14400


def test_string_embed_zero_candidates(gpu):
    emb = make_embedder("string", 2, 20, seed=3, n_orb=1, thresh=50.0)  # everything clashes
    with pytest.raises(ZeroCandidatesError):
        embeds.string_embed(emb)
    assert emb.logs and "did not find any suitable disposition" in emb.logs[-1]
