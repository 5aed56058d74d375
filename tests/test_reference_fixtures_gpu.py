"""The reference's OWN test fixtures (firecode/tests/embed_string, embed_cyclical, embed_chelotropic) through
the CUDA path: the problems and the outputs of the UNMODIFIED reference are stored in tests/golden/*.npz
(oracle/make_golden.py).  Near-threshold decisions (the string fixture has structural ties: 10-degree steps
against a 10-degree TFD threshold) are compared through the port conditioned on the decisions the GPU lists."""

import os

import numpy as np
import pytest

from firecode_b200 import embeds
from oracle import port
from test_oracle_pinning import GOLDEN, _string_problem_from_npz, cyclical_problem_from_npz

pytestmark = pytest.mark.gpu


def test_reference_string_fixture(gpu):
    z = np.load(os.path.join(GOLDEN, "embed_string.npz"))
    prob = _string_problem_from_npz(z)
    poses, rep = embeds.string_screen(prob)
    ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
    ref = port.string_embed(prob, ties=ties)
    assert not [k for k in ref["ties"].seen if k not in ties.forced]
    assert np.array_equal(rep.kept_indices, ref["kept"])
    assert np.abs(poses - ref["poses"]).max() < 1e-5
    # the decisions the GPU listed are genuine ties of the reference arithmetic; where it took the same side as
    # the reference the kept poses ARE the reference's
    own = port.string_embed(prob)
    if np.array_equal(own["kept"], rep.kept_indices):
        assert poses.shape == z["ref_structures"].shape and np.abs(poses - z["ref_structures"]).max() < 1e-5


@pytest.mark.parametrize("name", ["embed_cyclical", "embed_chelotropic"])
def test_reference_cyclical_fixtures(gpu, name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    prob = cyclical_problem_from_npz(z)
    poses, constrained, rep = embeds.cyclical_screen(prob)
    ties = port.Ties(eps=1e-6, forced=rep.forced_decisions())
    ref = port.cyclical_embed_bimol(prob, ties=ties)
    assert not [k for k in ref["ties"].seen if k not in ties.forced]
    assert np.array_equal(rep.kept_indices, ref["kept"])
    if len(rep.ties) == 0:
        assert poses.shape == z["ref_structures"].shape
        assert np.abs(poses - z["ref_structures"]).max() < 1e-5
        assert np.array_equal(constrained, z["ref_constrained"])
    else:
        assert np.abs(poses - ref["poses"]).max() < 1e-5
