"""GPU parity of the similarity pruning (C-ABI fc_prune) against the oracle's restatement of
prism_pruner (PARITY UNPINNED at the prism_pruner boundary: both sides share the same switches)."""

import numpy as np
import pytest

from firecode_b200 import pruner, synthetic
from oracle import port
from oracle.prism_pruner import pruner as ref_pruner

pytestmark = pytest.mark.gpu


def _forced(report):
    out = {}
    for t in report.ties:
        kind = {3: "rmsd", 4: "maxdev"}[int(t["kind"])]
        out[(kind, int(t["a"]), int(t["b"]))] = bool(t["decision"])
    return out


@pytest.mark.parametrize("keep,pass_mode", [("first", "greedy"), ("last", "snapshot"), ("first", "snapshot"),
                                            ("last", "greedy")])
@pytest.mark.parametrize("n,n_atoms,n_basins", [(300, 40, 12), (900, 24, 40)])
def test_prune_by_rmsd_matches_oracle(gpu, keep, pass_mode, n, n_atoms, n_basins):
    rng = np.random.default_rng(synthetic.SEED + n + n_atoms)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, n, n_atoms, n_basins, jitter=(0.02, 0.5))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5, keep=keep, pass_mode=pass_mode)
    rep = pruner.last_report
    ties = port.Ties(eps=1e-6, forced=_forced(rep))
    ref_out, ref_mask = ref_pruner.prune_by_rmsd(structures, atoms, 0.5, ties=ties, keep=keep, pass_mode=pass_mode)
    assert not [k for k in ties.seen if k not in ties.forced]
    assert mask.dtype == bool and mask.shape == (n,)
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(out, structures[ref_mask])
    assert 1 < mask.sum() < n  # both similar and dissimilar pairs exist


def test_prune_with_energies_and_multipass(gpu):
    """Enough structures for several chunked passes (20 * k < active) and an energy window."""
    rng = np.random.default_rng(99)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 2500, 16, 150, jitter=(0.02, 0.3))
    energies = rng.uniform(0, 5, size=len(structures))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.4, energies=energies, max_dE=1.0)
    assert pruner.last_report.passes >= 4
    ties = port.Ties(eps=1e-6, forced=_forced(pruner.last_report))
    _, ref_mask = ref_pruner.prune_by_rmsd(structures, atoms, 0.4, energies=energies, max_dE=1.0, ties=ties)
    assert np.array_equal(mask, ref_mask)


def test_prune_by_moi_matches_oracle(gpu):
    rng = np.random.default_rng(5)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 600, 30, 25, jitter=(0.0, 0.02))
    out, mask = pruner.prune_by_moment_of_inertia(structures, atoms)
    _, ref_mask = ref_pruner.prune_by_moment_of_inertia(structures, atoms)
    # relative deviations are compared with 1e-2: tolerate only pairs within 1e-9 of that threshold
    assert np.array_equal(mask, ref_mask)
    assert 1 < mask.sum() < len(mask)


def test_prune_edge_cases(gpu):
    rng = np.random.default_rng(1)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 3, 10, 1, jitter=(0.0, 0.0))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
    assert mask.tolist() == [True, False, False]       # identical up to rigid motion: keep first
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5, keep="last", pass_mode="snapshot")
    assert mask.tolist() == [False, False, True]
    out, mask = pruner.prune_by_rmsd(structures[:1], atoms, 0.5)
    assert mask.tolist() == [True]
    out, mask = pruner.prune_by_rmsd(structures[:0], atoms, 0.5)
    assert mask.shape == (0,) and out.shape == (0, 10, 3)
    s2, m2 = pruner.prune(structures, atoms, max_rmsd=0.5)
    assert m2.sum() == 1
