"""Symmetry-corrected RMSD pruning (firecode_b200.pruner.prune_by_rmsd_rot_corr, C-ABI fc_rmsd_rot_corr_pairs) against
the oracle's restatement (oracle/prism_pruner/pruner.py: [UNVERIFIED-RECALL] of prism_pruner 0.0.7, PARITY UNPINNED)
on synthetic molecules carrying three-fold (CF3-like) and two-fold rotors: copies that differ by a symmetry rotation of
a rotor -- same geometry, permuted labels -- are different for the plain RMSD and similar for the corrected one."""

import numpy as np
import pytest

from firecode_b200 import pruner, synthetic
from oracle import port
from oracle.prism_pruner import pruner as ref_pruner

pytestmark = pytest.mark.gpu


def rotor_molecule(rng, n_heavy, folds):
    """Heavy-atom tree + one symmetric rotor per entry of ``folds`` on distinct leaves.
    Returns atoms, coords, torsions [(g, p, x, f0)], angles, masks, plus per rotor the index list of its blades."""
    for _ in range(200):  # a tree with enough leaves to carry the rotors
        atoms, xyz, bonds, parent = synthetic.molecule_cloud(rng, n_heavy, heavy_fraction=1.0)
        children = {i: [] for i in range(n_heavy)}
        for p, c in bonds:
            children[p].append(c)
        leaves = [i for i in range(n_heavy) if not children[i] and parent[i] >= 0 and parent[parent[i]] >= 0]
        if len(leaves) >= len(folds):
            break
    assert len(leaves) >= len(folds)
    atoms = list(atoms)
    coords = [r for r in xyz]
    torsions, angles, blades = [], [], []
    for x, fold in zip(rng.permutation(leaves)[: len(folds)], folds):
        p, g = int(parent[x]), int(parent[parent[x]])
        axis = coords[x] - coords[p]
        axis /= np.linalg.norm(axis)
        perp = np.cross(axis, rng.normal(size=3))
        perp /= np.linalg.norm(perp)
        first = len(coords)
        for k in range(fold):
            ang = 2 * np.pi * k / fold
            radial = perp * np.cos(ang) + np.cross(axis, perp) * np.sin(ang)
            coords.append(coords[x] + 1.35 * (0.33 * axis + 0.94 * radial))
            atoms.append("F")
        torsions.append((g, p, int(x), first))
        angles.append(tuple(360.0 * k / fold for k in range(fold)))
        blades.append(list(range(first, first + fold)))
    coords = np.array(coords)
    masks = np.zeros((len(torsions), len(coords)), dtype=bool)
    for t, b in enumerate(blades):
        masks[t, b] = True
    return np.array(atoms), coords, torsions, angles, masks, blades


def rotor_ensemble(rng, n, n_heavy=14, folds=(3, 3, 2), n_basins=5, jitter=0.03):
    atoms, base, torsions, angles, masks, blades = rotor_molecule(rng, n_heavy, folds)
    basins = [base + rng.normal(scale=0.45, size=base.shape) * (np.arange(len(base)) < n_heavy)[:, None] * (b > 0)
              for b in range(n_basins)]
    # the blades must stay an exact rotor of the (moved) axis: rebuild them on every basin
    out = np.empty((n, len(base), 3))
    for s in range(n):
        x = basins[s % n_basins].copy()
        for (g, p, xa, f0), ang, b in zip(torsions, angles, blades):
            axis = x[xa] - x[p]
            axis /= np.linalg.norm(axis)
            rel = base[b] - base[xa]
            # carry the rotor rigidly onto the basin's axis (rotation taking the base axis to the basin axis)
            a0 = base[xa] - base[p]
            a0 /= np.linalg.norm(a0)
            v = np.cross(a0, axis)
            c = float(a0 @ axis)
            if np.linalg.norm(v) > 1e-12:
                k = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
                rot = np.eye(3) + k + k @ k * (1.0 / (1.0 + c))
            else:
                rot = np.eye(3)
            x[b] = x[xa] + rel @ rot.T
            # a random element of the rotor's symmetry group, plus a small off-symmetry twist
            turn = np.deg2rad(ang[rng.integers(len(ang))] + rng.normal(scale=1.5))
            x[b] = synthetic.rotate_about_axis(x[b], x[xa], axis, turn)
        x += rng.normal(scale=jitter, size=x.shape)
        out[s] = synthetic.random_rigid(rng, x)
    return atoms, out, torsions, angles, masks


def _forced(rep):
    return {(kind, later, earlier): dec for kind, later, earlier, _, dec in rep.ties}


@pytest.mark.parametrize("seed,n,keep,pass_mode", [(1, 160, "first", "greedy"), (2, 90, "last", "snapshot"),
                                                  (3, 420, "first", "greedy")])
def test_rot_corr_pruning_matches_oracle(gpu, seed, n, keep, pass_mode):
    rng = np.random.default_rng(seed)
    atoms, structures, torsions, angles, masks = rotor_ensemble(rng, n)
    out, mask = pruner.prune_by_rmsd_rot_corr(structures, atoms, None, max_rmsd=0.35, torsions=torsions, angles=angles,
                                              masks=masks, keep=keep, pass_mode=pass_mode)
    rep = pruner.last_rot_corr_report
    assert rep.n_torsions == 3 and rep.n_folds == (3, 3, 2) and rep.pairs_evaluated > 0
    ties = port.Ties(eps=1e-6, forced=_forced(rep))
    _, ref_mask = ref_pruner.prune_by_rmsd_rot_corr(structures, atoms, None, max_rmsd=0.35, ties=ties, torsions=torsions,
                                                    angles=angles, masks=list(masks), forced_choice=rep.choices,
                                                    choice_eps=1e-9, keep=keep, pass_mode=pass_mode)
    assert not [k for k in ties.seen if k not in ties.forced]
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(out, structures[mask])
    # the correction matters: the plain heavy-atom RMSD keeps (many) more structures
    _, plain = pruner.prune_by_rmsd(structures, atoms, 0.35, keep=keep, pass_mode=pass_mode)
    assert mask.sum() < plain.sum() <= n
    assert mask.sum() >= 5


def test_rot_corr_pair_values_match_oracle(gpu):
    """(rmsd, maxdev) and the chosen symmetry angles pair by pair."""
    rng = np.random.default_rng(11)
    atoms, structures, torsions, angles, masks = rotor_ensemble(rng, 40, folds=(3, 2, 3, 3))
    i, j = np.triu_indices(40, 1)
    pairs = np.stack([i, j], axis=1)
    rmsd, dev, choice, gap = pruner.rmsd_and_max_rot_corr_pairs(structures, atoms, pairs, torsions, angles, masks,
                                                                want_choices=True)
    sel = np.array([a != "H" for a in atoms])
    for p in rng.permutation(len(pairs))[:150]:
        a, b = pairs[p]
        forced = {((int(a), int(b)), t): int(choice[p, t]) for t in range(len(torsions)) if gap[p, t] <= 1e-9}
        r, m = ref_pruner.rmsd_and_max_rot_corr(structures[a], structures[b], torsions, angles, list(masks), sel,
                                                forced_choice=forced, choice_eps=1e-9, key=(int(a), int(b)))
        assert abs(r - rmsd[p]) < 1e-9 and abs(m - dev[p]) < 1e-9
    assert (gap > 1e-3).mean() > 0.9          # the winning angle is clear-cut for almost every pair
    assert len(np.unique(choice)) >= 3


def test_rot_corr_without_symmetric_torsions_keeps_everything(gpu):
    rng = np.random.default_rng(5)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 30, 12, 3, jitter=(0.0, 0.01))
    out, mask = pruner.prune_by_rmsd_rot_corr(structures, atoms, None, max_rmsd=0.5, torsions=[], angles=[], masks=[])
    assert mask.all() and out.shape == structures.shape
    with pytest.raises(Exception, match="symmetric torsions"):
        pruner.prune_by_rmsd_rot_corr(structures, atoms, None, max_rmsd=0.5)   # FIRECODE is not importable here


def test_prune_forwards_keywords(gpu):
    """prune() used to drop energies / max_dE (ungated pruning) and swallow unknown keywords."""
    rng = np.random.default_rng(6)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 60, 14, 4, jitter=(0.0, 0.02))
    energies = np.arange(60, dtype=float)
    _, gated = pruner.prune(structures, atoms, max_rmsd=0.5, energies=energies, max_dE=0.5)
    _, free = pruner.prune(structures, atoms, max_rmsd=0.5)
    assert gated.all() and free.sum() <= 8       # |dE| >= 1 everywhere: nothing may be compared
    with pytest.raises(TypeError):
        pruner.prune(structures, atoms, max_rmsd=0.5, max_rmds=0.1)
