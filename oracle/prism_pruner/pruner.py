"""prism_pruner.pruner restated (TEST INFRASTRUCTURE; PARITY UNPINNED, see package docstring).

Call sites: firecode/embedder.py:1452,1472,1489; firecode/ensemble.py:211,230,253;
firecode/operators.py:613-632; firecode/interfaces/goat.py:399; atropisomer_module.py:504.
In-tree structural analogue of the driver: firecode/torsion_module.py:957-1043.

Driver: out_mask = ones(n); for k in K_SCHEDULE: if k == 1 or MIN_PER_CHUNK * k < active:
split the array in k contiguous chunks of n // k structures (last chunk takes the remainder) and
resolve similar pairs inside every chunk; dissimilar pairs are cached so they are evaluated once.
Which member of a similar pair is dropped and whether a pass reads a snapshot of the mask are the
unpinned switches conventions.PRUNE_KEEP / PRUNE_PASS_MODE.
"""

from __future__ import annotations

import numpy as np

from . import conventions
from .algebra import get_inertia_moments
from .periodic_table import MASSES_TABLE
from .rmsd import rmsd_and_max

K_SCHEDULE = (500_000, 200_000, 100_000, 50_000, 20_000, 10_000, 5000, 2000, 1000, 500, 200, 100,
              50, 20, 10, 5, 2, 1)


def chunk_bounds(n, k):
    """[(first, last)] for k contiguous chunks of n // k structures; the last takes the rest."""
    size = int(n // k)
    out = []
    for c in range(int(k)):
        first = c * size
        last = n if c == k - 1 else size * (c + 1)
        out.append((first, last))
    return out


class PruneStats:
    def __init__(self):
        self.eval_calls = 0
        self.cache_calls = 0
        self.passes = []


def prune_mask(n, evaluate_sim, energies=None, max_dE=0.0, keep=None, pass_mode=None,
               min_per_chunk=None, stats=None):
    """Run the multi-pass chunked pruning driver over n structures and return the bool mask."""
    keep = conventions.PRUNE_KEEP if keep is None else keep
    pass_mode = conventions.PRUNE_PASS_MODE if pass_mode is None else pass_mode
    min_per_chunk = conventions.PRUNE_MIN_PER_CHUNK if min_per_chunk is None else min_per_chunk
    assert keep in ("first", "last") and pass_mode in ("greedy", "snapshot")
    use_e = energies is not None
    if use_e:
        energies = np.asarray(energies, dtype=float)

    mask = np.ones(n, dtype=bool)
    cache = set()

    def similar(i, j):
        a, b = (i, j) if i < j else (j, i)
        if (a, b) in cache:
            if stats is not None:
                stats.cache_calls += 1
            return False
        if use_e and not (abs(energies[a] - energies[b]) < max_dE):
            return False
        if stats is not None:
            stats.eval_calls += 1
        if evaluate_sim(a, b):
            return True
        cache.add((a, b))
        return False

    for k in K_SCHEDULE:
        active = int(np.count_nonzero(mask))
        if not (k == 1 or min_per_chunk * k < active):
            continue
        if stats is not None:
            stats.passes.append((k, active))
        in_mask = mask.copy() if pass_mode == "snapshot" else mask
        out_mask = mask.copy() if pass_mode == "snapshot" else mask
        for first, last in chunk_bounds(n, k):
            order = range(first, last) if keep == "first" else range(last - 1, first - 1, -1)
            if pass_mode == "greedy":
                # NMS sweep: a structure that is still active drops every later (keep-first) /
                # earlier (keep-last) active structure similar to it.
                for i in order:
                    if not mask[i]:
                        continue
                    others = range(i + 1, last) if keep == "first" else range(i - 1, first - 1, -1)
                    for j in others:
                        if mask[j] and similar(i, j):
                            mask[j] = False
            else:
                # snapshot: structure i is dropped if any structure on its "keeper" side that was
                # active at the start of the pass is similar to it.
                for i in order:
                    if not in_mask[i]:
                        continue
                    others = range(first, i) if keep == "first" else range(i + 1, last)
                    for j in others:
                        if in_mask[j] and similar(i, j):
                            out_mask[i] = False
                            break
        mask = out_mask
    return mask


def _heavy_mask(atoms):
    return np.array([str(a) != "H" for a in atoms], dtype=bool)


def prune_by_rmsd(structures, atoms, max_rmsd=0.25, max_dev=None, energies=None, max_dE=0.0,
                  debugfunction=None, logfunction=None, stats=None, ties=None, **switches):
    """Drop structures whose heavy-atom Kabsch RMSD to a kept one is < max_rmsd and whose largest
    atomic deviation is < max_dev (default 2 * max_rmsd). Returns (structures[mask], mask)."""
    structures = np.asarray(structures, dtype=float)
    max_dev = conventions.PRUNE_MAXDEV_FACTOR * max_rmsd if max_dev is None else max_dev
    sel = _heavy_mask(atoms) if conventions.PRUNE_RMSD_HEAVY_ONLY else np.ones(len(atoms), bool)
    work = structures[:, sel, :]

    def evaluate_sim(i, j):
        rmsd, maxdev = rmsd_and_max(work[i], work[j], center=True)
        if ties is None:
            return rmsd < max_rmsd and maxdev < max_dev
        # near-threshold bookkeeping for the parity tests (oracle.port.Ties): keys use (later, earlier)
        import operator

        return ties.decide(("rmsd", j, i), rmsd, max_rmsd, operator.lt) and \
            ties.decide(("maxdev", j, i), maxdev, max_dev, operator.lt)

    mask = prune_mask(len(structures), evaluate_sim, energies, max_dE, stats=stats, **switches)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_rmsd - kept {int(mask.sum())}/{len(mask)}")
    return structures[mask], mask


def principal_moments(structures, atoms):
    masses = np.array([MASSES_TABLE[str(a)] for a in atoms])
    return np.array([get_inertia_moments(s, masses) for s in np.asarray(structures, dtype=float)])


def prune_by_moment_of_inertia(structures, atoms, max_deviation=None, energies=None, max_dE=0.0,
                               debugfunction=None, logfunction=None, stats=None, **switches):
    """Drop structures whose three principal moments of inertia are all within max_deviation
    (relative, default 1 %) of a kept one (CHANGELOG.md:256)."""
    structures = np.asarray(structures, dtype=float)
    max_deviation = conventions.MOI_MAX_DEVIATION if max_deviation is None else max_deviation
    moi = principal_moments(structures, atoms)

    def evaluate_sim(i, j):
        return bool(np.all(np.abs(moi[i] - moi[j]) / moi[i] < max_deviation))

    mask = prune_mask(len(structures), evaluate_sim, energies, max_dE, stats=stats, **switches)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_moment_of_inertia - kept {int(mask.sum())}/{len(mask)}")
    return structures[mask], mask


def prune_by_rmsd_rot_corr(structures, atoms, graph, max_rmsd=0.25, max_dev=None, energies=None,
                           max_dE=0.0, logfunction=None, debugfunction=None, stats=None,
                           **switches):
    """Symmetry-corrected RMSD pruning. The torsion-symmetry enumeration lives in the absent
    prism_pruner.torsion_module; this restatement falls back to plain heavy-atom RMSD pruning
    (a "next" row)."""
    return prune_by_rmsd(structures, atoms, max_rmsd=max_rmsd, max_dev=max_dev, energies=energies,
                         max_dE=max_dE, debugfunction=debugfunction, stats=stats, **switches)


def prune(structures, atoms, max_rmsd=0.25, logfunction=None, debugfunction=None, **kw):
    """MOI pruning followed by RMSD pruning; returns (structures[mask], mask)."""
    structures = np.asarray(structures, dtype=float)
    s1, m1 = prune_by_moment_of_inertia(structures, atoms, debugfunction=debugfunction)
    s2, m2 = prune_by_rmsd(s1, atoms, max_rmsd=max_rmsd, debugfunction=debugfunction)
    mask = m1.copy()
    mask[np.flatnonzero(m1)] = m2
    return s2, mask
