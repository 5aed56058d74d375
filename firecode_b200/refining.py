"""The refining steps that follow the embed in FIRECODE's pipeline, batched on the GPU: the callers on the
far side of the embedding screen.

Each function takes the embedder (duck-typed, exactly the attributes the reference method reads) and does what
the ``RunEmbedding`` method of the same name does (/root/reference/firecode/embedder.py):

* ``compenetration_refining``  (embedder.py:1954-1995)  per-structure ``compenetration_check`` loop -> one call
* ``fitness_refining``         (embedder.py:1997-2039)  per-structure ``fitness_check`` loop -> one call
* ``similarity_refining``      (embedder.py:1410-1514)  TFD / MOI / RMSD / symmetry-corrected RMSD cascade with
  ``apply_mask`` over the dependent attributes

A maintainer can assign them as methods (``RunEmbedding.fitness_refining = refining.fitness_refining``); the
log lines are the reference's.
"""

from __future__ import annotations

import time

import numpy as np

from . import _lib
from .errors import ZeroCandidatesError

NEAR_EPS = 1e-6


def fitness_check_batch(structures, constrained_indices, constrained_distances, threshold=5.0, return_errors=False):
    """``fitness_check(structure, constraints, targets, threshold)`` (optimization_methods.py:163-180) for a batch:
    structures (P, N, 3); constrained_indices (P, M, 2) int; constrained_distances (P, M) float with None / NaN
    where a constraint has no target.  Returns the bool mask (error < threshold), optionally the errors."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    assert x.ndim == 3 and x.shape[2] == 3
    p = len(x)
    if p == 0:
        empty = np.zeros(0, dtype=bool)
        return (empty, np.zeros(0)) if return_errors else empty
    pairs = np.ascontiguousarray(np.asarray(constrained_indices, dtype=np.int32).reshape(p, -1, 2))
    m = pairs.shape[1]
    tg = np.array([[np.nan if t is None else float(t) for t in row] for row in constrained_distances],
                  dtype=np.float64).reshape(p, m)
    err = np.zeros(p, dtype=np.float64)
    _lib.check(lib.fc_fitness_batch(x.ctypes.data, p, x.shape[1], pairs.ctypes.data, tg.ctypes.data, m,
                                    err.ctypes.data), "fc_fitness_batch")
    mask = err < threshold
    return (mask, err) if return_errors else mask


def fitness_check(coords, constraints, targets, threshold):
    """Single-structure form with the reference's signature (optimization_methods.py:163)."""
    constraints = np.asarray(list(constraints), dtype=np.int32).reshape(1, -1, 2)
    return bool(fitness_check_batch(np.asarray(coords)[None], constraints, [list(targets)], threshold)[0])


def apply_mask(embedder, attributes, mask):
    """embedder.py:1399-1408."""
    for attr in attributes:
        if hasattr(embedder, attr):
            try:
                setattr(embedder, attr, getattr(embedder, attr)[mask])
            except IndexError:
                pass


def _zero_candidates_check(embedder):
    if len(embedder.structures) == 0:
        if hasattr(embedder, "log_warnings"):
            embedder.log_warnings()
        raise ZeroCandidatesError()


def compenetration_refining(embedder):
    """embedder.py:1954-1995: multi-fragment clash check of every structure (skipped for the embeds that check
    while embedding), then the energies / exit_status initialisation."""
    from .utils import compenetration_check_structures

    if embedder.embed not in ("string", "cyclical", "monomolecular"):
        embedder.log("--> Checking structures for compenetrations")
        t_start = time.perf_counter()
        mask = compenetration_check_structures(embedder.structures, embedder.ids, thresh=embedder.options.clash_thresh,
                                               max_clashes=embedder.options.max_clashes)
        apply_mask(embedder, ("structures", "constrained_indices"), mask)
        if False in mask:
            embedder.log(f"Discarded {int((~mask).sum())} candidates for compenetration ({int(mask.sum())} left, "
                         f"{time.perf_counter() - t_start:.3f} s)")
        else:
            embedder.log(f"All {len(mask)} structures passed the compenetration check")
        embedder.log()
        _zero_candidates_check(embedder)
    embedder.energies = np.full(len(embedder.structures), 1e10, dtype=float)
    embedder.exit_status = np.zeros(len(embedder.structures), dtype=bool)


def fitness_refining(embedder, threshold=5.0, verbose=False):
    """embedder.py:1997-2039: drop the structures whose constrained distances deviate, in sum, by ``threshold``
    or more from the imposed pairing distances."""
    if verbose:
        embedder.log(" \n--> Fitness pruning - removing inaccurate structures")
    ci = np.asarray(embedder.constrained_indices)
    if len(embedder.structures):
        targets = [[embedder.get_pairing_dists_from_constrained_indices(c) for c in constraints] for constraints in ci]
        mask, err = fitness_check_batch(embedder.structures, ci, targets, threshold, return_errors=True)
        embedder.b200_fitness_near = np.flatnonzero(np.abs(err - threshold) <= NEAR_EPS)
    else:
        mask = np.ones(0, dtype=bool)
    apply_mask(embedder, ("structures", "energies", "constrained_indices", "exit_status"), mask)
    if False in mask:
        embedder.log(f"Discarded {int((~mask).sum())} candidates for unfitness ({int(mask.sum())} left)")
    elif verbose:
        embedder.log("All candidates meet the imposed criteria.")
    embedder.log()
    _zero_candidates_check(embedder)


def similarity_refining(embedder, tfd=False, moi=True, rmsd=True, verbose=False):
    """embedder.py:1410-1514 with the pruning functions of this package.  The 1e5-structure guards of the
    reference are kept (they change which structures reach the next step)."""
    from . import graphs, pruner, torsion

    if verbose:
        embedder.log("--> Similarity Processing")
    before = len(embedder.structures)
    attr = ("constrained_indices", "energies", "exit_status")
    debug = getattr(embedder, "debuglog", None)
    if tfd and len(embedder.objects) > 1 and hasattr(embedder, "embed_graph") and embedder.embed_graph.is_single_molecule:
        quadruplets = graphs.quadruplets(embedder.embed_graph)
        if len(quadruplets) > 0:
            embedder.structures, mask = torsion.prune_conformers_tfd(embedder.structures, quadruplets, verbose=verbose)
            apply_mask(embedder, attr, mask)
            if False in mask:
                embedder.log(f"Discarded {int((~mask).sum())} structures for TFD similarity ({int(mask.sum())} left)")
    if moi:
        if len(embedder.structures) <= 1e5:
            embedder.structures, mask = pruner.prune_by_moment_of_inertia(embedder.structures, embedder.atoms,
                                                                          debugfunction=debug)
            apply_mask(embedder, attr, mask)
            if False in mask:
                embedder.log(f"Discarded {int((~mask).sum())} candidates for MOI similarity ({int(mask.sum())} left)")
        else:
            embedder.log("Skipped MOI pruning (>100k structures)")
    if rmsd:
        if len(embedder.structures) <= 1e5:
            embedder.structures, mask = pruner.prune_by_rmsd(embedder.structures, embedder.atoms, embedder.options.rmsd,
                                                             debugfunction=debug)
            apply_mask(embedder, attr, mask)
            if False in mask:
                embedder.log(f"Discarded {int((~mask).sum())} candidates for RMSD similarity ({int(mask.sum())} left)")
            if len(embedder.structures) <= 1e3 and hasattr(embedder, "embed_graph"):
                embedder.structures, mask = pruner.prune_by_rmsd_rot_corr(
                    embedder.structures, embedder.atoms, embedder.embed_graph, max_rmsd=embedder.options.rmsd,
                    logfunction=(embedder.log if verbose else None), debugfunction=debug)
                apply_mask(embedder, attr, mask)
                if False in mask:
                    embedder.log(f"Discarded {int((~mask).sum())} candidates for symmetry-corrected RMSD similarity "
                                 f"({int(mask.sum())} left)")
            elif hasattr(embedder, "embed_graph"):
                embedder.log("Skipped rotationally-corrected RMSD pruning (>1k structures)")
        else:
            embedder.log("Skipped RMSD pruning (>100k structures)")
    if verbose and len(embedder.structures) == before:
        embedder.log(f"All structures passed the similarity check.{' ' * 15}")
    embedder.log()
