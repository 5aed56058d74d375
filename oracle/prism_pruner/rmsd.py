"""prism_pruner.rmsd restated (TEST INFRASTRUCTURE; see package docstring).

Call sites: firecode/utils.py:499 (center default False), firecode/embedder.py:1784
(center=True), firecode/hypermolecule_class.py:77.
"""

from __future__ import annotations

import numpy as np


def get_alignment_matrix(p, q):
    """Kabsch rotation for already-centred (n,3) arrays: minimises |p @ R - q|."""
    cov = p.T @ q
    v, s, w = np.linalg.svd(cov)
    if (np.linalg.det(v) * np.linalg.det(w)) < 0.0:
        s[-1] = -s[-1]
        v[:, -1] = -v[:, -1]
    return v @ w


def rmsd_and_max(p, q, center=False):
    """(RMSD, max per-atom deviation) after optimal superposition of p onto q."""
    p = np.array(p, dtype=float)
    q = np.array(q, dtype=float)
    if center:
        p = p - p.mean(axis=0)
        q = q - q.mean(axis=0)
    rot = get_alignment_matrix(p, q)
    diff = p @ rot - q
    rmsd = np.sqrt((diff * diff).sum() / len(diff))
    max_delta = np.sqrt((diff * diff).sum(axis=1)).max()
    return float(rmsd), float(max_delta)
