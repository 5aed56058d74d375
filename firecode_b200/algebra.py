"""Rotation / alignment helpers with the reference's names (firecode/algebra.py and the
``prism_pruner.rmsd`` / ``prism_pruner.algebra`` functions it re-exports).  Per-structure arithmetic
runs in the CUDA library (batched C-ABI calls); closed-form 3-vector geometry stays on the host
(SURVEY.md 8a row a11)."""

from __future__ import annotations

import numpy as np

from . import _lib
from .utils import rot_mat_from_pointer, rotation_matrix_from_vectors  # noqa: F401  (re-exported)


def _ptr(a):
    return None if a is None else a.ctypes.data


def rmsd_and_max_batch(ref, structures, center=False):
    """(rmsd (n,), maxdev (n,)) of every structure against ``ref`` after optimal superposition:
    prism_pruner.rmsd.rmsd_and_max(ref, structure, center) batched (call sites utils.py:499,
    embedder.py:1784)."""
    lib = _lib.load(require_device=True)
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    x = np.ascontiguousarray(structures, dtype=np.float64)
    if x.ndim == 2:
        x = x[None]
    assert ref.shape == x.shape[1:] and ref.shape[1] == 3
    n = len(x)
    rmsd, maxdev = np.zeros(n), np.zeros(n)
    _lib.check(lib.fc_rmsd_and_max_batch(_ptr(ref), _ptr(x), n, ref.shape[0], 1 if center else 0, _ptr(rmsd),
                                         _ptr(maxdev)), "fc_rmsd_and_max_batch")
    return rmsd, maxdev


def rmsd_and_max(p, q, center=False):
    """prism_pruner.rmsd.rmsd_and_max for one pair."""
    r, m = rmsd_and_max_batch(p, np.asarray(q, dtype=np.float64)[None], center=center)
    return float(r[0]), float(m[0])


def self_clash_counts(coords, bonded=None, thresh=1.0):
    """Per structure: ordered atom pairs with 0 < d < 0.5 A (algebra.py:52-54 count_clashes) and
    ordered non-bonded pairs i != j with d < thresh (utils.py:534-540)."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(coords, dtype=np.float64)
    if x.ndim == 2:
        x = x[None]
    n, n_atoms = x.shape[:2]
    b = None if bonded is None else np.ascontiguousarray(bonded, dtype=np.uint8)
    close, nonbonded = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    _lib.check(lib.fc_self_clash_batch(_ptr(x), n, n_atoms, _ptr(b), float(thresh), _ptr(close), _ptr(nonbonded)),
               "fc_self_clash_batch")
    return close, nonbonded


def count_clashes(coords):
    """Number of atomic distances between 0 and 0.5 A, ordered pairs (algebra.py:52-54)."""
    return int(self_clash_counts(coords)[0][0])


def point_angle(p1, p2, p3):
    """Planar angle p1-p2-p3 in degrees (algebra.py:23-25)."""
    a = np.asarray(p1, dtype=float) - np.asarray(p2, dtype=float)
    b = np.asarray(p3, dtype=float) - np.asarray(p2, dtype=float)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    return float(np.arccos(np.clip(a @ b, -1.0, 1.0)) * 180 / np.pi)


def align_vec_pair(ref, tgt):
    """Rotation that optimally aligns the two ``tgt`` vectors onto the two ``ref`` vectors
    (algebra.py:28-49): Kabsch on the 2-vector covariance.  Set-up scale (once per group on the host
    side of the API; the embeds evaluate it on the device, fc_math.cuh kabsch_from_cov)."""
    ref, tgt = np.asarray(ref, dtype=float), np.asarray(tgt, dtype=float)
    B = ref[0][:, None] * tgt[0][None, :] + ref[1][:, None] * tgt[1][None, :]
    u, s, vh = np.linalg.svd(B)
    if np.linalg.det(u @ vh) < 0:
        u[:, -1] = -u[:, -1]
    return np.ascontiguousarray(u @ vh)
