"""Torsion rotation with clash filtering, reference call surface:
``rotate_dihedral(coords, torsion, angle, mask=...)`` (prism_pruner.utils, call form at
torsion_module.py:529,537), ``torsion_comp_check(coords, torsion, mask, thresh, max_clashes)``
(torsion_module.py:894-918) and ``get_rotation_mask(graph, torsion)`` (torsion_module.py:354-382),
plus the batched ``torsion_scan`` over (conformer, torsion, angle) grids that the GPU runs
(C-ABI ``fc_torsion_scan``).
"""

from __future__ import annotations

import numpy as np

from . import _lib, conventions
from .embeds import _ptr

STATUS_PASS, STATUS_NEAR = 1, 4


def get_rotation_mask(graph, torsion):
    """Atoms that move when rotating about the i2-i3 bond: everything reachable from i4 without
    crossing i2-i3, except i3 itself (torsion_module.py:354-382).  Host-side graph walk."""
    _, i2, i3, i4 = (int(t) for t in torsion)
    seen, stack = {i4}, [i4]
    while stack:
        k = stack.pop()
        for nb in graph.neighbors(k):
            if (k == i2 and nb == i3) or (k == i3 and nb == i2):
                continue
            if nb not in seen:
                seen.add(nb)
                stack.append(nb)
    mask = np.array([i in seen for i in graph.nodes], dtype=bool)
    mask[i3] = False
    return mask


def torsion_scan(coords, torsions, masks, angles, thresh=1.5, max_clashes=0, want_coords=True,
                 want_min_dist=False):
    """Rotate every conformer about every torsion by every angle and clash-check the result.

    coords (C, N, 3) or (N, 3); torsions (T, 4) int; masks (T, N) bool; angles (A,) degrees.
    Returns dict: passed (C, T, A) bool, near (C, T, A) bool (min moved-static distance within 1e-6
    of thresh), coords (C, T, A, N, 3) if want_coords, min_dist (C, T, A) if want_min_dist."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
    if x.ndim == 2:
        x = x[None]
    tors = np.ascontiguousarray(np.asarray(torsions, dtype=np.int32).reshape(-1, 4))
    m = np.ascontiguousarray(np.asarray(masks, dtype=bool).reshape(len(tors), -1).astype(np.uint8))
    ang = np.ascontiguousarray(np.asarray(angles, dtype=np.float64).ravel())
    c, n = x.shape[:2]
    assert m.shape[1] == n
    t, a = len(tors), len(ang)
    status = np.zeros((c, t, a), dtype=np.uint8)
    out = np.zeros((c, t, a, n, 3), dtype=np.float64) if want_coords else None
    dmin = np.zeros((c, t, a), dtype=np.float64) if want_min_dist else None
    rc = lib.fc_torsion_scan(_ptr(x), c, n, _ptr(tors), t, _ptr(m), _ptr(ang), a, float(thresh), int(max_clashes),
                             int(conventions.ROT_HANDEDNESS), int(conventions.TORSION_AXIS_SIGN), _ptr(out),
                             _ptr(status), _ptr(dmin))
    _lib.check(rc, "fc_torsion_scan")
    res = {"passed": (status & STATUS_PASS).astype(bool), "near": (status & STATUS_NEAR).astype(bool)}
    if want_coords:
        res["coords"] = out
    if want_min_dist:
        res["min_dist"] = dmin
    return res


def rotate_dihedral(coords, dihedral, angle, mask=None, indices_to_be_moved=None):
    """Rotate the atoms selected by ``mask`` about the i2-i3 axis of ``dihedral`` by ``angle``
    degrees (prism_pruner.utils.rotate_dihedral). Returns a new (N, 3) array."""
    x = np.asarray(coords, dtype=np.float64)
    if mask is None:
        mask = np.zeros(len(x), dtype=bool)
        mask[list(indices_to_be_moved)] = True
    res = torsion_scan(x, [dihedral], [mask], [angle], thresh=0.0, want_coords=True)
    return res["coords"][0, 0, 0]


def torsion_comp_check(coords, torsion, mask, thresh=1.5, max_clashes=0):
    """True if the already-rotated structure has at most ``max_clashes`` moved-static atom pairs
    closer than ``thresh`` (torsion_module.py:894-918)."""
    res = torsion_scan(coords, [torsion], [mask], [0.0], thresh=thresh, max_clashes=max_clashes, want_coords=False)
    return bool(res["passed"][0, 0, 0])
