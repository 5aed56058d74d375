"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the in-tree hot path of
FIRECODE on plain arrays.  It travels to the GPU box (where /root/reference does not exist) and is
the checker the CUDA path is compared against; tests/test_oracle_pinning.py pins it against the
UNMODIFIED reference (oracle.loader) in the CPU container and against tests/golden/.

Parity status: functions that only restate in-tree reference code are pinned against the
reference itself; everything that goes through oracle.prism_pruner is "parity unpinned"
(see oracle/prism_pruner/__init__.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""

from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist


# ------------------------------------------------------------------------------------------------
# compenetration check -- firecode/utils.py:507-575
# ------------------------------------------------------------------------------------------------
def compenetration_check(coords, ids=None, thresh=1.0, max_clashes=0):
    """Fragment-based branches of utils.py:544-575 (ids given)."""
    coords = np.asarray(coords, dtype=float)
    assert ids is not None
    if len(ids) == 2:
        m1, m2 = coords[: ids[0]], coords[ids[0]:]
        return int(np.count_nonzero(cdist(m2, m1) < thresh)) <= max_clashes  # utils.py:551
    n1, n2 = ids[0], ids[0] + ids[1]
    m1, m2, m3 = coords[:n1], coords[n1:n2], coords[n2:]
    clashes = 0
    for x, y in ((m2, m1), (m3, m2), (m1, m3)):  # utils.py:563-571, `<=` and early return
        clashes += int(np.count_nonzero(cdist(x, y) <= thresh))
        if clashes > max_clashes:
            return False
    return True


def place(frag, xf):
    """get_embed expression embeds.py:815-817 for one molecule: (R @ X.T).T + t."""
    rot = np.asarray(xf[:9], dtype=float).reshape(3, 3)
    return (rot @ np.asarray(frag, dtype=float).T).T + np.asarray(xf[9:12], dtype=float)


def clash_batch(frag_a, frag_b, xf, thresh=1.0, max_clashes=0, conf_a=None, conf_b=None,
                strict=True, chunk=2048):
    """Reference decision for every pose of a batch, vectorised over poses.

    Returns (mask (n,) bool, dmin (n,) f64, closest (n,) f64 = min |d - thresh|)."""
    a = np.asarray(frag_a, dtype=float)
    b = np.asarray(frag_b, dtype=float)
    a = a[None] if a.ndim == 2 else a
    b = b[None] if b.ndim == 2 else b
    xf = np.asarray(xf, dtype=float).reshape(-1, 12)
    n = len(xf)
    ca = np.zeros(n, dtype=int) if conf_a is None else np.asarray(conf_a)
    cb = np.zeros(n, dtype=int) if conf_b is None else np.asarray(conf_b)
    mask = np.zeros(n, dtype=bool)
    dmin = np.zeros(n)
    closest = np.zeros(n)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        rot = xf[s:e, :9].reshape(-1, 3, 3)
        placed = np.einsum("pij,pnj->pni", rot, b[cb[s:e]]) + xf[s:e, None, 9:12]
        diff = placed[:, :, None, :] - a[ca[s:e]][:, None, :, :]
        d = np.sqrt((diff * diff).sum(axis=-1))  # cdist(m2, m1): euclidean
        hits = (d < thresh) if strict else (d <= thresh)
        mask[s:e] = hits.reshape(e - s, -1).sum(axis=1) <= max_clashes
        dmin[s:e] = d.reshape(e - s, -1).min(axis=1)
        closest[s:e] = np.abs(d - thresh).reshape(e - s, -1).min(axis=1)
    return mask, dmin, closest
