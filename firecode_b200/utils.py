"""Host-side helpers of the embedding screen with the reference's names and semantics
(firecode/utils.py).  Closed-form 3-vector geometry stays on the host; the
per-pose work runs in the CUDA library.
"""

from __future__ import annotations

import numpy as np

from .errors import TriangleError


def cartesian_product(*arrays):
    """Rows of the cartesian product in the reference's enumeration order (utils.py:219-221):
    numpy meshgrid 'xy' indexing, i.e. for two inputs the FIRST index varies fastest, for three
    inputs the order is (second outermost, first, third fastest) -- SURVEY.md quirk N1."""
    grids = np.meshgrid(*[np.asarray(a) for a in arrays])
    return np.stack(grids, -1).reshape(-1, len(arrays))


def rot_mat_from_pointer(pointer, angle_deg):
    """Rotation about ``pointer`` by ``angle_deg`` degrees through a unit quaternion (x, y, z, w);
    handedness follows firecode_b200.conventions.ROT_HANDEDNESS (prism_pruner.algebra contract)."""
    from . import conventions

    p = np.asarray(pointer, dtype=float)
    half = conventions.ROT_HANDEDNESS * float(angle_deg) * np.pi / 360.0
    x, y, z = p / np.linalg.norm(p) * np.sin(half)
    w = np.cos(half)
    return np.array([
        [x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w],
    ])


def rotation_matrix_from_vectors(vec1, vec2):
    """Rotation taking the direction of vec1 onto vec2 (utils.py:224-249): Rodrigues form
    I + K + K^2 (1-c)/s^2; antiparallel inputs give a half turn about z, parallel the identity."""
    vec1 = np.asarray(vec1, dtype=float)
    vec2 = np.asarray(vec2, dtype=float)
    assert vec1.shape == (3,) and vec2.shape == (3,)
    a = vec1 / np.linalg.norm(vec1)
    b = vec2 / np.linalg.norm(vec2)
    v = np.cross(a, b)
    s = np.linalg.norm(v)
    if s != 0:
        c = float(np.dot(a, b))
        k = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
        return np.eye(3) + k + k @ k * ((1 - c) / s**2)
    if np.linalg.norm(a + b) == 0:
        return rot_mat_from_pointer(np.array([0.0, 0.0, 1.0]), 180)
    return np.eye(3)


def polygonize(lengths):
    """Start/end points of the pivot vectors of a cyclical embed (utils.py:252-312).
    Two lengths: two centred collinear segments, second orientation flips segment 2;
    three lengths: the triangle with its 8 vertex-order orientations."""
    lengths = np.asarray(lengths, dtype=float)
    assert len(lengths) in (2, 3)
    base = np.zeros((len(lengths), 2, 3))
    if len(lengths) == 2:
        for i in range(2):
            base[i, 0, 0] = -lengths[i] / 2
            base[i, 1, 0] = +lengths[i] / 2
        out = np.stack([base, base.copy()])
        out[1, 1] *= -1
        return out
    if not all(lengths[i] < lengths[i - 1] + lengths[i - 2] for i in (0, 1, 2)):
        raise TriangleError(f"Impossible to build a triangle with sides {lengths}")
    l0, l1, l2 = lengths
    x = (l0**2 - l1**2 + l2**2) / (2 * (l0**2) ** 0.5)
    y = (l2**2 - x**2) ** 0.5
    base[0, 1] = (l0, 0, 0)
    base[1, 0] = (l0, 0, 0)
    base[1, 1] = (x, y, 0)
    base[2, 0] = (x, y, 0)
    out = np.stack([base.copy() for _ in range(8)])
    flips = {1: (2,), 2: (1,), 3: (1, 2), 4: (0,), 5: (0, 1), 6: (0, 2), 7: (0, 1, 2)}
    for t, vecs in flips.items():
        for v in vecs:
            out[t, v] = out[t, v][::-1].copy()
    return out


def compenetration_check(coords, graph=None, ids=None, thresh=1.0, max_clashes=0):
    """Single-structure form of the clash test (utils.py:507-575), evaluated on the GPU.

    ids of length 2: ``#{|m2_j - m1_i| < thresh} <= max_clashes`` (utils.py:544-551);
    ids of length 3: the three blocks (m2,m1), (m3,m2), (m1,m3) with ``<=`` (utils.py:553-575).
    The hot path never calls this per pose -- the screens batch it (firecode_b200.clash)."""
    coords = np.ascontiguousarray(np.asarray(coords, dtype=np.float64))
    if ids is None:
        # not fragment-based (utils.py:523-542): quick count of pairs closer than 0.5 A, then -- given
        # a graph -- the non-bonded pairs below thresh.  The reference loop counts ORDERED pairs and
        # tests the running count at the top of each iteration; its last iteration is always the
        # diagonal entry (N-1, N-1), so it returns False iff the full count exceeds max_clashes.
        from .algebra import self_clash_counts

        bonded = None
        if graph is not None:
            bonded = np.zeros((len(coords), len(coords)), dtype=np.uint8)
            for a, b in graph.edges:
                bonded[a, b] = bonded[b, a] = 1
        close, nonbonded = self_clash_counts(coords[None], bonded, thresh)
        if close[0] > max_clashes:
            return False
        if graph is None:
            return True
        return not (nonbonded[0] > max_clashes)
    return bool(compenetration_check_structures(coords[None], ids, thresh=thresh, max_clashes=max_clashes)[0])


def compenetration_check_structures(structures, ids, thresh=1.0, max_clashes=0, return_closest=False):
    """``compenetration_check(structure, ids=ids, thresh, max_clashes)`` for every structure of a batch in one
    GPU call: the loop of RunEmbedding.compenetration_refining (embedder.py:1954-1975).  Returns the bool mask
    (and, on request, min |d - thresh| per structure)."""
    from . import _lib

    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    assert x.ndim == 3 and x.shape[2] == 3
    idv = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).ravel())
    counts = np.zeros(len(x), dtype=np.int64)
    closest = np.zeros(len(x), dtype=np.float64)
    _lib.check(lib.fc_structure_clash_batch(x.ctypes.data, len(x), x.shape[1], idv.ctypes.data, len(idv), float(thresh),
                                            counts.ctypes.data, closest.ctypes.data), "fc_structure_clash_batch")
    mask = counts <= max_clashes
    return (mask, closest) if return_closest else mask


def rmsd_similarity(ref, structures, rmsd_thr=0.5):
    """True if any structure has uncentred all-atom Kabsch RMSD < rmsd_thr and max deviation
    < 2 * rmsd_thr to ``ref`` (utils.py:494-504).  The embeds run this filter fused inside
    fc_cyclical_screen / fc_cyclical3_screen; this stand-alone form evaluates the whole batch of
    structures in one call of fc_rmsd_and_max_batch."""
    from .algebra import rmsd_and_max_batch

    structures = np.asarray(structures, dtype=np.float64)
    if structures.size == 0:
        return False
    rmsd, maxdev = rmsd_and_max_batch(ref, structures, center=False)
    return bool(np.any((rmsd < rmsd_thr) & (maxdev < 2 * rmsd_thr)))


def xyz_text(atoms, structures, titles=None):
    """The xyz text of a batch of structures, byte-identical to looping the reference's ``write_xyz``
    (utils.py:105-116) over them; formatted by several host threads in the library (fc_xyz_format)."""
    import ctypes as C

    from . import _lib

    lib = _lib.load(require_device=False)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    assert x.ndim == 3 and x.shape[2] >= 3, f"{x.shape=}"
    if x.shape[2] > 3:
        x = np.ascontiguousarray(x[:, :, :3])
    n, n_atoms = x.shape[:2]
    sym = [str(a).encode() for a in atoms]
    assert len(sym) == n_atoms, f"{len(sym)=} != {n_atoms=}"
    stride = max([len(s) for s in sym] + [1])
    symbuf = b"".join(s.ljust(stride, b"\0") for s in sym)
    if titles is None:
        tbuf = None
    else:
        titles = [titles] * n if isinstance(titles, str) else list(titles)
        assert len(titles) == n
        tbuf = b"".join(str(t).encode() + b"\0" for t in titles)
    cap = n * (n_atoms * (stride + 56) + 64) + sum(len(str(t)) for t in (titles or []))
    for _ in range(2):
        out = C.create_string_buffer(max(cap, 1))
        need = C.c_int64(0)
        rc = lib.fc_xyz_format(symbuf, stride, x.ctypes.data, n, n_atoms, tbuf, out, cap, C.byref(need))
        if rc == 0:
            return out.raw[: need.value].decode()
        if need.value <= cap:
            _lib.check(rc, "fc_xyz_format")
        cap = need.value
    _lib.check(rc, "fc_xyz_format")


def write_xyz(atoms, coords, output, title="temp"):
    """utils.py:105-116: same signature, same bytes."""
    atoms, coords = np.asarray(atoms), np.asarray(coords)
    assert atoms.shape[0] == coords.shape[0], f"{atoms.shape[0]=} != {coords.shape[0]=}"
    assert coords.shape[1] >= 3, f"{coords.shape[1]=}"
    output.write(xyz_text(atoms, coords[None], [title]))


def write_xyz_batch(atoms, structures, output, titles=None):
    """All structures of an ensemble in one call (the reference loops write_xyz, e.g. embedder.py:1698-1716)."""
    output.write(xyz_text(atoms, structures, titles))
