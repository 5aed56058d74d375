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


# ------------------------------------------------------------------------------------------------
# TFD ensemble pruning (torsion_module.py:957-1076)
# ------------------------------------------------------------------------------------------------
_TFD_SCHEDULE = (5e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5000, 2000, 1000, 500, 200, 100, 50, 20, 10, 5, 2, 1)
last_tfd_ties = None  # near-threshold decisions of the last prune_conformers_tfd call


def _get_tf_mat(structures, quadruplets):
    """Torsion fingerprint matrix (n, Q) in degrees (torsion_module.py:1046-1053), on the GPU."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    q = np.ascontiguousarray(np.asarray(quadruplets, dtype=np.int64).reshape(-1, 4))
    tf = np.zeros((len(x), len(q)), dtype=np.float64)
    _lib.check(lib.fc_tfd_fingerprints(_ptr(x), len(x), x.shape[1], _ptr(q), len(q), _ptr(tf)), "fc_tfd_fingerprints")
    return tf


def get_torsion_fingerprint(coords, quadruplets):
    """Dihedral of every quadruplet of one structure (torsion_module.py:1070-1076)."""
    return _get_tf_mat(np.asarray(coords, dtype=np.float64)[None], quadruplets)[0]


def tfd_similarity(tfp1, tfp2, thresh=10):
    """torsion_module.py:1056-1067 for one pair (two-row batch through the GPU first-match search)."""
    tf = np.ascontiguousarray(np.stack([np.asarray(tfp1, dtype=np.float64), np.asarray(tfp2, dtype=np.float64)]))
    return bool(_first_match(tf, [(0, 2)], thresh)[0][0] == 1)


def _first_match(tf, chunks, thresh, tie_cap=1 << 16):
    import ctypes as C

    from .embeds import TIE_DTYPE

    lib = _lib.load(require_device=True)
    tf = np.ascontiguousarray(tf, dtype=np.float64)
    n, nq = tf.shape
    start = np.ascontiguousarray([c[0] for c in chunks], dtype=np.int64)
    length = np.ascontiguousarray([c[1] for c in chunks], dtype=np.int64)
    first = np.full(n, -1, dtype=np.int64)
    ties = np.zeros(tie_cap, dtype=TIE_DTYPE)
    n_ties = C.c_int64(0)
    _lib.check(lib.fc_tfd_first_match(_ptr(tf), n, nq, _ptr(start), _ptr(length), len(chunks), float(thresh), _ptr(first),
                                      _ptr(ties), tie_cap, C.byref(n_ties)), "fc_tfd_first_match")
    return first, ties[: min(int(n_ties.value), tie_cap)]


def prune_conformers_tfd(structures, quadruplets, thresh=10, verbose=False):
    """Drop-in for firecode.torsion_module.prune_conformers_tfd (torsion_module.py:957-1043):
    ``(structures[mask], mask)``.  Fingerprints and the per-chunk first-match searches run on the GPU; the
    cluster resolution keeps the reference's own networkx calls (its outcome depends on set / graph
    iteration order).  Quirks reproduced: the last subdivision of a pass ends at the number of active
    structures, and masked-out structures keep taking part in later passes."""
    global last_tfd_ties
    from networkx import Graph, connected_components

    structures = np.asarray(structures)
    n = structures.shape[0]
    tf = _get_tf_mat(structures, quadruplets)
    final_mask = np.ones(n, dtype=bool)
    all_ties = []
    for k in _TFD_SCHEDULE:
        num_active_str = int(np.count_nonzero(final_mask))
        if not (k == 1 or 5 * k < num_active_str):
            continue
        if verbose:
            print(f"Working on subgroups with k={k} ({num_active_str} candidates left) {' ' * 10}", end="\r")
        d = int(n // k)
        chunks = []
        for step in range(int(k)):
            if step == k - 1:
                chunks.append((d * step, len(range(d * step, num_active_str))))
            else:
                chunks.append((d * step, len(range(d * step, int(d * (step + 1))))))
        first, ties = _first_match(tf, chunks, thresh)
        all_ties.append(ties)
        for start, length in chunks:
            if length < 2:
                continue
            matches = set()
            found = first[start:start + length]
            for i_rel in np.flatnonzero(found >= 0):  # ascending i, as the reference inserts them
                matches.add((int(i_rel), int(found[i_rel]) - start))
            if not matches:
                continue
            g = Graph(matches)
            subgraphs = [g.subgraph(c) for c in connected_components(g)]
            groups = [tuple(graph.nodes) for graph in subgraphs]
            for group in groups:
                for i in set(group) - {group[0]}:  # of each cluster, keep the first structure
                    final_mask[i + start] = False
    last_tfd_ties = np.concatenate(all_ties) if all_ties else None
    return structures[final_mask], final_mask


# ------------------------------------------------------------------------------------------------
# conformational search drivers (torsion_module.py:436-586, 726-891)
# ------------------------------------------------------------------------------------------------
def csearch_apply(starts, torsions, masks, angle_sets, thresh=1.5):
    """Inner loop of the csearch drivers (torsion_module.py:512-552 = 813-856), batched on the GPU: every
    angle set is applied, torsion by torsion with the 5-degree back-off, to every starting structure.
    Returns (coords (S, A, N, 3), rotated_bonds (S, A) int32, near (S, A) bool)."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(starts, dtype=np.float64))
    if x.ndim == 2:
        x = x[None]
    tors = np.ascontiguousarray(np.asarray(torsions, dtype=np.int32).reshape(-1, 4))
    m = np.ascontiguousarray(np.asarray(masks, dtype=bool).reshape(len(tors), -1).astype(np.uint8))
    sets = np.ascontiguousarray(np.asarray(angle_sets, dtype=np.int32).reshape(-1, max(len(tors), 1)))
    s_, n = x.shape[:2]
    a = len(sets) if len(tors) else 0
    out = np.zeros((s_, a, n, 3), dtype=np.float64)
    rotated = np.zeros((s_, a), dtype=np.int32)
    near = np.zeros((s_, a), dtype=np.uint8)
    rc = lib.fc_csearch_apply(_ptr(x), s_, n, _ptr(tors), len(tors), _ptr(m), _ptr(sets), a, float(thresh),
                              int(conventions.ROT_HANDEDNESS), int(conventions.TORSION_AXIS_SIGN), _ptr(out), _ptr(rotated),
                              _ptr(near))
    _lib.check(rc, "fc_csearch_apply")
    return out, rotated, near.astype(bool)


def most_diverse_conformers(n, structures):
    """torsion_module.py:574-586 ("TEMP: JUST RETURNS THE TOP n STRUCTURES": a sorted random choice)."""
    if len(structures) <= n:
        return list(np.array(structures))
    indices = np.sort(np.random.choice(len(structures), size=n))
    return list(np.array(structures)[indices])


def _torsion_tables(torsions, graph):
    tors = [tuple(int(i) for i in t.torsion) for t in torsions]
    return tors, [get_rotation_mask(graph, t) for t in tors]


def random_csearch(atoms, coords, torsions, graph, constrained_indices=None, n_out=100, max_tries=10000, rotations=None,
                   title="test", logfunction=print, interactive_print=True, write_torsions=False, batch=4096):
    """Drop-in for firecode.torsion_module.random_csearch (torsion_module.py:436-571): the shuffled angle
    sets are applied in batches on the GPU and consumed in the reference's order, including its stopping
    rule (the `a == max_tries` test only fires when a structure is appended)."""
    from .utils import cartesian_product

    if logfunction is not None:
        logfunction(f"\n--> Random dihedral CSearch on {title}\n    mode 2 (random) - {len(torsions)} torsions")
    angles = cartesian_product(*[t.get_angles() for t in torsions])
    if rotations is not None:
        angles = angles[np.count_nonzero(angles, axis=1) == rotations]
    np.random.shuffle(angles)  # same call as the reference: identical RNG state -> identical order
    tors, masks = _torsion_tables(torsions, graph)
    coords = np.asarray(coords, dtype=np.float64)
    new_structures = []
    done = False
    for lo in range(0, len(angles), batch):
        out, rotated, _ = csearch_apply(coords, tors, masks, angles[lo:lo + batch])
        for k in np.flatnonzero(rotated[0] != 0):
            new_structures.append(out[0, k])
            if len(new_structures) == n_out or lo + k == max_tries:
                done = True
                break
        if done:
            break
    if logfunction is not None:
        exhaustiveness = len(new_structures) / np.prod([t.n_fold for t in torsions])
        logfunction(f"  Generated {len(new_structures)} conformers, (est. {round(100 * exhaustiveness, 2)} % of the total "
                    "conformational space)")
    return np.array(new_structures)


def clustered_csearch(atoms, coords, torsions, graph, charge=0, mult=1, constrained_indices=None, n=100, n_out=100,
                      title="test", logfunction=print, interactive_print=True, write_torsions=False, debug=False):
    """Drop-in for firecode.torsion_module.clustered_csearch (torsion_module.py:726-891): all angle sets of
    the (single) torsion group are applied to every starting point in one GPU call, the starting point itself
    leads its block as in the reference, then TFD pruning (GPU) and the reference's "most diverse" cut."""
    from .utils import cartesian_product

    grouped_torsions = [torsions]
    if logfunction is not None:
        logfunction(f"\n--> Clustered CSearch on {title}\n    - {len(torsions)} torsions in 1 group - {[len(torsions)]}")
    output_structures = []
    starting_points = [np.asarray(coords, dtype=np.float64)]
    new_structures = []
    for tg, torsions_group in enumerate(grouped_torsions):
        angles = cartesian_product(*[t.get_angles() for t in torsions_group])
        tors, masks = _torsion_tables(torsions_group, graph)
        out, rotated, _ = csearch_apply(np.array(starting_points), tors, masks, angles)
        new_structures = []
        for s_, sp in enumerate(starting_points):
            new_structures.append(sp)
            for k in np.flatnonzero(rotated[s_] != 0):
                new_structures.append(out[s_, k])
        torsion_array = np.array([t.torsion for t in torsions])
        if tg + 1 != len(grouped_torsions) and n is not None and len(new_structures) > n:
            new_structures = most_diverse_conformers(n, new_structures)
        output_structures.extend(new_structures)
        starting_points = new_structures
    output_structures = list(prune_conformers_tfd(np.array(output_structures), torsion_array)[0])
    if len(new_structures) > n_out:
        output_structures = most_diverse_conformers(n_out, output_structures)
    if logfunction is not None:
        exhaustiveness = len(output_structures) / np.prod([t.n_fold for t in torsions])
        logfunction(f"  Selected the most diverse {len(output_structures)} conformers, corresponding\n"
                    f"  to about {round(100 * exhaustiveness, 2)} % of the total conformational space")
    return np.array(output_structures)

