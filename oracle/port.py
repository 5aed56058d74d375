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


def pose7_to_xf(pose7):
    """Compact poses (n, 7) float32 {qx, qy, qz, qw, tx, ty, tz} -> (n, 12) float64 transforms, the expansion
    include/firecode_b200.h defines for FC_POSE_Q7 (every operation rounded on its own, in that order)."""
    p = np.asarray(pose7, dtype=np.float32).reshape(-1, 7).astype(np.float64)
    x, y, z, w = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    xx, yy, zz, ww = x * x, y * y, z * z, w * w
    n = ((xx + yy) + zz) + ww
    s = 2.0 / n
    xy, xz, yz, xw, yw, zw = x * y, x * z, y * z, x * w, y * w, z * w
    xf = np.empty((len(p), 12))
    xf[:, 0] = 1.0 - s * (yy + zz)
    xf[:, 1] = s * (xy - zw)
    xf[:, 2] = s * (xz + yw)
    xf[:, 3] = s * (xy + zw)
    xf[:, 4] = 1.0 - s * (xx + zz)
    xf[:, 5] = s * (yz - xw)
    xf[:, 6] = s * (xz - yw)
    xf[:, 7] = s * (yz + xw)
    xf[:, 8] = 1.0 - s * (xx + yy)
    xf[:, 9:12] = p[:, 4:7]
    return xf


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


# ------------------------------------------------------------------------------------------------
# shared helpers (reference arithmetic, float64)
# ------------------------------------------------------------------------------------------------
def _shim():
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import prism_pruner.algebra as alg
    import prism_pruner.rmsd as rmsd

    return alg, rmsd


def rotation_matrix_from_vectors(vec1, vec2):
    """utils.py:224-249."""
    alg, _ = _shim()
    a = vec1 / np.linalg.norm(vec1)
    b = vec2 / np.linalg.norm(vec2)
    v = np.cross(a, b)
    if np.linalg.norm(v) != 0:
        c = np.dot(a, b)
        s = np.linalg.norm(v)
        kmat = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        return np.eye(3) + kmat + kmat.dot(kmat) * ((1 - c) / (s**2))
    if np.linalg.norm(a + b) == 0:
        return alg.rot_mat_from_pointer(np.array([0, 0, 1]), 180)
    return np.eye(3)


def align_vec_pair(ref, tgt):
    """algebra.py:28-49: rotation taking the two tgt vectors onto the two ref vectors."""
    B = np.zeros((3, 3))
    for j in range(2):
        B += np.outer(ref[j], tgt[j])
    u, s, vh = np.linalg.svd(B)
    if np.linalg.det(u @ vh) < 0:
        u[:, -1] = -u[:, -1]
    return np.ascontiguousarray(u @ vh)


def torsion_fingerprint(coords, quadruplets):
    """torsion_module.py:1070-1076."""
    alg, _ = _shim()
    out = np.zeros(len(quadruplets))
    for i, (i1, i2, i3, i4) in enumerate(quadruplets):
        out[i] = alg.dihedral([coords[i1], coords[i2], coords[i3], coords[i4]])
    return out


def tfd_sum(tfp1, tfp2):
    """Sum of wrapped absolute torsion differences, torsion_module.py:1056-1067."""
    deltas = np.abs(tfp1 - tfp2)
    deltas = np.abs(deltas - (deltas > 180) * 360)
    return float(np.sum(deltas))


class Ties:
    """Near-threshold bookkeeping shared by the ports.

    A decision whose value lies within ``eps`` of its threshold is recorded in ``seen`` and, when
    ``forced`` holds an entry for the same key, takes that decision instead of its own (this is
    how a parity test conditions the oracle on the near-threshold decisions listed by the GPU)."""

    def __init__(self, eps=0.0, forced=None):
        self.eps = eps
        self.forced = forced or {}
        self.seen = {}

    def decide(self, key, value, threshold, op):
        own = bool(op(value, threshold))
        if self.eps > 0 and abs(value - threshold) <= self.eps:
            self.seen[key] = (value, own)
            if key in self.forced:
                return bool(self.forced[key])
        return own


# ------------------------------------------------------------------------------------------------
# string embed -- firecode/embeds.py:51-158
# ------------------------------------------------------------------------------------------------
def string_transform(prob, c1, c2, a1, a2, angle):
    """Rotation / position of molecule 2 for one tuple, embeds.py:121-135."""
    alg, _ = _shim()
    p1, p2 = prob.centers[0][c1][a1], prob.centers[1][c2][a2]
    ref_vec, mol_vec = prob.vecs[0][c1][a1], prob.vecs[1][c2][a2]
    rot = rotation_matrix_from_vectors(mol_vec, -ref_vec)
    if angle != 0:
        rot = alg.rot_mat_from_pointer(ref_vec, angle) @ rot
    return rot, p1 - rot @ p2


def string_embed(prob, ties=None, want_poses=True):
    """Reference loop of string_embed on a StringProblem (firecode_b200.problem).

    Returns dict: kept (indices of accepted tuples in enumeration order), poses (P, N1+N2, 3),
    clash_pass (bool per tuple), dmin (per tuple), ties (Ties)."""
    import operator

    ties = ties or Ties()
    n1 = prob.coords[0].shape[1]
    n2 = prob.coords[1].shape[1]
    accepted_tfp, accepted_idx, poses = [], [], []
    n = prob.n_poses
    clash_pass = np.zeros(n, dtype=bool)
    dmin = np.zeros(n)
    for pose in range(n):
        c1, c2, a1, a2, angle = prob.decode(pose)
        rot, pos = string_transform(prob, c1, c2, a1, a2, angle)
        structure = np.concatenate([prob.coords[0][c1], (rot @ prob.coords[1][c2].T).T + pos])
        d = cdist(structure[n1:], structure[:n1])
        dmin[pose] = d.min()
        # utils.py:551 with max_clashes = 0: no pair strictly below the threshold
        ok = not ties.decide(("clash", pose), float(d.min()), prob.thresh, operator.lt)
        clash_pass[pose] = ok
        if not ok:
            continue
        tfp = torsion_fingerprint(structure, prob.quadruplets)
        new = True
        for ref_pose, ref in zip(accepted_idx, accepted_tfp):  # never-evicting cache (quirk N2)
            if ties.decide(("tfd", pose, ref_pose), tfd_sum(tfp, ref), 10.0, operator.lt):
                new = False
                break
        if new:
            accepted_tfp.append(tfp)
            accepted_idx.append(pose)
            if want_poses:
                poses.append(structure)
    return {"kept": np.array(accepted_idx, dtype=np.int64),
            "poses": np.array(poses).reshape(len(poses), n1 + n2, 3) if want_poses else None,
            "clash_pass": clash_pass, "dmin": dmin, "ties": ties}


# ------------------------------------------------------------------------------------------------
# cyclical embed, bimolecular fast path -- firecode/embeds.py:588-750
# ------------------------------------------------------------------------------------------------
def polygonize(lengths):
    """utils.py:252-312 (two-segment case and triangle case)."""
    from firecode_b200.utils import polygonize as _poly  # same closed form, pinned in tests

    return _poly(lengths)


def cyclical_reactive_indices(pivot_ids, n, n_mols):
    """embeds.py:753-784: atom couples facing each other for orientation ``n``."""
    def orient(i, ids):
        return list(reversed(ids)) if swaps[n][i] else list(ids)

    if n_mols == 2:
        swaps = [(0, 0), (0, 1)]
        o = [orient(i, ids) for i, ids in enumerate(pivot_ids)]
        return [(int(o[0][0]), int(o[1][0])), (int(o[0][1]), int(o[1][1]))]
    swaps = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)]
    o = [orient(i, ids) for i, ids in enumerate(pivot_ids)]
    couples = [(o[0][1], o[1][0]), (o[1][1], o[2][0]), (o[2][1], o[0][0])]
    return [tuple(sorted((int(a), int(b)))) for a, b in couples]


def pairing_filter(prob, ids):
    """embeds.py:473-476 / 638-641 including quirk N10 (tuple_in_collection never matches an
    ndarray collection)."""
    if not prob.pairings:
        return True
    def in_internal(pair):
        if prob.internal_constraints_is_array:
            return False  # list-of-lists never equals a tuple (embeds.py:820-826)
        return tuple(pair) in [tuple(x) for x in prob.internal_constraints]
    return all((tuple(pair) in [tuple(i) for i in ids]) or in_internal(pair) for pair in prob.pairings)


def cyclical_molecule_transform(prob, i, conf, pivot_vec, pivot_mean, vec_pair, direction, angle):
    """Rotation / position of molecule i for one angle, embeds.py:494-554 (= 649-709)."""
    alg, _ = _shim()
    start, end = vec_pair
    reactive_coords = prob.coords[i][conf][prob.reactive[i]]
    atomic_pivot_mean = np.mean(reactive_coords, axis=0)
    mol_direction = pivot_mean - atomic_pivot_mean
    if np.all(mol_direction == 0.0):
        mol_direction = pivot_mean
    alignment = align_vec_pair(np.array([end - start, direction]), np.array([pivot_vec, mol_direction]))
    if len(reactive_coords) == 2:
        axis = alignment @ (reactive_coords[0] - reactive_coords[1])
    else:
        axis = alignment @ pivot_vec
    step = alg.rot_mat_from_pointer(axis, angle)
    center = alignment @ atomic_pivot_mean
    rotation = step @ alignment
    pos = np.mean(vec_pair, axis=0) - alignment @ pivot_mean
    position = center - step @ center + pos
    return rotation, position


def rmsd_and_max(p, q):
    """prism_pruner.rmsd.rmsd_and_max(p, q, center=False) as called from utils.py:499."""
    _, rmsd = _shim()
    return rmsd.rmsd_and_max(p, q, center=False)


def cyclical_groups_bimol(prob):
    """Valid (conformer pair, pivot pair, orientation) groups in the reference's loop order.
    Returns list of dicts: conf (c0, c1), piv (p0, p1), v, ids."""
    from firecode_b200.utils import cartesian_product

    groups = []
    n_conf = [len(c) for c in prob.coords]
    for conf_ids in cartesian_product(*[np.arange(n) for n in n_conf]):
        counts = [len(prob.pivot_vec[m][conf_ids[m]]) for m in range(2)]
        if min(counts) == 0:
            continue
        for pi in cartesian_product(*[np.arange(k) for k in counts]):
            pv = [prob.pivot_vec[m][conf_ids[m]][pi[m]] for m in range(2)]
            norms = np.linalg.norm(np.array(pv), axis=1)
            if abs(norms[0] - norms[1]) > prob.max_norm_delta:  # embeds.py:624
                continue
            for v in range(2):
                ids = cyclical_reactive_indices([prob.pivot_ids[m][conf_ids[m]][pi[m]] for m in range(2)], v, 2)
                if pairing_filter(prob, ids):
                    groups.append({"conf": tuple(int(c) for c in conf_ids), "piv": tuple(int(p) for p in pi),
                                   "v": v, "ids": ids, "norms": norms})
    return groups


def cyclical_embed_bimol(prob, ties=None, rmsd_thr=1.0, want_poses=True):
    """Reference loop of _fast_bimol_rigid_cyclical_embed on a CyclicalProblem.

    Pose index = (index of the group in cyclical_groups_bimol order) * n_angles + angle index."""
    import operator

    ties = ties or Ties()
    assert prob.n_mols == 2
    groups = cyclical_groups_bimol(prob)
    n_ang = len(prob.angles)
    n0 = prob.coords[0].shape[1]
    directions = np.array([[0, 1, 0], [0, -1, 0]])
    kept, poses, constrained = [], [], []
    clash_pass = np.zeros(len(groups) * n_ang, dtype=bool)
    for g, grp in enumerate(groups):
        c = grp["conf"]
        vecs = polygonize(grp["norms"])[grp["v"]]
        angular, angular_idx = [], []
        for ai, angles in enumerate(prob.angles):
            parts = []
            for i in range(2):
                rot, pos = cyclical_molecule_transform(
                    prob, i, c[i], prob.pivot_vec[i][c[i]][grp["piv"][i]], prob.pivot_mean[i][c[i]][grp["piv"][i]],
                    vecs[i], directions[i], angles[i])
                parts.append((rot @ prob.coords[i][c[i]].T).T + pos)
            structure = np.concatenate(parts)
            pose = g * n_ang + ai
            d = cdist(structure[n0:], structure[:n0])
            ok = not ties.decide(("clash", pose), float(d.min()), prob.thresh, operator.lt)
            clash_pass[pose] = ok
            if not ok:
                continue
            similar = False
            for ref_pose, ref in zip(angular_idx, angular):  # utils.py:494-504
                r, m = rmsd_and_max(structure, ref)
                if ties.decide(("rmsd", pose, ref_pose), r, rmsd_thr, operator.lt) and \
                        ties.decide(("maxdev", pose, ref_pose), m, 2 * rmsd_thr, operator.lt):
                    similar = True
                    break
            if not similar:
                angular.append(structure)
                angular_idx.append(pose)
                kept.append(pose)
                constrained.append(grp["ids"])
                if want_poses:
                    poses.append(structure)
    n_tot = sum(c.shape[1] for c in prob.coords)
    return {"kept": np.array(kept, dtype=np.int64), "groups": groups,
            "poses": np.array(poses).reshape(len(poses), n_tot, 3) if want_poses else None,
            "constrained": np.array(constrained, dtype=np.int64).reshape(len(kept), 2, 2),
            "clash_pass": clash_pass, "ties": ties}


# ------------------------------------------------------------------------------------------------
# cyclical embed, trimolecular body -- firecode/embeds.py:188-585
# ------------------------------------------------------------------------------------------------
def triangle_vertices(norms):
    """Planar triangle (0,0), (l0,0), (x,y) used by embeds.py:198-209 and 288-299."""
    vertices = np.zeros((3, 2))
    vertices[1] = np.array([norms[0], 0])
    a = np.power(norms[0], 2)
    b = np.power(norms[1], 2)
    c = np.power(norms[2], 2)
    x = (a - b + c) / (2 * a**0.5)
    y = (c - x**2) ** 0.5
    vertices[2] = np.array([x, y])
    return vertices


def get_directions3(norms):
    """embeds.py:188-254 for three molecules.  ``norms`` is mutated in place in the right-triangle
    case (embeds.py:232-237, quirk N7) exactly as the reference does."""
    alg, _ = _shim()
    vertices = triangle_vertices(norms)
    a = vertices[1, 0]
    b = vertices[2, 0]
    c = vertices[2, 1]
    x = a / 2
    y = (b**2 + c**2 - a * b) / (2 * c)
    cc = np.array([x, y])
    v0, v1, v2 = vertices
    meanpoint1 = np.mean((v0, v1), axis=0)
    meanpoint2 = np.mean((v1, v2), axis=0)
    meanpoint3 = np.mean((v2, v0), axis=0)
    dir1 = cc - meanpoint1
    dir2 = cc - meanpoint2
    dir3 = cc - meanpoint3
    if np.any([np.all(d == 0) for d in (dir1, dir2, dir3)]):
        norms[0] += 1e-5
        dir1, dir2, dir3 = [t[:-1] for t in get_directions3(norms)]
    angle0_obtuse = alg.vec_angle(v1 - v0, v2 - v0) > 90
    angle1_obtuse = alg.vec_angle(v0 - v1, v2 - v1) > 90
    angle2_obtuse = alg.vec_angle(v0 - v2, v1 - v2) > 90
    dir1 = -dir1 if angle2_obtuse else dir1
    dir2 = -dir2 if angle0_obtuse else dir2
    dir3 = -dir3 if angle1_obtuse else dir3
    dir1 = alg.normalize(np.concatenate((dir1, [0])))
    dir2 = alg.normalize(np.concatenate((dir2, [0])))
    dir3 = alg.normalize(np.concatenate((dir3, [0])))
    return np.vstack((dir1, dir2, dir3))


def facing_table(prob, constrained_indices):
    """r[m, partner] = index (in molecule m) of the reactive atom facing ``partner``
    (embeds.py:328-353; matched through the cumnum of conformer 0's reactive atoms)."""
    pairings = [[(-1, -1), (-1, -1)] for _ in constrained_indices]
    for i, c in enumerate(constrained_indices):
        for m in range(prob.n_mols):
            for index, cumnum in prob.ratoms0[m]:
                if cumnum == c[0]:
                    pairings[i][0] = (m, int(index))
                if cumnum == c[1]:
                    pairings[i][1] = (m, int(index))
    r = np.zeros((3, 3), dtype=int)
    for first, second in pairings:
        r[first[0], second[0]] = first[1]
        r[second[0], first[0]] = second[1]
    return r


def adjust_directions(prob, norms, directions, constrained_indices, triangle_vectors, pivot_vec, pivot_mean,
                      conf_ids, choice=None):
    """embeds.py:256-407.  Returns (directions (3,3), index of the chosen candidate, cost gap to the
    runner-up).  ``choice`` forces the candidate (parity tests condition the oracle on a listed
    near-tie of the 343-point grid search)."""
    alg, _ = _shim()
    from firecode_b200.utils import cartesian_product

    p0, p1, p2 = [end - start for start, end in triangle_vectors]
    p0_mean, p1_mean, p2_mean = [np.mean((end, start), axis=0) for start, end in triangle_vectors]
    vertices = triangle_vertices(norms)
    v0, v1, v2 = [np.concatenate((v, [0])) for v in vertices]
    rot, pos = [], []
    for i in (0, 1, 2):
        start, end = triangle_vectors[i]
        mol_direction = pivot_mean[i] - np.mean(prob.coords[i][conf_ids[i]][prob.reactive[i]], axis=0)
        if np.all(mol_direction == 0.0):
            mol_direction = pivot_mean[i]
        rot.append(align_vec_pair(np.array([end - start, directions[i]]), np.array([pivot_vec[i], mol_direction])))
        pos.append(np.mean(triangle_vectors[i], axis=0) - rot[i] @ pivot_mean[i])
    r = facing_table(prob, constrained_indices)
    # reactive atom positions are read from CONFORMER 0 (embeds.py:359-366, quirk N6)
    a01 = rot[0] @ prob.coords[0][0][r[0, 1]] + pos[0]
    a02 = rot[0] @ prob.coords[0][0][r[0, 2]] + pos[0]
    a10 = rot[1] @ prob.coords[1][0][r[1, 0]] + pos[1]
    a12 = rot[1] @ prob.coords[1][0][r[1, 2]] + pos[1]
    a20 = rot[2] @ prob.coords[2][0][r[2, 0]] + pos[2]
    a21 = rot[2] @ prob.coords[2][0][r[2, 1]] + pos[2]
    steps = 6
    angle_range = 30
    step_angle = 2 * angle_range / steps
    angles_list = cartesian_product(*[range(steps + 1) for _ in range(3)]) * step_angle - angle_range
    costs, dirs = [], []
    for angles in angles_list:
        rot0 = alg.rot_mat_from_pointer(p0, angles[0])
        new_a01 = rot0 @ a01
        new_a02 = rot0 @ a02
        d0 = p0_mean - np.mean((new_a01, new_a02), axis=0)
        rot1 = alg.rot_mat_from_pointer(p1, angles[1])
        new_a10 = rot1 @ a10
        new_a12 = rot1 @ a12
        d1 = p1_mean - np.mean((new_a10, new_a12), axis=0)
        rot2 = alg.rot_mat_from_pointer(p2, angles[2])
        new_a20 = rot2 @ a20
        new_a21 = rot2 @ a21
        d2 = p2_mean - np.mean((new_a20, new_a21), axis=0)
        cost = 0
        cost += alg.vec_angle(v0 - new_a02, new_a20 - v0)
        cost += alg.vec_angle(v1 - new_a01, new_a10 - v1)
        cost += alg.vec_angle(v2 - new_a21, new_a12 - v2)
        costs.append(cost)
        dirs.append((d0, d1, d2))
    order = sorted(range(len(costs)), key=lambda k: costs[k])  # stable: first minimum wins (embeds.py:405)
    best = order[0]
    gap = costs[order[1]] - costs[best]
    if choice is not None:
        best = int(choice)
    return np.array(dirs[best]), best, float(gap)


def cyclical_groups_trimol(prob, conf_tuple_range=None):
    """(conformer triple, pivot triple) super-groups that form a triangle, in the reference's loop
    order (embeds.py:414-462), each with its 8 orientations' atom couples and pairing-filter flags."""
    from firecode_b200.utils import cartesian_product

    out = []
    n_conf = [len(c) for c in prob.coords]
    for t, conf_ids in enumerate(cartesian_product(*[np.arange(n) for n in n_conf])):
        if conf_tuple_range is not None and not (conf_tuple_range[0] <= t < conf_tuple_range[1]):
            continue
        counts = [len(prob.pivot_vec[m][conf_ids[m]]) for m in range(3)]
        if min(counts) == 0:
            continue
        for pi in cartesian_product(*[np.arange(k) for k in counts]):
            pv = [prob.pivot_vec[m][conf_ids[m]][pi[m]] for m in range(3)]
            norms = np.linalg.norm(np.array(pv), axis=1)
            if not all(norms[i] < norms[i - 1] + norms[i - 2] for i in (0, 1, 2)):  # embeds.py:447
                continue
            pid = [prob.pivot_ids[m][conf_ids[m]][pi[m]] for m in range(3)]
            ids = [cyclical_reactive_indices(pid, v, 3) for v in range(8)]
            out.append({"conf": tuple(int(c) for c in conf_ids), "piv": tuple(int(p) for p in pi),
                        "norms": norms, "ids": ids, "active": [pairing_filter(prob, i) for i in ids]})
    return out


def cyclical_embed_trimol(prob, ties=None, rmsd_thr=1.0, want_poses=True, forced_choice=None, choice_eps=0.0,
                          conf_tuple_range=None):
    """Reference loop of cyclical_embed for three molecules (embeds.py:409-585) on a CyclicalProblem.

    Group = (super-group, orientation v) passing the pairing filter, numbered in loop order;
    pose index = group * n_angles + angle index.  ``forced_choice`` {group: candidate} overrides the
    grid-search argmin of _adjust_directions where its runner-up gap is below ``choice_eps``."""
    import operator

    ties = ties or Ties()
    assert prob.n_mols == 3
    supers = cyclical_groups_trimol(prob, conf_tuple_range)
    n_ang = len(prob.angles)
    kept, poses, constrained, groups, clash_pass = [], [], [], [], []
    near_choice = {}
    for sg in supers:
        c = sg["conf"]
        pv = [prob.pivot_vec[m][c[m]][sg["piv"][m]] for m in range(3)]
        pm = [prob.pivot_mean[m][c[m]][sg["piv"][m]] for m in range(3)]
        norms = sg["norms"].copy()
        polygon_vectors = polygonize(norms)          # embeds.py:453 (before any perturbation of norms)
        directions = get_directions3(norms)          # may perturb norms[0] in place (N7)
        for v in range(8):
            if not sg["active"][v]:
                continue
            g = len(groups)
            vecs = polygon_vectors[v]
            ids = sg["ids"][v]
            force = None
            directions_in = directions
            directions, best, gap = adjust_directions(prob, norms, directions_in, ids, vecs, pv, pm, c)
            if gap <= choice_eps:
                near_choice[g] = (best, gap)
                if forced_choice is not None and g in forced_choice and forced_choice[g] != best:
                    force = forced_choice[g]
                    directions, best, gap = adjust_directions(prob, norms, directions_in, ids, vecs, pv, pm, c,
                                                              choice=force)
            groups.append({"conf": c, "piv": sg["piv"], "v": v, "ids": ids, "directions": directions.copy(),
                           "choice": best, "gap": gap})
            angular, angular_idx = [], []
            for ai, angles in enumerate(prob.angles):
                parts = []
                for i in range(3):
                    rot, pos = cyclical_molecule_transform(prob, i, c[i], pv[i], pm[i], vecs[i], directions[i],
                                                           angles[i])
                    parts.append((rot @ prob.coords[i][c[i]].T).T + pos)
                structure = np.concatenate(parts)
                pose = g * n_ang + ai
                n1, n2 = prob.ids[0], prob.ids[0] + prob.ids[1]
                m1, m2, m3 = structure[:n1], structure[n1:n2], structure[n2:]
                ok = True
                for blk, (x, y) in enumerate(((m2, m1), (m3, m2), (m1, m3))):   # utils.py:563-571
                    d = float(cdist(x, y).min())
                    if ties.decide(("clash3", pose, blk), d, prob.thresh, operator.le):
                        ok = False
                        break
                clash_pass.append(ok)
                if not ok:
                    continue
                similar = False
                for ref_pose, ref in zip(angular_idx, angular):  # utils.py:494-504
                    r, m = rmsd_and_max(structure, ref)
                    if ties.decide(("rmsd", pose, ref_pose), r, rmsd_thr, operator.lt) and \
                            ties.decide(("maxdev", pose, ref_pose), m, 2 * rmsd_thr, operator.lt):
                        similar = True
                        break
                if not similar:
                    angular.append(structure)
                    angular_idx.append(pose)
                    kept.append(pose)
                    constrained.append(ids)
                    if want_poses:
                        poses.append(structure)
    n_tot = sum(c.shape[1] for c in prob.coords)
    return {"kept": np.array(kept, dtype=np.int64), "groups": groups,
            "poses": np.array(poses).reshape(len(poses), n_tot, 3) if want_poses else None,
            "constrained": np.array(constrained, dtype=np.int64).reshape(len(kept), 3, 2),
            "clash_pass": np.array(clash_pass, dtype=bool), "ties": ties, "near_choice": near_choice}


# ------------------------------------------------------------------------------------------------
# torsion rotation + clash -- torsion_module.py:354-382, 894-918 and prism_pruner.utils.rotate_dihedral
# ------------------------------------------------------------------------------------------------
def rotation_mask(graph, torsion):
    """torsion_module.py:354-382 (the graph is left unchanged)."""
    from networkx import shortest_path

    _, i2, i3, i4 = torsion
    graph.remove_edge(i2, i3)
    reachable = shortest_path(graph, i4).keys()
    graph.add_edge(i2, i3)
    mask = np.array([i in reachable for i in graph.nodes], dtype=bool)
    mask[i3] = False
    return mask


def torsion_comp_check(coords, torsion, mask, thresh=1.5, max_clashes=0):
    """torsion_module.py:894-918."""
    _, i2, i3, _ = torsion
    antimask = ~mask
    antimask[i2] = False
    antimask[i3] = False
    d = cdist(coords[antimask], coords[mask])
    return int(np.count_nonzero(d < thresh)) <= max_clashes, (float(d.min()) if d.size else np.inf)


def torsion_scan(coords, torsions, masks, angles, thresh=1.5, max_clashes=0):
    """rotate_dihedral + torsion_comp_check over (conformer, torsion, angle)."""
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    from prism_pruner.utils import rotate_dihedral

    coords = np.asarray(coords, dtype=float)
    c, n = coords.shape[:2]
    out = np.zeros((c, len(torsions), len(angles), n, 3))
    passed = np.zeros((c, len(torsions), len(angles)), dtype=bool)
    dmin = np.zeros((c, len(torsions), len(angles)))
    for ci in range(c):
        for ti, (tor, mask) in enumerate(zip(torsions, masks)):
            for ai, ang in enumerate(angles):
                new = rotate_dihedral(coords[ci], tuple(int(t) for t in tor), ang, mask=np.asarray(mask, dtype=bool))
                out[ci, ti, ai] = new
                passed[ci, ti, ai], dmin[ci, ti, ai] = torsion_comp_check(new, tor, np.asarray(mask, dtype=bool), thresh, max_clashes)
    return out, passed, dmin


# ------------------------------------------------------------------------------------------------
# TFD ensemble pruning -- firecode/torsion_module.py:957-1067 (called from embedder.py:1430-1437, the
# `tfd` branch of similarity_refining, and torsion_module.py:875)
# ------------------------------------------------------------------------------------------------
TFD_SCHEDULE = (5e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5000, 2000, 1000, 500, 200, 100, 50, 20, 10, 5, 2, 1)


def tf_mat(structures, quadruplets):
    """torsion_module.py:1046-1053."""
    return np.array([torsion_fingerprint(s, quadruplets) for s in structures]).reshape(len(structures), len(quadruplets))


def tfd_chunks(n, k, num_active):
    """(start, length) of the k subdivisions of a pass (torsion_module.py:984-991): d = n // k structures
    each, the LAST one ends at the number of still-active structures (not at n) -- reproduced as is."""
    d = int(n // k)
    out = []
    for step in range(int(k)):
        if step == k - 1:
            length = len(range(d * step, num_active))
        else:
            length = len(range(d * step, int(d * (step + 1))))
        out.append((d * step, length))
    return out


def tfd_resolve_chunk(matches):
    """torsion_module.py:1022-1037: clusters of the match graph, the first node of each cluster (in
    networkx's iteration order of the sub-graph) survives.  Returns the rejected relative indices."""
    from networkx import Graph, connected_components

    g = Graph(matches)
    subgraphs = [g.subgraph(c) for c in connected_components(g)]
    groups = [tuple(graph.nodes) for graph in subgraphs]
    best_of_cluster = [group[0] for group in groups]
    rejects = []
    for members, best in zip(groups, best_of_cluster):
        for i in set(members) - {best}:
            rejects.append(i)
    return rejects


def prune_conformers_tfd(structures, quadruplets, thresh=10, ties=None, first_match=None):
    """Reference driver of the TFD pruning on plain arrays.  Per chunk every structure i is matched with
    the FIRST later structure of the chunk whose torsion-difference sum is below ``thresh`` (the cache of
    known-dissimilar pairs only saves work); masked-out structures still take part (the loops never look
    at the mask).  ``first_match(tf, chunks) -> (n,) int64`` (absolute index of the first match or -1)
    replaces the pair loops when given (the CUDA path plugs in here in the parity tests)."""
    import operator

    ties = ties or Ties()
    structures = np.asarray(structures, dtype=float)
    n = len(structures)
    tf = tf_mat(structures, quadruplets)
    final_mask = np.ones(n, dtype=bool)
    for k in TFD_SCHEDULE:
        num_active = int(np.count_nonzero(final_mask))
        if not (k == 1 or 5 * k < num_active):
            continue
        chunks = tfd_chunks(n, k, num_active)
        first = None if first_match is None else first_match(tf, chunks)
        for start, length in chunks:
            matches = set()
            for i_rel in range(length):
                i_abs = i_rel + start
                if first is not None:
                    if first[i_abs] >= 0:
                        matches.add((i_rel, int(first[i_abs]) - start))
                    continue
                for j_rel in range(i_rel + 1, length):
                    j_abs = j_rel + start
                    if ties.decide(("tfd", j_abs, i_abs), tfd_sum(tf[i_abs], tf[j_abs]), float(thresh), operator.lt):
                        matches.add((i_rel, j_rel))
                        break
            for i in tfd_resolve_chunk(matches):
                final_mask[i + start] = False
    return structures[final_mask], final_mask


# ------------------------------------------------------------------------------------------------
# csearch inner loop -- firecode/torsion_module.py:512-552 (random_csearch) = 813-856 (clustered_csearch)
# ------------------------------------------------------------------------------------------------
def csearch_apply(start, torsions, masks, angle_set, thresh=1.5):
    """One angle set applied to one starting structure.  Returns (coords, rotated_bonds, closest) where
    closest is the smallest |d - thresh| over every clash check made on the way."""
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    from prism_pruner.utils import rotate_dihedral

    new_coords = np.copy(np.asarray(start, dtype=float))
    rotated_bonds = 0
    closest = np.inf

    def check(x, tor, mask):
        nonlocal closest
        _, i2, i3, _ = tor
        antimask = ~mask
        antimask[i2] = False
        antimask[i3] = False
        d = cdist(x[antimask], x[mask])
        if d.size:
            closest = min(closest, float(np.abs(d - thresh).min()))
        return int(np.count_nonzero(d < thresh)) <= 0   # torsion_module.py:918

    for t, tor in enumerate(torsions):
        angle = int(angle_set[t])
        if angle == 0:
            continue
        mask = np.asarray(masks[t], dtype=bool)
        tor = tuple(int(i) for i in tor)
        temp = rotate_dihedral(new_coords, tor, angle, mask=mask)
        if not check(temp, tor, mask):
            for _ in range(angle // 5):
                temp = rotate_dihedral(temp, tor, -5, mask=mask)
                if check(temp, tor, mask):
                    rotated_bonds += 1
                    break
        else:
            rotated_bonds += 1
        new_coords = temp
    return new_coords, rotated_bonds, closest



# ------------------------------------------------------------------------------------------------
# bond-graph post-filters (firecode/utils.py:341-400)
# ------------------------------------------------------------------------------------------------
def bond_set(atoms, coords):
    """Sorted bonds of ``graphize(atoms, coords)`` (prism_pruner shim) without self loops."""
    _shim()  # puts the prism_pruner shim on sys.path
    from prism_pruner.graph_manipulations import graphize

    return {tuple(sorted((int(a), int(b)))) for a, b in graphize(atoms, coords).edges if a != b}


def assembly_bonds(mols_graphs):
    """utils.py:371-377: the union of the fragments' bonds, shifted by the atoms before each fragment."""
    bonds, pos = set(), 0
    for graph in mols_graphs:
        for a, b in graph.edges:
            if a != b:
                bonds.add(tuple(sorted((int(a) + pos, int(b) + pos))))
        pos += len(graph.nodes)
    return bonds


def scramble_delta(atoms, structure, excluded_atoms, mols_graphs):
    """The bonds scramble_check counts (utils.py:379-391): symmetric difference of expected and found bonds minus the
    ones touching an excluded atom."""
    bonds, new_bonds = assembly_bonds(mols_graphs), bond_set(atoms, structure)
    delta = (bonds | new_bonds) - (bonds & new_bonds)
    excluded = {int(a) for a in excluded_atoms}
    return {b for b in delta if not (b[0] in excluded or b[1] in excluded)}


def scramble_check(atoms, structure, excluded_atoms, mols_graphs, max_newbonds=0):
    return len(scramble_delta(atoms, structure, excluded_atoms, mols_graphs)) <= max_newbonds


def molecule_delta(atoms, old_coords, new_coords):
    old_bonds, new_bonds = bond_set(atoms, old_coords), bond_set(atoms, new_coords)
    return (old_bonds | new_bonds) - (old_bonds & new_bonds)


def molecule_check(atoms, old_coords, new_coords, max_newbonds=0):
    """utils.py:341-353."""
    return len(molecule_delta(atoms, old_coords, new_coords)) <= max_newbonds


def bond_near_threshold(atoms, coords, eps=1e-6):
    """Number of atom pairs whose distance lies within eps of their bonding limit."""
    _shim()
    from prism_pruner.graph_manipulations import d_min_bond

    atoms = np.asarray(atoms)
    coords = np.asarray(coords, dtype=float)
    d = np.sqrt(((coords[:, None] - coords[None]) ** 2).sum(-1))
    n = len(atoms)
    return sum(1 for i in range(n) for j in range(i + 1, n) if abs(d[i, j] - d_min_bond(atoms[i], atoms[j])) <= eps)
