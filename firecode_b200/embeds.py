"""Embed entry points with the reference's names and calling convention
(firecode/embeds.py: ``fn(embedder) -> ndarray (P, N, 3)``, sets ``embedder.constrained_indices``,
logs through ``embedder.log`` and raises ``ZeroCandidatesError`` when nothing survives).

All per-pose work (transforms, clash screen, similarity filters, coordinates of the survivors)
runs in the CUDA library through the C-ABI; this module only extracts plain arrays from the
embedder (firecode_b200.problem) and wraps the results.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib, conventions, problem
from .errors import ZeroCandidatesError

TIE_CLASH, TIE_TFD, TIE_RMSD, TIE_MAXDEV = 1, 2, 3, 4


def pretty_num(n) -> str:
    if n < 1e3:
        return str(n)
    if n < 1e6:
        return str(round(n / 1e3, 2)) + " k"
    return str(round(n / 1e6, 2)) + " M"


@dataclass
class ScreenReport:
    """What the GPU screen did, kept on ``embedder.b200_report`` after an embed call."""

    n_poses: int = 0
    n_clash_pass: int = 0
    n_fp64_rechecks: int = 0
    n_kept: int = 0
    kept_indices: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    status: np.ndarray | None = None
    ties: np.ndarray | None = None  # structured array (a, b, value, kind, decision)
    n_ties_total: int = 0
    conventions: dict = field(default_factory=conventions.as_dict)
    n_mols: int = 2
    group_choice: np.ndarray | None = None  # trimolecular: grid-search candidate per group
    group_gap: np.ndarray | None = None     # ... and its cost gap to the runner-up (degrees)

    def forced_decisions(self):
        """Near-threshold decisions as the key -> bool mapping oracle.port.Ties understands."""
        out = {}
        if self.ties is None:
            return out
        for t in self.ties:
            if t["kind"] == TIE_CLASH and self.n_mols == 3:
                out[("clash3", int(t["a"]), -1 - int(t["b"]))] = bool(t["decision"])  # b = -1 - block
            elif t["kind"] == TIE_CLASH:
                out[("clash", int(t["a"]))] = bool(t["decision"])
            elif t["kind"] == TIE_TFD:
                out[("tfd", int(t["a"]), int(t["b"]))] = bool(t["decision"])
            elif t["kind"] == TIE_RMSD:
                out[("rmsd", int(t["a"]), int(t["b"]))] = bool(t["decision"])
            elif t["kind"] == TIE_MAXDEV:
                out[("maxdev", int(t["a"]), int(t["b"]))] = bool(t["decision"])
        return out


TIE_DTYPE = np.dtype([("a", np.int64), ("b", np.int64), ("value", np.float64), ("kind", np.int32),
                      ("decision", np.int32)])


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class _Result:
    """RAII wrapper of the opaque fc_result handle."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle
        counts = np.zeros(10, dtype=np.int64)
        _lib.check(lib.fc_result_counts(handle, counts.ctypes.data_as(_lib.c_i64p)), "fc_result_counts")
        (self.n_poses, self.n_clash_pass, self.n_rechecked, self.n_kept, self.n_ties, self.n_atoms,
         self.n_surv, self.n_quads, self.n_pairs, self.n_groups) = (int(x) for x in counts)

    def _get(self, fn, shape, dtype):
        out = np.empty(shape, dtype=dtype)  # filled completely by the library
        if out.size:
            _lib.check(getattr(self.lib, fn)(self.handle, _ptr(out)), fn)
        return out

    def status(self):
        return self._get("fc_result_status", (self.n_poses,), np.uint8)

    def survivors(self):
        return self._get("fc_result_survivors", (self.n_surv,), np.int64)

    def fingerprints(self):
        return self._get("fc_result_fingerprints", (self.n_surv, self.n_quads), np.float64)

    def kept_indices(self):
        return self._get("fc_result_kept_indices", (self.n_kept,), np.int64)

    def kept_coords(self):
        return self._get("fc_result_kept_coords", (self.n_kept, self.n_atoms, 3), np.float64)

    def constrained(self):
        return self._get("fc_result_constrained", (self.n_kept, self.n_pairs, 2), np.int32)

    def groups(self):
        choice = np.zeros(self.n_groups, dtype=np.int32)
        gap = np.zeros(self.n_groups, dtype=np.float64)
        if self.n_groups:
            _lib.check(self.lib.fc_result_groups(self.handle, _ptr(choice), _ptr(gap)), "fc_result_groups")
        return choice, gap

    def ties(self, cap=1 << 20):
        n = min(self.n_ties, cap)
        out = np.zeros(max(n, 1), dtype=TIE_DTYPE)
        total = self.lib.fc_result_ties(self.handle, _ptr(out), n)
        return out[: min(n, max(total, 0))]

    def close(self):
        if self.handle:
            self.lib.fc_result_free(self.handle)
            self.handle = None

    def __del__(self):
        self.close()


def _string_problem_c(prob: problem.StringProblem, max_clashes=0):
    """ctypes struct + the arrays it points into (kept alive by the caller)."""
    keep = {
        "coords1": np.ascontiguousarray(prob.coords[0], dtype=np.float64),
        "coords2": np.ascontiguousarray(prob.coords[1], dtype=np.float64),
        "centers1": np.ascontiguousarray(prob.centers[0], dtype=np.float64),
        "vecs1": np.ascontiguousarray(prob.vecs[0], dtype=np.float64),
        "centers2": np.ascontiguousarray(prob.centers[1], dtype=np.float64),
        "vecs2": np.ascontiguousarray(prob.vecs[1], dtype=np.float64),
        "angles": np.ascontiguousarray(prob.angles, dtype=np.float64),
        "quadruplets": np.ascontiguousarray(prob.quadruplets, dtype=np.int64).reshape(-1, 4),
    }
    c = _lib.StringProblemC()
    c.coords1, c.n_conf1, c.n_atoms1 = _ptr(keep["coords1"]), keep["coords1"].shape[0], keep["coords1"].shape[1]
    c.coords2, c.n_conf2, c.n_atoms2 = _ptr(keep["coords2"]), keep["coords2"].shape[0], keep["coords2"].shape[1]
    c.centers1, c.vecs1, c.k1 = _ptr(keep["centers1"]), _ptr(keep["vecs1"]), keep["centers1"].shape[1]
    c.centers2, c.vecs2, c.k2 = _ptr(keep["centers2"]), _ptr(keep["vecs2"]), keep["centers2"].shape[1]
    c.angles, c.n_angles = _ptr(keep["angles"]), len(keep["angles"])
    c.quadruplets, c.n_quads = _ptr(keep["quadruplets"]), len(keep["quadruplets"])
    c.thresh, c.max_clashes = float(prob.thresh), int(max_clashes)
    c.rot_handedness = int(conventions.ROT_HANDEDNESS)
    c.tfd_thresh = 10.0
    return c, keep


def string_screen(prob: problem.StringProblem):
    """Run the string screen on the current CUDA device. Returns (poses, ScreenReport)."""
    lib = _lib.load(require_device=True)
    c, keep = _string_problem_c(prob)
    handle = C.c_void_p()
    _lib.check(lib.fc_string_screen(C.byref(c), C.byref(handle)), "fc_string_screen")
    res = _Result(lib, handle)
    try:
        report = ScreenReport(n_poses=res.n_poses, n_clash_pass=res.n_clash_pass,
                              n_fp64_rechecks=res.n_rechecked, n_kept=res.n_kept,
                              kept_indices=res.kept_indices(), status=res.status(), ties=res.ties(),
                              n_ties_total=res.n_ties)
        poses = res.kept_coords()
    finally:
        res.close()
    del keep
    return poses, report


def string_embed(embedder):
    """Drop-in for firecode.embeds.string_embed (embeds.py:51-158)."""
    assert len(embedder.objects) == 2
    embedder.log(f"\n--> Performing string embed ({pretty_num(embedder.candidates)} candidates)")
    prob = problem.string_problem(embedder)
    poses, report = string_screen(prob)
    embedder.b200_report = report
    if len(poses) == 0:
        s = (
            "\n--> Cyclical embed did not find any suitable disposition of molecules.\n"
            + "    This is probably because the two molecules cannot find a correct interlocking pose.\n"
            + "    Try expanding the conformational space with the firecode_search> operator or see the SHRINK keyword."
        )
        embedder.log(s, p=False)
        raise ZeroCandidatesError(s)
    embedder.constrained_indices = np.repeat(prob.constrained[None], len(poses), axis=0)
    return poses


def get_embed(mols, conf_ids):
    """Coordinates of every molecule placed by its ``rotation`` / ``position`` (embeds.py:808-817).
    Host-side convenience for callers that assemble one structure; the screens never call it."""
    return np.concatenate([(np.asarray(m.rotation) @ np.asarray(m.coords[c]).T).T + np.asarray(m.position)
                           for m, c in zip(mols, conf_ids)])


# ------------------------------------------------------------------------------------------------
# cyclical embed (bimolecular path)
# ------------------------------------------------------------------------------------------------
def cyclical_groups(prob: problem.CyclicalProblem):
    """Group table of the bimolecular cyclical embed in the reference's loop order (embeds.py:596-641): conformer
    pairs (first index fastest) x pivot pairs (first index fastest) x 2 orientations, minus pivot pairs whose norms
    differ by more than max_norm_delta and arrangements that miss a user pairing (an ndarray ``internal_constraints``
    never matches: SURVEY.md quirk N10, embeds.py:820-826).  Enumerated by the library (C-ABI fc_cyclical_groups, host
    code; the numpy version of round 1 took 24 of the 46 ms of a 2 x 50-conformer embed)."""
    assert prob.n_mols == 2
    lib = _lib.load()
    tabs = []
    for m in range(2):
        o, v = _csr(prob.pivot_vec[m], 3, np.float64)
        _, mp = _csr(prob.pivot_mean[m], 3, np.float64)
        _, ids = _csr(prob.pivot_ids[m], 2, np.int64)
        tabs.append((o, v, mp, ids))
    n_conf = np.array([len(c) for c in prob.coords], dtype=np.int32)
    pairings = np.ascontiguousarray(np.array(prob.pairings, dtype=np.int64).reshape(-1, 2))
    internal = np.ascontiguousarray(np.array([] if prob.internal_constraints_is_array else prob.internal_constraints,
                                             dtype=np.int64).reshape(-1, 2))
    k0 = tabs[0][0][1:] - tabs[0][0][:-1]
    k1 = tabs[1][0][1:] - tabs[1][0][:-1]
    cap = 2 * int(k0.sum()) * int(k1.sum())      # every (pivot pair, orientation) of every conformer pair
    out = {"conf": np.empty((cap, 2), dtype=np.int32), "pivot": np.empty((cap, 2, 3)), "mean": np.empty((cap, 2, 3)),
           "vecs": np.empty((cap, 2, 2, 3)), "dirs": np.empty((cap, 2, 3)), "ids": np.empty((cap, 2, 2), dtype=np.int32)}
    n_groups = C.c_int64(0)
    rc = lib.fc_cyclical_groups(_ptr(n_conf), *(_ptr(t) for t in tabs[0]), *(_ptr(t) for t in tabs[1]),
                                float(prob.max_norm_delta), _ptr(pairings), len(pairings), _ptr(internal), len(internal),
                                cap, C.byref(n_groups), _ptr(out["conf"]), _ptr(out["pivot"]), _ptr(out["mean"]),
                                _ptr(out["vecs"]), _ptr(out["dirs"]), _ptr(out["ids"]))
    _lib.check(rc, "fc_cyclical_groups")
    g = int(n_groups.value)
    assert g <= cap
    return {k: v[:g] for k, v in out.items()}


def cyclical_screen(prob: problem.CyclicalProblem, rmsd_thresh=1.0, group_range=None, groups=None):
    """Run the bimolecular cyclical screen on the current CUDA device.  ``group_range`` = (lo, hi)
    restricts the call to a slice of the group table (multi-GPU sharding).
    Returns (poses, constrained_indices, ScreenReport)."""
    lib = _lib.load(require_device=True)
    groups = groups if groups is not None else cyclical_groups(prob)
    if group_range is not None:
        groups = {k: np.ascontiguousarray(v[group_range[0]:group_range[1]]) for k, v in groups.items()}
    keep = {"coords": [np.ascontiguousarray(c, dtype=np.float64) for c in prob.coords],
            "reactive": [np.ascontiguousarray(r, dtype=np.int64) for r in prob.reactive],
            "angles": np.ascontiguousarray(prob.angles, dtype=np.float64).reshape(-1, 2), **groups}
    c = _lib.CyclicalProblemC()
    c.n_mols = 2
    for m in range(2):
        c.coords[m] = keep["coords"][m].ctypes.data
        c.n_conf[m], c.n_atoms[m] = keep["coords"][m].shape[:2]
        c.reactive[m] = keep["reactive"][m].ctypes.data
        c.n_reactive[m] = len(keep["reactive"][m])
    c.n_groups = len(groups["conf"])
    c.group_conf, c.group_pivot, c.group_mean = _ptr(groups["conf"]), _ptr(groups["pivot"]), _ptr(groups["mean"])
    c.group_vecs, c.group_dirs, c.group_ids = _ptr(groups["vecs"]), _ptr(groups["dirs"]), _ptr(groups["ids"])
    c.n_pairs = 2
    c.angles, c.n_angles = _ptr(keep["angles"]), len(keep["angles"])
    c.thresh, c.max_clashes = float(prob.thresh), 0
    c.rot_handedness = int(conventions.ROT_HANDEDNESS)
    c.rmsd_thresh = float(rmsd_thresh)
    handle = C.c_void_p()
    _lib.check(lib.fc_cyclical_screen(C.byref(c), C.byref(handle)), "fc_cyclical_screen")
    res = _Result(lib, handle)
    try:
        report = ScreenReport(n_poses=res.n_poses, n_clash_pass=res.n_clash_pass,
                              n_fp64_rechecks=res.n_rechecked, n_kept=res.n_kept,
                              kept_indices=res.kept_indices(), status=res.status(), ties=res.ties(),
                              n_ties_total=res.n_ties)
        poses = res.kept_coords()
        constrained = res.constrained().astype(np.int64)
    finally:
        res.close()
    del keep
    return poses, constrained, report


def _csr(per_conf, width, dtype):
    """list over conformers of (P, width) arrays -> (offsets (C+1,), rows (sum P, width))."""
    counts = np.array([len(a) for a in per_conf], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rows = (np.concatenate([np.asarray(a, dtype=dtype).reshape(-1, width) for a in per_conf])
            if len(per_conf) else np.zeros((0, width), dtype=dtype))
    return np.ascontiguousarray(offsets), np.ascontiguousarray(rows.reshape(-1, width).astype(dtype))


def cyclical3_screen(prob: problem.CyclicalProblem, rmsd_thresh=1.0, conf_tuple_range=None, want_status=True,
                     want_coords=True):
    """Run the trimolecular cyclical screen (embeds.py:409-585) on the current CUDA device.
    ``conf_tuple_range`` = (lo, hi) restricts the call to a slice of the conformer-triple enumeration
    (multi-GPU sharding: groups never straddle a slice).  Returns (poses, constrained, ScreenReport)."""
    lib = _lib.load(require_device=True)
    assert prob.n_mols == 3
    keep = {"coords": [np.ascontiguousarray(c, dtype=np.float64) for c in prob.coords],
            "reactive": [np.ascontiguousarray(r, dtype=np.int64) for r in prob.reactive],
            "ratoms0": [np.ascontiguousarray(r, dtype=np.int64).reshape(-1, 2) for r in prob.ratoms0],
            "angles": np.ascontiguousarray(prob.angles, dtype=np.float64).reshape(-1, 3),
            "pairings": np.ascontiguousarray(np.array(prob.pairings, dtype=np.int64).reshape(-1, 2)),
            "internal": np.ascontiguousarray(np.array(
                [] if prob.internal_constraints_is_array else prob.internal_constraints, dtype=np.int64).reshape(-1, 2)),
            "csr": []}
    c = _lib.Cyclical3ProblemC()
    for m in range(3):
        c.coords[m] = keep["coords"][m].ctypes.data
        c.n_conf[m], c.n_atoms[m] = keep["coords"][m].shape[:2]
        c.reactive[m] = keep["reactive"][m].ctypes.data
        c.n_reactive[m] = len(keep["reactive"][m])
        off, vec = _csr(prob.pivot_vec[m], 3, np.float64)
        _, mean = _csr(prob.pivot_mean[m], 3, np.float64)
        _, ids = _csr(prob.pivot_ids[m], 2, np.int64)
        keep["csr"].append((off, vec, mean, ids))
        c.pivot_offsets[m], c.pivot_vec[m] = off.ctypes.data, vec.ctypes.data
        c.pivot_mean[m], c.pivot_ids[m] = mean.ctypes.data, ids.ctypes.data
        c.ratoms0[m], c.n_ratoms0[m] = keep["ratoms0"][m].ctypes.data, len(keep["ratoms0"][m])
    c.angles, c.n_angles = _ptr(keep["angles"]), len(keep["angles"])
    c.pairings, c.n_pairings = _ptr(keep["pairings"]), len(keep["pairings"])
    c.internal, c.n_internal = _ptr(keep["internal"]), len(keep["internal"])
    c.thresh, c.rot_handedness, c.rmsd_thresh = float(prob.thresh), int(conventions.ROT_HANDEDNESS), float(rmsd_thresh)
    c.conf_tuple_lo, c.conf_tuple_hi = (0, 0) if conf_tuple_range is None else (int(conf_tuple_range[0]), int(conf_tuple_range[1]))
    c.flags = (0 if want_status else 1) | (0 if want_coords else 2)
    handle = C.c_void_p()
    _lib.check(lib.fc_cyclical3_screen(C.byref(c), C.byref(handle)), "fc_cyclical3_screen")
    res = _Result(lib, handle)
    try:
        choice, gap = res.groups()
        report = ScreenReport(n_poses=res.n_poses, n_clash_pass=res.n_clash_pass,
                              n_fp64_rechecks=res.n_rechecked, n_kept=res.n_kept,
                              kept_indices=res.kept_indices(), status=res.status() if want_status else None,
                              ties=res.ties(), n_ties_total=res.n_ties, n_mols=3, group_choice=choice, group_gap=gap)
        n_tot = sum(int(x.shape[1]) for x in keep["coords"])
        poses = res.kept_coords() if want_coords else np.zeros((0, n_tot, 3))
        constrained = res.constrained().astype(np.int64)
    finally:
        res.close()
    del keep
    return poses, constrained, report


def cyclical_embed(embedder, max_norm_delta: float = 5.0):
    """Drop-in for firecode.embeds.cyclical_embed (embeds.py:180-585): two molecules ("cyclical" and
    "chelotropic" embeds, embeds.py:184-185 -> 588-750) or three (embeds.py:409-585)."""
    if len(embedder.objects) not in (2, 3):
        raise ValueError("cyclical_embed needs two or three molecules")
    embedder.log(f"\n--> Performing {embedder.embed} embed ({pretty_num(embedder.candidates)} candidates)")
    prob = problem.cyclical_problem(embedder, max_norm_delta=max_norm_delta)
    if len(embedder.objects) == 3:
        # the per-pose status bytes are not part of the reference's interface (216 MB at BASELINE config C2): they
        # come back only when the caller asks for them (tests do, through embedder.b200_want_status)
        poses, constrained, report = cyclical3_screen(prob, want_status=bool(getattr(embedder, "b200_want_status", False)))
    else:
        poses, constrained, report = cyclical_screen(prob)
    embedder.b200_report = report
    embedder.constrained_indices = constrained
    if len(poses) == 0:
        s = (
            "\n--> Cyclical embed did not find any suitable disposition of molecules.\n"
            + "    This is probably because one molecule has two reactive centers at a great distance,\n"
            + "    preventing the other two molecules from forming a closed, cyclical structure."
        )
        embedder.log(s, p=False)
        raise ZeroCandidatesError(s)
    return poses
