"""Similarity pruning with the call surface FIRECODE uses from ``prism_pruner.pruner``
(embedder.py:1452,1472,1489; ensemble.py:211,230,253; operators.py:613-632; goat.py:399):
``fn(structures, atoms, ...) -> (structures[mask], mask)``, mask = numpy bool array in input order.

The pair similarities and the order-dependent keep rule run on the GPU (C-ABI ``fc_prune``).
prism_pruner itself is absent from the reference tree, so the two conventions that cannot be
verified (which member of a similar pair survives; whether a pass reads a snapshot of the mask) are
the switches ``firecode_b200.conventions.PRUNE_KEEP`` / ``PRUNE_PASS_MODE``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib, conventions
from .embeds import TIE_DTYPE, _ptr
from .pt import MASSES


@dataclass
class PruneReport:
    passes: int = 0
    pairs_tiled: int = 0     # active pairs whose 3x3 covariance was accumulated
    pairs_solved: int = 0    # pairs that needed the FP64 eigen-solve
    pairs_skipped: int = 0   # pairs known dissimilar from an earlier pass (same chunk survivors)
    ties: np.ndarray | None = None
    n_ties_total: int = 0
    keep: str = "first"
    pass_mode: str = "greedy"
    wall_ms: float = 0.0            # inside the library call
    screen_ms: float = 0.0          # CUDA-event time of the tensor-core screen launches (0 when another screen ran)
    screen_launches: int = 0
    screen_pair_slots: float = 0.0  # 2048 per 128 x 16 tile the screen evaluated
    screen_candidates: int = 0      # pairs handed to the FP64 exact kernel
    n_sel: int = 0                  # atoms per structure that enter the RMSD
    screen_tiles_planned: int = 0   # 128 x 16 tiles of the passes' work items ...
    screen_tiles_multiplied: int = 0  # ... and those the screen did not skip by the shape bound (sigma ranges too far apart)


last_report: PruneReport | None = None


def _chunk_bounds(mask, k, chunk_over):
    """Index ranges [(first, last)] of the k chunks of a pass (conventions.PRUNE_CHUNK_OVER)."""
    n = len(mask)
    if chunk_over == "full":
        size = n // k
        return [(c * size, n if c == k - 1 else size * (c + 1)) for c in range(k)]
    act = np.flatnonzero(mask)
    size = max(1, len(act) // k)
    starts = [0] + [int(act[c * size]) if c * size < len(act) else n for c in range(1, k)]
    return [(starts[c], starts[c + 1] if c + 1 < k else n) for c in range(k)]


def _run(structures, mode, sel, masses, max_rmsd, max_dev, moi_dev, energies, max_dE, keep, pass_mode,
         tie_cap=1 << 16, shard=None, chunk_over=None, want_structures=True):
    """``shard`` = (rank, world, allgather) runs the multi-GPU form with host-staged lists: allgather(bytes ndarray)
    must return the rank-order concatenation of every rank's buffer; ``shard`` = (rank, world, None, allgather_dev) the
    device-resident form (C-ABI fc_prune_sharded_dev).  firecode_b200.dist supplies both."""
    global last_report
    lib = _lib.load(require_device=True)
    keep = conventions.PRUNE_KEEP if keep is None else keep
    pass_mode = conventions.PRUNE_PASS_MODE if pass_mode is None else pass_mode
    chunk_over = conventions.PRUNE_CHUNK_OVER if chunk_over is None else chunk_over
    assert keep in ("first", "last") and pass_mode in ("greedy", "snapshot") and chunk_over in ("full", "active")
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    assert x.ndim == 3 and x.shape[2] == 3
    n, n_atoms = x.shape[:2]
    mask = np.ones(n, dtype=np.uint8)
    stats = np.zeros(4, dtype=np.int64)
    ties = np.zeros(max(tie_cap, 1), dtype=TIE_DTYPE)
    n_ties = C.c_int64(0)
    sel_arr = None if sel is None else np.ascontiguousarray(sel, dtype=np.int32)
    mass_arr = None if masses is None else np.ascontiguousarray(masses, dtype=np.float64)
    e_arr = None if energies is None else np.ascontiguousarray(energies, dtype=np.float64)
    if e_arr is not None:
        assert len(e_arr) == n
    args = (_ptr(x), n, n_atoms, mode, _ptr(sel_arr), 0 if sel_arr is None else len(sel_arr),
            _ptr(mass_arr), float(max_rmsd), float(max_dev), float(moi_dev), _ptr(e_arr),
            float(max_dE), 1 if keep == "first" else 0,
            (1 if pass_mode == "snapshot" else 0) | (2 if chunk_over == "active" else 0),
            int(conventions.PRUNE_MIN_PER_CHUNK), _ptr(mask), _ptr(stats), _ptr(ties), tie_cap, C.byref(n_ties))
    if shard is None:
        rc = lib.fc_prune(*args)
    elif len(shard) == 4 and shard[3] is not None:
        # device all-gather: shard[3](send_ptr, recv_ptr, nbytes, stream_ptr) exchanges device buffers (NCCL); every rank
        # uploads 1 / world of the structures, the lists never leave the GPUs until the union is sorted
        rank, world, _, allgather_dev = shard

        def _gather_dev(send, recv, nbytes, stream, ctx):
            try:
                allgather_dev(int(send or 0), int(recv or 0), int(nbytes), int(stream or 0))
                return 0
            except Exception:  # pragma: no cover - reported through the C-ABI error path
                import traceback

                traceback.print_exc()
                return 1

        cb = _lib.ALLGATHER_DEV_FN(_gather_dev)
        rc = lib.fc_prune_sharded_dev(*args, int(rank), int(world), cb, None)
    else:
        rank, world, allgather = shard[:3]
        hold = {}

        def _gather(send, send_bytes, recv, recv_bytes, ctx):
            try:
                buf = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(send_bytes,)).copy() \
                    if send_bytes else np.zeros(0, dtype=np.uint8)
                out = np.ascontiguousarray(allgather(buf), dtype=np.uint8)
                hold["buf"] = out  # owned here until the next call
                recv[0] = out.ctypes.data if out.size else None
                recv_bytes[0] = out.size
                return 0
            except Exception:  # pragma: no cover - reported through the C-ABI error path
                import traceback

                traceback.print_exc()
                return 1

        cb = _lib.ALLGATHER_FN(_gather)
        rc = lib.fc_prune_sharded(*args, int(rank), int(world), cb, None)
    _lib.check(rc, "fc_prune")
    tm = np.zeros(6, dtype=np.float64)
    lib.fc_prune_timing(_ptr(tm))
    tl = np.zeros(2, dtype=np.float64)
    lib.fc_prune_tiles(_ptr(tl))
    last_report = PruneReport(screen_tiles_planned=int(tl[0]), screen_tiles_multiplied=int(tl[1]), wall_ms=float(tm[0]), screen_ms=float(tm[1]), screen_launches=int(tm[2]),
                              screen_pair_slots=float(tm[3]), screen_candidates=int(tm[4]), n_sel=int(tm[5]),
                              passes=int(stats[0]), pairs_tiled=int(stats[1]), pairs_solved=int(stats[2]),
                              pairs_skipped=int(stats[3]),
                              ties=ties[: min(int(n_ties.value), tie_cap)], n_ties_total=int(n_ties.value),
                              keep=keep, pass_mode=pass_mode)
    mask = mask.astype(bool)
    # want_structures=False: the mask alone (every rank of a sharded call would otherwise write its own copy of the kept
    # structures -- 232 MB per rank at BASELINE config C4 -- through the one host memory they share)
    return (_take(lib, x, mask) if want_structures else None), mask


def _take(lib, x, mask):
    """``x[mask]`` for a C-contiguous array, copied by several host threads (fc_take_rows)."""
    n_keep = int(np.count_nonzero(mask))
    out = np.empty((n_keep,) + x.shape[1:], dtype=x.dtype)
    if n_keep:
        row_bytes = x.dtype.itemsize * int(np.prod(x.shape[1:]))
        m8 = mask.view(np.uint8)
        _lib.check(lib.fc_take_rows(_ptr(x), row_bytes, _ptr(m8), len(mask), _ptr(out), n_keep), "fc_take_rows")
    return out


def prune_by_rmsd(structures, atoms, max_rmsd=0.25, max_dev=None, energies=None, max_dE=0.0,
                  debugfunction=None, logfunction=None, keep=None, pass_mode=None, shard=None, chunk_over=None,
                  want_structures=True):
    """Heavy-atom, centred Kabsch RMSD pruning: a structure is dropped when a kept one has
    rmsd < max_rmsd and max atomic deviation < max_dev (default 2 * max_rmsd)."""
    atoms = np.asarray(atoms)
    max_dev = conventions.PRUNE_MAXDEV_FACTOR * max_rmsd if max_dev is None else max_dev
    if conventions.PRUNE_RMSD_HEAVY_ONLY:
        sel = np.flatnonzero(np.array([str(a) != "H" for a in atoms]))
    else:
        sel = np.arange(len(atoms))
    out, mask = _run(structures, 0, sel, None, max_rmsd, max_dev, 0.0, energies, max_dE, keep, pass_mode, shard=shard,
                     chunk_over=chunk_over, want_structures=want_structures)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_rmsd (firecode_b200) - kept {int(mask.sum())}/{len(mask)}")
    return out, mask


def prune_by_moment_of_inertia(structures, atoms, max_deviation=None, energies=None, max_dE=0.0,
                               debugfunction=None, logfunction=None, keep=None, pass_mode=None, shard=None,
                               chunk_over=None, want_structures=True):
    """Drop structures whose three principal moments of inertia are all within ``max_deviation``
    (relative, default 1 %) of a kept structure (CHANGELOG.md:256)."""
    max_deviation = conventions.MOI_MAX_DEVIATION if max_deviation is None else max_deviation
    masses = np.array([MASSES[str(a)] for a in np.asarray(atoms)])
    out, mask = _run(structures, 1, None, masses, 0.0, 0.0, max_deviation, energies, max_dE, keep, pass_mode, shard=shard,
                     chunk_over=chunk_over, want_structures=want_structures)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_moment_of_inertia (firecode_b200) - kept {int(mask.sum())}/{len(mask)}")
    return out, mask


K_SCHEDULE = (500_000, 200_000, 100_000, 50_000, 20_000, 10_000, 5000, 2000, 1000, 500, 200, 100, 50, 20, 10, 5, 2, 1)


@dataclass
class RotCorrReport:
    n_torsions: int = 0
    n_folds: tuple = ()
    passes: int = 0
    pairs_evaluated: int = 0
    ties: list = None            # [(kind, later, earlier, value, decision)] within 1e-6 of a threshold
    choices: dict = None         # {((earlier, later), torsion): angle index} where the two best angles tie (gap <= 1e-9)
    keep: str = "first"
    pass_mode: str = "greedy"


last_rot_corr_report: RotCorrReport | None = None


def host_symmetry_torsions(coords, atoms, graph):
    """The symmetric torsions of a structure, found by the HOST application's own torsion machinery (FIRECODE's
    torsion_module.py:70-432 and prism_pruner.utils.get_double_bonds_indices, which a plugin runs inside of): graph
    chemistry at set-up scale, O(atoms), not part of the per-pair hot path.  Selection as prune_by_rmsd_rot_corr's
    ([UNVERIFIED-RECALL], see oracle/prism_pruner/pruner.py:symmetry_torsions).  Returns (quadruplets, angles, masks)."""
    try:
        from firecode.torsion_module import _get_hydrogen_bonds, _get_rotation_mask, _get_torsions, _is_nondummy
        from prism_pruner.utils import get_double_bonds_indices
    except Exception as exc:
        raise _lib.FirecodeB200Error(
            "prune_by_rmsd_rot_corr needs the symmetric torsions of the molecule: pass torsions=/angles=/masks= or run "
            f"inside FIRECODE (firecode.torsion_module / prism_pruner.utils not importable: {exc})") from exc
    atoms = np.asarray(atoms)
    torsions = _get_torsions(graph, hydrogen_bonds=_get_hydrogen_bonds(atoms, coords, graph),
                             double_bonds=get_double_bonds_indices(coords, atoms), keepdummy=True, mode="symmetry")
    torsions = [t for t in torsions if not (_is_nondummy(t.i2, t.i3, graph) and _is_nondummy(t.i3, t.i2, graph))]
    torsions = [t for t in torsions if "H" not in [str(atoms[i]) for i in t.torsion]]
    angles = [tuple(float(x) for x in t.get_angles()) for t in torsions]
    quads = [tuple(t.torsion) if _is_nondummy(t.i2, t.i3, graph) else tuple(reversed(t.torsion)) for t in torsions]
    masks = [np.asarray(_get_rotation_mask(graph, q), dtype=bool) for q in quads]
    return quads, angles, masks


def rmsd_and_max_rot_corr_pairs(structures, atoms, pairs, torsions, angles, masks, want_choices=False):
    """Symmetry-corrected (rmsd, maxdev) of structure pairs {earlier, later} on the GPU (C-ABI fc_rmsd_rot_corr_pairs)."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    n, n_atoms = x.shape[:2]
    atoms = np.asarray(atoms)
    sel = np.flatnonzero(np.array([str(a) != "H" for a in atoms])) if conventions.PRUNE_RMSD_HEAVY_ONLY else np.arange(n_atoms)
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    tor = np.ascontiguousarray(np.asarray(torsions, dtype=np.int32).reshape(-1, 4))
    n_tors = len(tor)
    msk = np.ascontiguousarray(np.asarray(masks, dtype=np.uint8).reshape(n_tors, n_atoms)) if n_tors else None
    flat = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64) for a in angles])) if n_tors else None
    off = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(a) for a in angles])]).astype(np.int32))
    rmsd = np.empty(len(pairs))
    dev = np.empty(len(pairs))
    choice = np.zeros((len(pairs), max(n_tors, 1)), dtype=np.int32) if want_choices else None
    gap = np.full((len(pairs), max(n_tors, 1)), np.inf) if want_choices else None
    step = 1 << 22
    for lo in range(0, len(pairs), step):
        hi = min(len(pairs), lo + step)
        rc = lib.fc_rmsd_rot_corr_pairs(_ptr(x), n, n_atoms, _ptr(sel), len(sel), _ptr(tor) if n_tors else None, n_tors,
                                        _ptr(msk), _ptr(flat), _ptr(off), _ptr(pairs[lo:hi]), hi - lo,
                                        int(conventions.ROT_HANDEDNESS), int(conventions.TORSION_AXIS_SIGN),
                                        _ptr(rmsd[lo:hi]), _ptr(dev[lo:hi]),
                                        None if choice is None else _ptr(choice[lo:hi]),
                                        None if gap is None else _ptr(gap[lo:hi]))
        _lib.check(rc, "fc_rmsd_rot_corr_pairs")
    if want_choices:
        return rmsd, dev, choice[:, :n_tors], gap[:, :n_tors]
    return rmsd, dev


def prune_by_rmsd_rot_corr(structures, atoms, graph, max_rmsd=0.25, max_dev=None, energies=None,
                           max_dE=0.0, logfunction=None, debugfunction=None, keep=None, pass_mode=None,
                           torsions=None, angles=None, masks=None, chunk_over=None):
    """Symmetry-corrected RMSD pruning (embedder.py:1485-1496, ensemble.py:253, operators.py:626): structures that
    differ only by a rotation of a locally symmetric group (methyl, tert-butyl, CF3, phenyl ...) count as similar.
    The pair distance is the symmetry-corrected RMSD of fc_rmsd_rot_corr_pairs, evaluated on the GPU for every
    active pair of a chunk; the multi-pass chunked driver and its keep rule are those of prune_by_rmsd.  The
    symmetric torsions come from the host application's torsion machinery unless passed in.  PARITY UNPINNED: the
    algorithm is [UNVERIFIED-RECALL] of prism_pruner 0.0.7 (csrc/fc_rotcorr.cu states it)."""
    global last_rot_corr_report
    keep = conventions.PRUNE_KEEP if keep is None else keep
    pass_mode = conventions.PRUNE_PASS_MODE if pass_mode is None else pass_mode
    chunk_over = conventions.PRUNE_CHUNK_OVER if chunk_over is None else chunk_over
    assert keep in ("first", "last") and pass_mode in ("greedy", "snapshot") and chunk_over in ("full", "active")
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    n = len(x)
    max_dev = conventions.PRUNE_MAXDEV_FACTOR * max_rmsd if max_dev is None else max_dev
    if torsions is None:
        torsions, angles, masks = host_symmetry_torsions(x[0], atoms, graph) if n else ([], [], [])
    rep = RotCorrReport(n_torsions=len(torsions), n_folds=tuple(len(a) for a in angles), ties=[], choices={},
                        keep=keep, pass_mode=pass_mode)
    last_rot_corr_report = rep
    mask = np.ones(n, dtype=bool)
    if len(torsions) == 0 or n < 2:
        return x[mask], mask
    if logfunction is not None:
        logfunction(f"Rotationally-corrected RMSD pruning: {len(torsions)} symmetric torsions ({list(rep.n_folds)}-fold)")
    e = None if energies is None else np.asarray(energies, dtype=np.float64)
    for k in K_SCHEDULE:
        active = int(np.count_nonzero(mask))
        if not (k == 1 or conventions.PRUNE_MIN_PER_CHUNK * k < active):
            continue
        rep.passes += 1
        bounds = _chunk_bounds(mask, k, chunk_over)
        # every pair of structures active at the start of the pass, chunk by chunk, in one GPU call
        blocks = []
        for first, last in bounds:
            idx = first + np.flatnonzero(mask[first:last])
            if len(idx) > 1:
                i, j = np.triu_indices(len(idx), 1)
                blocks.append(np.stack([idx[i], idx[j]], axis=1))
        if not blocks:
            continue
        pairs = np.concatenate(blocks)
        if e is not None:
            pairs = pairs[np.abs(e[pairs[:, 0]] - e[pairs[:, 1]]) < max_dE]
        if len(pairs) == 0:
            continue
        rmsd, dev, choice, gap = rmsd_and_max_rot_corr_pairs(x, atoms, pairs, torsions, angles, masks, want_choices=True)
        rep.pairs_evaluated += len(pairs)
        similar = (rmsd < max_rmsd) & (dev < max_dev)
        for kind, val, thr in (("rmsd", rmsd, max_rmsd), ("maxdev", dev, max_dev)):
            for p in np.flatnonzero(np.abs(val - thr) <= 1e-6):
                rep.ties.append((kind, int(pairs[p, 1]), int(pairs[p, 0]), float(val[p]), bool(val[p] < thr)))
        for p, t in zip(*np.nonzero(gap <= 1e-9)):
            rep.choices[((int(pairs[p, 0]), int(pairs[p, 1])), int(t))] = int(choice[p, t])
        sim_pairs = pairs[similar]
        # ordered resolution of the pass (the keep rule of prune_by_rmsd) over the similar pairs only
        chunk_of = np.empty(n, dtype=np.int64)
        for c, (first, last) in enumerate(bounds):
            chunk_of[first:last] = c
        partners = {}
        for a, b in sim_pairs:
            partners.setdefault(int(a), []).append(int(b))
            partners.setdefault(int(b), []).append(int(a))
        if pass_mode == "greedy":
            order = range(n) if keep == "first" else range(n - 1, -1, -1)
            for i in order:
                if not mask[i]:
                    continue
                for j in partners.get(i, ()):
                    if mask[j] and (j > i if keep == "first" else j < i):
                        mask[j] = False
        else:
            snap = mask.copy()
            for i in range(n):
                if snap[i] and any(snap[j] and (j < i if keep == "first" else j > i) for j in partners.get(i, ())):
                    mask[i] = False
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_rmsd_rot_corr (firecode_b200) - kept {int(mask.sum())}/{len(mask)}")
    lib = _lib.load()
    return _take(lib, x, mask), mask


_PRUNE_KW = ("energies", "max_dE", "keep", "pass_mode", "shard", "chunk_over")


def prune(structures, atoms, max_rmsd=0.25, logfunction=None, debugfunction=None, max_dev=None, max_deviation=None,
          **kw):
    """MOI pruning followed by RMSD pruning (interfaces/goat.py:399).  energies / max_dE / keep / pass_mode / shard
    reach both stages; anything else is a TypeError (a silently dropped energy window would un-gate the pruning)."""
    unknown = [k for k in kw if k not in _PRUNE_KW]
    if unknown:
        raise TypeError(f"prune() got unexpected keyword arguments {unknown}")
    structures = np.asarray(structures, dtype=np.float64)
    s1, m1 = prune_by_moment_of_inertia(structures, atoms, max_deviation=max_deviation, debugfunction=debugfunction, **kw)
    kw2 = dict(kw)
    if kw2.get("energies") is not None:
        kw2["energies"] = np.asarray(kw2["energies"])[m1]
    s2, m2 = prune_by_rmsd(s1, atoms, max_rmsd=max_rmsd, max_dev=max_dev, debugfunction=debugfunction, **kw2)
    mask = m1.copy()
    mask[np.flatnonzero(m1)] = m2
    return s2, mask
