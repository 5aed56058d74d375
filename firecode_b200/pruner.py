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


last_report: PruneReport | None = None


def _run(structures, mode, sel, masses, max_rmsd, max_dev, moi_dev, energies, max_dE, keep, pass_mode,
         tie_cap=1 << 16, shard=None):
    """``shard`` = (rank, world, allgather) runs the multi-GPU form: allgather(bytes ndarray) must
    return the rank-order concatenation of every rank's buffer (firecode_b200.dist supplies it)."""
    global last_report
    lib = _lib.load(require_device=True)
    keep = conventions.PRUNE_KEEP if keep is None else keep
    pass_mode = conventions.PRUNE_PASS_MODE if pass_mode is None else pass_mode
    assert keep in ("first", "last") and pass_mode in ("greedy", "snapshot")
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
            float(max_dE), 1 if keep == "first" else 0, 1 if pass_mode == "snapshot" else 0,
            int(conventions.PRUNE_MIN_PER_CHUNK), _ptr(mask), _ptr(stats), _ptr(ties), tie_cap, C.byref(n_ties))
    if shard is None:
        rc = lib.fc_prune(*args)
    else:
        rank, world, allgather = shard
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
    last_report = PruneReport(wall_ms=float(tm[0]), screen_ms=float(tm[1]), screen_launches=int(tm[2]),
                              screen_pair_slots=float(tm[3]), screen_candidates=int(tm[4]), n_sel=int(tm[5]),
                              passes=int(stats[0]), pairs_tiled=int(stats[1]), pairs_solved=int(stats[2]),
                              pairs_skipped=int(stats[3]),
                              ties=ties[: min(int(n_ties.value), tie_cap)], n_ties_total=int(n_ties.value),
                              keep=keep, pass_mode=pass_mode)
    mask = mask.astype(bool)
    return _take(lib, x, mask), mask


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
                  debugfunction=None, logfunction=None, keep=None, pass_mode=None, shard=None):
    """Heavy-atom, centred Kabsch RMSD pruning: a structure is dropped when a kept one has
    rmsd < max_rmsd and max atomic deviation < max_dev (default 2 * max_rmsd)."""
    atoms = np.asarray(atoms)
    max_dev = conventions.PRUNE_MAXDEV_FACTOR * max_rmsd if max_dev is None else max_dev
    if conventions.PRUNE_RMSD_HEAVY_ONLY:
        sel = np.flatnonzero(np.array([str(a) != "H" for a in atoms]))
    else:
        sel = np.arange(len(atoms))
    out, mask = _run(structures, 0, sel, None, max_rmsd, max_dev, 0.0, energies, max_dE, keep, pass_mode, shard=shard)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_rmsd (firecode_b200) - kept {int(mask.sum())}/{len(mask)}")
    return out, mask


def prune_by_moment_of_inertia(structures, atoms, max_deviation=None, energies=None, max_dE=0.0,
                               debugfunction=None, logfunction=None, keep=None, pass_mode=None, shard=None):
    """Drop structures whose three principal moments of inertia are all within ``max_deviation``
    (relative, default 1 %) of a kept structure (CHANGELOG.md:256)."""
    max_deviation = conventions.MOI_MAX_DEVIATION if max_deviation is None else max_deviation
    masses = np.array([MASSES[str(a)] for a in np.asarray(atoms)])
    out, mask = _run(structures, 1, None, masses, 0.0, 0.0, max_deviation, energies, max_dE, keep, pass_mode, shard=shard)
    if debugfunction is not None:
        debugfunction(f"DEBUG: prune_by_moment_of_inertia (firecode_b200) - kept {int(mask.sum())}/{len(mask)}")
    return out, mask


def prune_by_rmsd_rot_corr(structures, atoms, graph, max_rmsd=0.25, max_dev=None, energies=None,
                           max_dE=0.0, logfunction=None, debugfunction=None, keep=None, pass_mode=None):
    """Symmetry-corrected RMSD pruning (embedder.py:1489).  The torsion-symmetry enumeration of
    prism_pruner is a "next" row of the scope table: until it is built this
    entry point applies the plain heavy-atom RMSD criterion, which is the first of the two tests
    the corrected variant performs, and says so through ``logfunction``."""
    if logfunction is not None:
        logfunction("firecode_b200: rotationally-corrected RMSD pruning not built yet - plain RMSD criterion applied")
    return prune_by_rmsd(structures, atoms, max_rmsd=max_rmsd, max_dev=max_dev, energies=energies,
                         max_dE=max_dE, debugfunction=debugfunction, keep=keep, pass_mode=pass_mode)


def prune(structures, atoms, max_rmsd=0.25, logfunction=None, debugfunction=None, **kw):
    """MOI pruning followed by RMSD pruning (interfaces/goat.py:399)."""
    structures = np.asarray(structures, dtype=np.float64)
    s1, m1 = prune_by_moment_of_inertia(structures, atoms, debugfunction=debugfunction)
    s2, m2 = prune_by_rmsd(s1, atoms, max_rmsd=max_rmsd, debugfunction=debugfunction)
    mask = m1.copy()
    mask[np.flatnonzero(m1)] = m2
    return s2, mask
