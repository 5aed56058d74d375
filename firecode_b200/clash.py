"""Batched compenetration (clash) screen -- host side of the C-ABI `fc_clash_*` entry points.

Reference: firecode/utils.py:507-575 `compenetration_check`, called per pose at
embeds.py:139-141, 559-563, 718-722.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

STATUS_PASS, STATUS_RECHECKED, STATUS_NEAR = 1, 2, 4
NEAR_EPS = 1e-6


@dataclass
class ClashResult:
    status: np.ndarray  # (n_poses,) uint8 FC_STATUS_* bits
    min_dist: np.ndarray | None  # (n_poses,) f32 estimate (exact f64->f32 for rechecked poses)
    n_pass: int = 0
    n_rechecked: int = 0
    near_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    near_dist: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.float64))
    _mask: np.ndarray | None = None

    @property
    def mask(self) -> np.ndarray:
        """(n_poses,) bool, True = passes (no compenetration); derived from ``status`` on first use."""
        if self._mask is None:
            self._mask = (self.status & STATUS_PASS).astype(bool)
        return self._mask


def tile_poses(n_b: int) -> int:
    return int(_lib.load().fc_clash_tile_poses(int(n_b)))


def build_tiles(conf_a, conf_b, n_b: int) -> np.ndarray:
    """(n_tiles, 4) int32 {conf_a, conf_b, first_pose, count}: contiguous runs of equal conformer
    pairs cut into tiles of at most fc_clash_tile_poses(n_b) poses."""
    conf_a = np.asarray(conf_a, dtype=np.int64).ravel()
    conf_b = np.asarray(conf_b, dtype=np.int64).ravel()
    n = len(conf_a)
    if n == 0:
        return np.zeros((0, 4), dtype=np.int32)
    p = tile_poses(n_b)
    change = np.flatnonzero((np.diff(conf_a) != 0) | (np.diff(conf_b) != 0)) + 1
    starts = np.concatenate([[0], change])
    lengths = np.diff(np.concatenate([starts, [n]]))
    n_tiles = (lengths + p - 1) // p
    run_of_tile = np.repeat(np.arange(len(starts)), n_tiles)
    first_tile_of_run = np.concatenate([[0], np.cumsum(n_tiles)[:-1]])
    j = np.arange(int(n_tiles.sum())) - first_tile_of_run[run_of_tile]
    first = starts[run_of_tile] + j * p
    count = np.minimum(p, starts[run_of_tile] + lengths[run_of_tile] - first)
    tiles = np.stack([conf_a[first], conf_b[first], first, count], axis=1)
    assert first.max(initial=0) < 2**31
    return np.ascontiguousarray(tiles.astype(np.int32))


def _as_ensemble(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if x.ndim == 2:
        x = x[None]
    assert x.ndim == 3 and x.shape[2] == 3, "ensemble must be (n_conf, n_atoms, 3)"
    return x


def compenetration_check_batch(frag_a, frag_b, xf, thresh=1.0, max_clashes=0, conf_a=None,
                               conf_b=None, strict=True, want_min_dist=False, near_cap=4096):
    """Screen n poses of fragment B against fragment A on the GPU (host-buffer C-ABI path).

    frag_a, frag_b: (n_atoms, 3) or (n_conf, n_atoms, 3) float64 ensembles.
    xf: (n, 12) float64, R row-major then t: pose places B at ``R @ b + t``.
    conf_a, conf_b: optional (n,) conformer index per pose (default: conformer 0).
    Returns ClashResult; ``mask[p]`` equals the reference's
    ``compenetration_check(concat(a, R @ b + t), ids=[n_a, n_b], thresh, max_clashes)``.
    """
    lib = _lib.load(require_device=True)
    a = _as_ensemble(frag_a)
    b = _as_ensemble(frag_b)
    xf = np.ascontiguousarray(np.asarray(xf, dtype=np.float64).reshape(-1, 12))
    n = xf.shape[0]
    tiles = None
    if conf_a is not None or conf_b is not None:
        ca = np.zeros(n, dtype=np.int64) if conf_a is None else np.asarray(conf_a)
        cb = np.zeros(n, dtype=np.int64) if conf_b is None else np.asarray(conf_b)
        assert len(ca) == n and len(cb) == n
        assert n == 0 or (ca.min() >= 0 and ca.max() < a.shape[0] and cb.min() >= 0 and cb.max() < b.shape[0])
        tiles = build_tiles(ca, cb, b.shape[1])
    status = np.empty(n, dtype=np.uint8)
    min_dist = np.empty(n, dtype=np.float32) if want_min_dist else None
    counts = np.zeros(3, dtype=np.int64)
    near_idx = np.zeros(max(near_cap, 1), dtype=np.int64)
    near_dist = np.zeros(max(near_cap, 1), dtype=np.float64)

    def ptr(arr):
        return None if arr is None else arr.ctypes.data_as(C.c_void_p)

    rc = lib.fc_clash_batch(ptr(a), a.shape[0], a.shape[1], ptr(b), b.shape[0], b.shape[1], ptr(xf), n,
                            ptr(tiles), 0 if tiles is None else len(tiles), float(thresh),
                            int(max_clashes), 1 if strict else 0, ptr(status), ptr(min_dist),
                            ptr(counts), ptr(near_idx), ptr(near_dist), int(near_cap))
    _lib.check(rc, "fc_clash_batch")
    n_near = int(min(counts[2], near_cap))
    order = np.argsort(near_idx[:n_near], kind="stable")
    return ClashResult(status=status, min_dist=min_dist, n_pass=int(counts[0]),
                       n_rechecked=int(counts[1]), near_idx=near_idx[:n_near][order],
                       near_dist=near_dist[:n_near][order])


def screen_device(a_dev, b_dev, xf_dev, thresh, max_clashes=0, strict=True, tiles_dev=None,
                  status_out=None, min_dist_out=None, near=None, pose_index_base=0):
    """Device-resident clash screen on torch CUDA tensors (no copies, no synchronisation).

    a_dev (n_conf_a, n_a, 3) f64, b_dev (n_conf_b, n_b, 3) f64, xf_dev (n, 12) f64, optional
    tiles_dev (n_tiles, 4) int32.  ``near`` = (count int32[1], idx int64[cap], dist f64[cap])
    device tensors for the near-threshold list.  Runs on the current torch stream and returns the
    status tensor (uint8, FC_STATUS_* bits)."""
    import torch

    lib = _lib.load(require_device=True)
    assert a_dev.is_cuda and b_dev.is_cuda and xf_dev.is_cuda
    assert a_dev.dtype == torch.float64 and b_dev.dtype == torch.float64 and xf_dev.dtype == torch.float64
    assert a_dev.is_contiguous() and b_dev.is_contiguous() and xf_dev.is_contiguous()
    if a_dev.dim() == 2:
        a_dev = a_dev[None]
    if b_dev.dim() == 2:
        b_dev = b_dev[None]
    n = xf_dev.shape[0]
    if status_out is None:
        status_out = torch.empty(n, dtype=torch.uint8, device=xf_dev.device)
    stream = torch.cuda.current_stream(xf_dev.device).cuda_stream
    ncount = nidx = ndist = None
    cap = 0
    if near is not None:
        ncount, nidx, ndist = (t.data_ptr() for t in near)
        cap = near[1].numel()
    rc = lib.fc_clash_screen_dev(
        a_dev.data_ptr(), a_dev.shape[0], a_dev.shape[1], b_dev.data_ptr(), b_dev.shape[0],
        b_dev.shape[1], xf_dev.data_ptr(), n,
        None if tiles_dev is None else tiles_dev.data_ptr(),
        0 if tiles_dev is None else tiles_dev.shape[0], float(thresh), int(max_clashes),
        1 if strict else 0, status_out.data_ptr(),
        None if min_dist_out is None else min_dist_out.data_ptr(), ncount, nidx, ndist, cap,
        int(pose_index_base), stream)
    _lib.check(rc, "fc_clash_screen_dev")
    return status_out


def pack_mask_device(status_dev, bits_out=None):
    """status (n,) uint8 CUDA tensor -> survivor bitmask (ceil(n/32),) int32 CUDA tensor."""
    import torch

    lib = _lib.load(require_device=True)
    n = status_dev.numel()
    if bits_out is None:
        bits_out = torch.empty((n + 31) // 32, dtype=torch.int32, device=status_dev.device)
    stream = torch.cuda.current_stream(status_dev.device).cuda_stream
    _lib.check(lib.fc_pack_mask_dev(status_dev.data_ptr(), n, bits_out.data_ptr(), stream),
               "fc_pack_mask_dev")
    return bits_out
