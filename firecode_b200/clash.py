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


POSE_XF64, POSE_Q7 = 0, 1


class _PinnedBlock:
    """Owner of one fc_host_alloc block; numpy views keep it alive through their base chain."""

    def __init__(self, lib, nbytes):
        self._lib = lib
        self.node = C.c_int32(-1)
        self.ptr = lib.fc_host_alloc(nbytes, C.byref(self.node))
        if not self.ptr:
            raise _lib.FirecodeB200Error("fc_host_alloc failed: " + lib.fc_last_error().decode(errors="replace"))
        self.__array_interface__ = {"data": (self.ptr, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                self._lib.fc_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory on the NUMA node of the current CUDA device (C-ABI fc_host_alloc);
    the host-buffer entry points reach the PCIe rate only from such memory.  Freed with the last view of it."""
    lib = _lib.load(require_device=True)
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    block = _PinnedBlock(lib, max(1, n * dtype.itemsize))
    return np.asarray(block)[: n * dtype.itemsize].view(dtype).reshape(shape)


def cell_grid_meta(frag_a, thresh):
    """Geometry of the cell grid the screen builds for fragment A (test hook, C-ABI fc_clash_cell_meta):
    dict(origin (3,) = centre of cell (0,0,0), h, g, rc)."""
    lib = _lib.load(require_device=True)
    a = _as_ensemble(frag_a)
    out = (C.c_float * 8)()
    _lib.check(lib.fc_clash_cell_meta(a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], float(thresh), out),
               "fc_clash_cell_meta")
    return {"origin": np.array(out[0:3], dtype=np.float64), "h": float(out[3]), "g": int(out[4]), "rc": float(out[5])}


def pack_poses7(quat, trans):
    """(n, 4) quaternions (x, y, z, w) + (n, 3) translations -> (n, 7) float32 compact poses (FC_POSE_Q7)."""
    quat = np.asarray(quat).reshape(-1, 4)
    trans = np.asarray(trans).reshape(-1, 3)
    out = np.empty((len(quat), 7), dtype=np.float32)
    out[:, :4] = quat
    out[:, 4:] = trans
    return out


def unpack_bits(bits, n):
    """survivor bitmask (ceil(n/32),) uint32 -> (n,) bool"""
    return np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), bitorder="little")[:n].astype(bool)


@dataclass
class ClashBitsResult:
    bits: np.ndarray  # (ceil(n/32),) uint32: bit (i & 31) of word i >> 5 = pose i passes
    n_poses: int
    n_pass: int = 0
    n_rechecked: int = 0
    n_near: int = 0
    near_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    near_dist: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.float64))
    status: np.ndarray | None = None
    _mask: np.ndarray | None = None

    @property
    def mask(self) -> np.ndarray:
        if self._mask is None:
            self._mask = unpack_bits(self.bits, self.n_poses)
        return self._mask


def compenetration_check_batch_pose7(frag_a, frag_b, pose7, thresh=1.0, max_clashes=0, conf_a=None, conf_b=None,
                                     strict=True, want_status=False, near_cap=4096, bits_out=None):
    """Screen n compact poses (n, 7) float32 {quaternion x y z w, translation} on the GPU: 28 bytes per pose go
    to the device, one bit per pose comes back (C-ABI fc_clash_batch_pose7).  The pose a row stands for is the
    FP64 expansion documented in include/firecode_b200.h (oracle: port.pose7_to_xf); ``mask[p]`` then equals the
    reference's ``compenetration_check(concat(a, R @ b + t), ids=[n_a, n_b], thresh, max_clashes)``."""
    lib = _lib.load(require_device=True)
    a = _as_ensemble(frag_a)
    b = _as_ensemble(frag_b)
    pose7 = np.asarray(pose7)
    assert pose7.dtype == np.float32 and pose7.ndim == 2 and pose7.shape[1] == 7 and pose7.flags.c_contiguous, \
        "pose7 must be a C-contiguous (n, 7) float32 array"
    n = pose7.shape[0]
    tiles = None
    if conf_a is not None or conf_b is not None:
        ca = np.zeros(n, dtype=np.int64) if conf_a is None else np.asarray(conf_a)
        cb = np.zeros(n, dtype=np.int64) if conf_b is None else np.asarray(conf_b)
        assert len(ca) == n and len(cb) == n
        assert n == 0 or (ca.min() >= 0 and ca.max() < a.shape[0] and cb.min() >= 0 and cb.max() < b.shape[0])
        tiles = build_tiles(ca, cb, b.shape[1])
    n_words = (n + 31) // 32
    bits = np.empty(n_words, dtype=np.uint32) if bits_out is None else bits_out
    assert bits.dtype == np.uint32 and bits.size >= n_words and bits.flags.c_contiguous
    status = np.empty(n, dtype=np.uint8) if want_status else None
    counts = np.zeros(3, dtype=np.int64)
    near_idx = np.zeros(max(near_cap, 1), dtype=np.int64)
    near_dist = np.zeros(max(near_cap, 1), dtype=np.float64)

    def ptr(arr):
        return None if arr is None else arr.ctypes.data_as(C.c_void_p)

    rc = lib.fc_clash_batch_pose7(ptr(a), a.shape[0], a.shape[1], ptr(b), b.shape[0], b.shape[1], ptr(pose7), n,
                                  ptr(tiles), 0 if tiles is None else len(tiles), float(thresh), int(max_clashes),
                                  1 if strict else 0, ptr(bits), ptr(status), ptr(counts), ptr(near_idx),
                                  ptr(near_dist), int(near_cap))
    _lib.check(rc, "fc_clash_batch_pose7")
    n_near = int(min(counts[2], near_cap))
    order = np.argsort(near_idx[:n_near], kind="stable")
    return ClashBitsResult(bits=bits[:n_words], n_poses=n, n_pass=int(counts[0]), n_rechecked=int(counts[1]),
                           n_near=int(counts[2]), near_idx=near_idx[:n_near][order],
                           near_dist=near_dist[:n_near][order], status=status)


class DevicePrep:
    """Fragment A's device tables (C-ABI fc_clash_prepare_dev), reusable by several screens on one stream."""

    def __init__(self, a_dev, thresh, want_cells=True):
        import torch

        self.lib = _lib.load(require_device=True)
        if a_dev.dim() == 2:
            a_dev = a_dev[None]
        assert a_dev.is_cuda and a_dev.dtype == torch.float64 and a_dev.is_contiguous()
        self.a_dev = a_dev
        self.stream = torch.cuda.current_stream(a_dev.device).cuda_stream
        h = C.c_void_p()
        _lib.check(self.lib.fc_clash_prepare_dev(a_dev.data_ptr(), a_dev.shape[0], a_dev.shape[1], float(thresh),
                                                 1 if want_cells else 0, C.byref(h), self.stream), "fc_clash_prepare_dev")
        self.handle = h

    def free(self):
        if self.handle:
            self.lib.fc_clash_prep_free(self.handle, self.stream)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def screen_device_ex(prep, b_dev, poses_dev, pose_format=POSE_XF64, max_clashes=0, strict=True, tiles_dev=None,
                     status_out=None, bits_out=None, near=None, pose_index_base=0, recheck_count=None):
    """Device-resident screen on prepared tables (C-ABI fc_clash_screen_ex_dev): poses (n, 12) f64 or (n, 7) f32
    CUDA tensors; writes status bytes and / or the survivor bitmask; runs on the current torch stream."""
    import torch

    lib = prep.lib
    if b_dev.dim() == 2:
        b_dev = b_dev[None]
    assert b_dev.is_cuda and b_dev.dtype == torch.float64 and b_dev.is_contiguous() and poses_dev.is_contiguous()
    if pose_format == POSE_Q7:
        assert poses_dev.dtype == torch.float32 and poses_dev.shape[1] == 7
    else:
        assert poses_dev.dtype == torch.float64 and poses_dev.shape[1] == 12
    assert status_out is not None or bits_out is not None
    n = poses_dev.shape[0]
    stream = torch.cuda.current_stream(poses_dev.device).cuda_stream
    ncount = nidx = ndist = None
    cap = 0
    if near is not None:
        ncount, nidx, ndist = (t.data_ptr() for t in near)
        cap = near[1].numel()
    rc = lib.fc_clash_screen_ex_dev(
        prep.handle, prep.a_dev.data_ptr(), b_dev.data_ptr(), b_dev.shape[0], b_dev.shape[1], poses_dev.data_ptr(),
        int(pose_format), n, None if tiles_dev is None else tiles_dev.data_ptr(),
        0 if tiles_dev is None else tiles_dev.shape[0], int(max_clashes), 1 if strict else 0,
        None if status_out is None else status_out.data_ptr(), None if bits_out is None else bits_out.data_ptr(),
        None, ncount, nidx, ndist, cap, int(pose_index_base),
        None if recheck_count is None else recheck_count.data_ptr(), stream)
    _lib.check(rc, "fc_clash_screen_ex_dev")
    return status_out if status_out is not None else bits_out


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
