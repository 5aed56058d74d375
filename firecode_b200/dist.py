"""Multi-GPU sharding of the embedding screen: one process per GPU (torch.distributed, NCCL over
NVLink on the GPU box, gloo in the CPU tests).

The path shards over independent units: contiguous ranges of the candidate-tuple
index per rank, fragments replicated.  No data-path collective is needed for the clash screen; the
only exchange steps are
  * one all-gather of the per-rank survivor bitmasks (10 M poses = 1.25 MB in total), and
  * for the string embed, one all-gather of the survivors' torsion fingerprints followed by the
    ordered keep-first sweep, run redundantly on every rank so that the kept set is the
    reference's and identical everywhere.
Rank order = enumeration order, so concatenation reproduces the reference's pose order.
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous [lo, hi) range of rank ``rank``; the first n_items % world ranks get one more."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    return dist


def world_info(group=None):
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _device_for_backend(group=None):
    import torch

    dist = _dist()
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather_varlen(x: np.ndarray, group=None) -> np.ndarray:
    """All-gather numpy arrays whose first dimension differs per rank; returns their concatenation
    in rank order (every rank gets the same result)."""
    import torch

    dist = _dist()
    rank, world = world_info(group)
    x = np.ascontiguousarray(x)
    if world == 1:
        return x
    dev = _device_for_backend(group)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([x.shape[0]], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.cpu().numpy()
    row = int(np.prod(x.shape[1:])) if x.ndim > 1 else 1
    cap = int(counts.max())
    flat = np.zeros((cap, row), dtype=x.dtype)
    flat[: x.shape[0]] = x.reshape(x.shape[0], row)
    # bytes on the wire: dtype-agnostic and valid for both NCCL and gloo
    send = torch.from_numpy(flat.view(np.uint8).reshape(-1)).to(dev)
    recv = torch.empty(world * send.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = recv.cpu().numpy().reshape(world, cap, row * x.dtype.itemsize)
    out = [parts[r, : counts[r]].copy().view(x.dtype).reshape((counts[r],) + x.shape[1:]) for r in range(world)]
    return np.concatenate(out, axis=0)


def pack_bits(mask: np.ndarray) -> np.ndarray:
    """bool (n,) -> uint32 (ceil(n/32),), bit (i & 31) of word i >> 5 (layout of fc_pack_mask_dev)."""
    m = np.asarray(mask, dtype=bool)
    pad = (-len(m)) % 32
    m = np.concatenate([m, np.zeros(pad, dtype=bool)])
    return np.packbits(m.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel()


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    w = np.ascontiguousarray(words, dtype=np.uint32)
    return np.unpackbits(w.view(np.uint8).reshape(-1, 4), axis=1, bitorder="little").ravel()[:n].astype(bool)


def all_gather_mask(local_mask: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """Each rank holds the survivor mask of its shard_bounds range; returns the mask of all n_total
    poses on every rank (bitmask all-gather + ordered concatenation)."""
    rank, world = world_info(group)
    if world == 1:
        return np.asarray(local_mask, dtype=bool)
    words = all_gather_varlen(pack_bits(local_mask), group)
    out = np.zeros(n_total, dtype=bool)
    off = 0
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        nw = (hi - lo + 31) // 32
        out[lo:hi] = unpack_bits(words[off: off + nw], hi - lo)
        off += nw
    return out


def clash_screen_sharded(frag_a, frag_b, xf, thresh, max_clashes=0, group=None, screen=None):
    """Screen poses on all ranks: rank r takes poses shard_bounds(n, world, r).  ``screen`` is the
    per-shard screen (default: the CUDA path, firecode_b200.clash.compenetration_check_batch)."""
    rank, world = world_info(group)
    xf = np.asarray(xf, dtype=np.float64).reshape(-1, 12)
    lo, hi = shard_bounds(len(xf), world, rank)
    if screen is None:
        from .clash import compenetration_check_batch

        def screen(a, b, x):
            return compenetration_check_batch(a, b, x, thresh=thresh, max_clashes=max_clashes).mask
    local = screen(frag_a, frag_b, xf[lo:hi]) if hi > lo else np.zeros(0, dtype=bool)
    return all_gather_mask(local, len(xf), group)


def ordered_keep_first(labels, fingerprints, keep_fn, group=None):
    """Merge step of the sharded string embed: all ranks contribute the (pose index, fingerprint)
    of their clash survivors; every rank gets all of them in pose order and runs the ordered
    keep-first sweep ``keep_fn(fingerprints, labels) -> bool mask``."""
    all_labels = all_gather_varlen(np.asarray(labels, dtype=np.int64), group)
    all_fp = all_gather_varlen(np.asarray(fingerprints, dtype=np.float64), group)
    assert np.all(np.diff(all_labels) > 0), "rank order must equal enumeration order"
    keep = np.asarray(keep_fn(all_fp, all_labels), dtype=bool)
    return all_labels[keep], all_labels, keep


def string_embed_sharded(embedder, group=None):
    """string_embed over all ranks of ``group`` (same result on every rank, equal to the
    single-GPU / reference result)."""
    import ctypes as C

    from . import _lib, embeds, problem
    from .errors import ZeroCandidatesError

    rank, world = world_info(group)
    if world == 1:
        return embeds.string_embed(embedder)
    lib = _lib.load(require_device=True)
    prob = problem.string_problem(embedder)
    c, keep_alive = embeds._string_problem_c(prob)
    lo, hi = shard_bounds(prob.n_poses, world, rank)
    handle = C.c_void_p()
    _lib.check(lib.fc_string_stage1(C.byref(c), lo, hi, C.byref(handle)), "fc_string_stage1")
    res = embeds._Result(lib, handle)
    surv, fps = res.survivors(), res.fingerprints()
    res.close()

    def keep_fn(fp, labels):
        out = np.zeros(len(labels), dtype=np.uint8)
        n_ties = C.c_int64(0)
        fp = np.ascontiguousarray(fp.reshape(len(labels), -1))
        _lib.check(lib.fc_tfd_keepfirst(embeds._ptr(fp), embeds._ptr(labels), len(labels), fp.shape[1] if fp.size else 0,
                                        10.0, embeds._ptr(out), None, 0, C.byref(n_ties)), "fc_tfd_keepfirst")
        return out.astype(bool)

    kept, _, _ = ordered_keep_first(surv, fps.reshape(len(surv), -1), keep_fn, group)
    if len(kept) == 0:
        raise ZeroCandidatesError("string embed: no pose survived")
    n_tot = prob.coords[0].shape[1] + prob.coords[1].shape[1]
    poses = np.zeros((len(kept), n_tot, 3))
    kept = np.ascontiguousarray(kept, dtype=np.int64)
    _lib.check(lib.fc_string_materialize(C.byref(c), embeds._ptr(kept), len(kept), embeds._ptr(poses)),
               "fc_string_materialize")
    embedder.constrained_indices = np.repeat(prob.constrained[None], len(poses), axis=0)
    del keep_alive
    return poses


def cyclical_embed_sharded(embedder, group=None, screen=None, max_norm_delta=5.0):
    """cyclical_embed (embeds.py:180-585) over all ranks of ``group``.

    Units are whole groups: conformer-triple ranges for three molecules (the stateful direction
    search never crosses a conformer triple), slices of the host group table for two.  Each rank
    screens its units, then kept pose indices, coordinates and constrained indices are all-gathered
    in rank order = the reference's enumeration order; every rank returns the same arrays.
    ``screen(prob, lo, hi) -> (poses, constrained, n_poses_screened, kept_local)`` replaces the CUDA
    path in the CPU tests."""
    from . import embeds, problem
    from .errors import ZeroCandidatesError

    rank, world = world_info(group)
    prob = problem.cyclical_problem(embedder, max_norm_delta=max_norm_delta)
    if prob.n_mols == 3:
        n_units = int(np.prod([len(c) for c in prob.coords]))
        if screen is None:
            def screen(pb, lo, hi):
                poses, cons, rep = embeds.cyclical3_screen(pb, conf_tuple_range=(lo, hi), want_status=False)
                return poses, cons, rep.n_poses, rep.kept_indices
    else:
        groups = embeds.cyclical_groups(prob)
        n_units = len(groups["conf"])
        if screen is None:
            def screen(pb, lo, hi):
                poses, cons, rep = embeds.cyclical_screen(pb, group_range=(lo, hi), groups=groups)
                return poses, cons, rep.n_poses, rep.kept_indices
    lo, hi = shard_bounds(n_units, world, rank)
    n_tot = sum(int(c.shape[1]) for c in prob.coords)
    if hi > lo:
        poses, cons, n_local, kept_local = screen(prob, lo, hi)
    else:
        poses, cons, n_local, kept_local = (np.zeros((0, n_tot, 3)), np.zeros((0, prob.n_mols, 2), dtype=np.int64), 0,
                                            np.zeros(0, dtype=np.int64))
    counts = all_gather_varlen(np.array([n_local], dtype=np.int64), group)
    base = int(counts[:rank].sum())
    kept = all_gather_varlen(np.asarray(kept_local, dtype=np.int64) + base, group)
    poses = all_gather_varlen(np.asarray(poses, dtype=np.float64).reshape(-1, n_tot, 3), group)
    cons = all_gather_varlen(np.asarray(cons, dtype=np.int64).reshape(len(kept_local), prob.n_mols, 2), group)
    embedder.constrained_indices = cons
    embedder.b200_kept_indices = kept
    if len(poses) == 0:
        raise ZeroCandidatesError("cyclical embed: no pose survived")
    return poses


class _RawDeviceBytes:
    """A device address range as an object torch.as_tensor can wrap without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_all_gather(group=None):
    """The fc_allgather_dev_fn of the C-ABI on top of torch.distributed: gather(send_ptr, recv_ptr, nbytes, stream_ptr)
    leaves the rank-order concatenation of every rank's ``nbytes`` device bytes at ``recv_ptr``, ordered on the
    library's CUDA stream.  NCCL moves the bytes GPU to GPU (NVLink / NVSwitch on the box); with a gloo group (CPU
    tests, several ranks sharing one GPU) the pieces are staged through the host."""
    def gather(send_ptr, recv_ptr, nbytes, stream_ptr):
        import torch

        dist = _dist()
        backend = dist.get_backend(group)
        world = dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        send = torch.as_tensor(_RawDeviceBytes(send_ptr, nbytes), device=dev)
        recv = torch.as_tensor(_RawDeviceBytes(recv_ptr, nbytes * world), device=dev)
        ext = torch.cuda.ExternalStream(stream_ptr, device=dev)
        with torch.cuda.stream(ext):
            if backend == "nccl":
                dist.all_gather_into_tensor(recv, send, group=group)
            else:
                ext.synchronize()
                out = torch.empty(world * nbytes, dtype=torch.uint8)
                dist.all_gather_into_tensor(out, send.cpu(), group=group)
                recv.copy_(out)
                ext.synchronize()

    return gather


# Below this many pairs a pruning call is replicated on every rank instead of sharded.  Measured on 8 B200 at BASELINE
# config C4 (2e10 pairs, 5e9 decided by the screen): device-resident sharding returns the mask in 24 ms against 36 ms
# on one GPU (round 1, host-staged: 0.164 s against 0.085 s), so the break-even lies near that size.
PRUNE_SHARD_MIN_PAIRS = 1e10


def prune_sharded(structures, atoms, kind="rmsd", group=None, force_shard=False, host_staged=False, **kw):
    """prune_by_rmsd / prune_by_moment_of_inertia over all ranks of ``group``: work items of every pass
    are dealt round-robin to the ranks, the similar pairs each rank finds are all-gathered (8 bytes per
    pair) and every rank resolves the pass on the union -- the deterministic ordered merge -- so all ranks
    return the same (structures[mask], mask) as a single-GPU call.  By default the exchange steps run on the
    GPUs (C-ABI fc_prune_sharded_dev through ``device_all_gather``: every rank uploads 1 / world of the
    structures, pieces / counters / pair lists travel over NCCL); ``host_staged=True`` keeps the structures
    replicated and stages the lists through the host.  ``want_structures=False`` (forwarded to the pruning
    call) returns (None, mask): on a multi-GPU box every rank would otherwise write its own copy of the kept
    structures through the one host memory the ranks share.
    Ensembles with fewer than PRUNE_SHARD_MIN_PAIRS pairs are pruned redundantly on every rank instead
    (same result, no collectives) unless ``force_shard``."""
    from . import pruner

    rank, world = world_info(group)
    fn = pruner.prune_by_rmsd if kind == "rmsd" else pruner.prune_by_moment_of_inertia
    n = len(structures)
    if world == 1 or (not force_shard and 0.5 * n * (n - 1) < PRUNE_SHARD_MIN_PAIRS):
        return fn(structures, atoms, **kw)
    if host_staged:
        return fn(structures, atoms, shard=(rank, world, lambda buf: all_gather_varlen(buf, group)), **kw)
    return fn(structures, atoms, shard=(rank, world, None, device_all_gather(group)), **kw)


def torsion_scan_sharded(coords, torsions, masks, angles, group=None, scan=None, **kw):
    """Torsion scan (rotate_dihedral + torsion_comp_check over conformer x torsion x angle) with the
    conformers sharded over the ranks; the pass masks are all-gathered in rank order (= item order).
    Returns the (n_conf, n_tors, n_angles) bool array on every rank."""
    rank, world = world_info(group)
    coords = np.asarray(coords, dtype=np.float64)
    lo, hi = shard_bounds(len(coords), world, rank)
    if scan is None:
        from .torsion import torsion_scan

        def scan(x):
            return torsion_scan(x, torsions, masks, angles, want_coords=False, **kw)["passed"]
    shape = (len(torsions), len(np.atleast_1d(angles)))
    local = np.asarray(scan(coords[lo:hi]), dtype=bool).reshape((hi - lo,) + shape) if hi > lo else np.zeros((0,) + shape, bool)
    flat = all_gather_varlen(local.reshape(-1).astype(np.uint8), group).astype(bool)  # rank order = conformer order
    return flat.reshape((len(coords),) + shape)
