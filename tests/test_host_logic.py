"""CPU tests of the host-side logic of the product (no GPU, no reference tree)."""

import numpy as np
import pytest
from networkx import Graph

from firecode_b200 import embeds, graphs, problem, synthetic, torsion, utils
from firecode_b200.errors import TriangleError
from synth_embedder import make_embedder


def test_cartesian_product_order():
    # first index fastest for two inputs
    out = utils.cartesian_product(range(3), range(2))
    assert out.tolist() == [[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1]]
    out3 = utils.cartesian_product(range(2), range(2), range(2))
    assert out3.tolist() == [[0, 0, 0], [0, 0, 1], [1, 0, 0], [1, 0, 1], [0, 1, 0], [0, 1, 1], [1, 1, 0], [1, 1, 1]]


def test_rotation_helpers():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = rng.normal(size=3), rng.normal(size=3)
        r = utils.rotation_matrix_from_vectors(a, b)
        assert np.allclose(r @ r.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(r), 1)
        assert np.allclose(r @ (a / np.linalg.norm(a)), b / np.linalg.norm(b), atol=1e-12)
    z = utils.rot_mat_from_pointer(np.array([0, 0, 2.0]), 90)
    assert np.allclose(z @ np.array([1.0, 0, 0]), [0, 1, 0])  # right-handed about +z
    a = np.array([1.0, 0, 0])
    assert np.allclose(utils.rotation_matrix_from_vectors(a, a), np.eye(3))
    assert np.allclose(utils.rotation_matrix_from_vectors(a, -a) @ a, -a)


def test_polygonize():
    two = utils.polygonize([2.0, 3.0])
    assert two.shape == (2, 2, 2, 3)
    assert np.allclose(two[0, 1], [[-1.5, 0, 0], [1.5, 0, 0]]) and np.allclose(two[1, 1], [[1.5, 0, 0], [-1.5, 0, 0]])
    tri = utils.polygonize([3.0, 4.0, 5.0])
    assert tri.shape == (8, 3, 2, 3)
    for t in tri:  # every orientation is the same closed triangle
        lengths = sorted(np.linalg.norm(t[:, 1] - t[:, 0], axis=1))
        assert np.allclose(lengths, [3, 4, 5])
    with pytest.raises(TriangleError):
        utils.polygonize([1.0, 1.0, 3.0])


def test_string_problem_extraction_and_decode():
    emb = make_embedder("string", 3, 20, seed=1, n_orb=2)
    prob = problem.string_problem(emb)
    assert prob.n_poses == 3 * 3 * 4 * 36
    assert prob.centers[0].shape == (3, 2, 3) and prob.quadruplets.shape[1] == 4
    # decode follows the reference's loop nest: conformer pairs (first index fastest), centre pairs, angles
    seen = [prob.decode(p)[:4] for p in range(0, prob.n_poses, 36)]
    assert seen[0] == (0, 0, 0, 0) and seen[1] == (0, 0, 1, 0) and seen[4] == (1, 0, 0, 0)
    assert prob.decode(37)[4] == 10.0


def test_cyclical_groups_table():
    emb = make_embedder("cyclical", 2, 16, seed=4, n_reactive=2, n_orb=2)
    prob = problem.cyclical_problem(emb)
    g = embeds.cyclical_groups(prob)
    assert g["conf"].shape[1] == 2 and g["vecs"].shape[1:] == (2, 2, 3) and len(g["conf"]) == 2 * 2 * 16 * 2
    # a pairing that no arrangement contains removes every group (the reference's quirk N10 included)
    emb.pairings_table = {"a": (0, 1)}
    prob2 = problem.cyclical_problem(emb)
    assert len(embeds.cyclical_groups(prob2)["conf"]) == 0


def test_rotation_mask_and_tiles():
    g = Graph([(0, 1), (1, 2), (2, 3), (3, 4), (2, 5)])
    mask = torsion.get_rotation_mask(g, (0, 1, 2, 3))
    assert mask.tolist() == [False, False, False, True, True, True]
    from firecode_b200 import clash

    ca = np.array([0, 0, 0, 1, 1, 2]); cb = np.array([0, 0, 1, 1, 1, 0])
    tiles = clash.build_tiles(ca, cb, 150)
    assert tiles[:, 2].tolist() == [0, 2, 3, 5] and tiles[:, 3].tolist() == [2, 1, 2, 1]
    p = clash.tile_poses(150)
    big = clash.build_tiles(np.zeros(3 * p + 1, int), np.zeros(3 * p + 1, int), 150)
    assert big[:, 3].tolist() == [p, p, p, 1]


def test_sum_graph_and_quadruplets():
    g1 = Graph([(0, 1), (1, 2)]); g2 = Graph([(0, 1), (1, 2)])
    s = graphs.sum_graph((g1, g2), [(2, 3)])
    assert sorted(s.edges()) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]
    q = graphs.quadruplets(s)
    assert q.shape == (3, 4)


def test_synthetic_generators_are_seeded():
    a = synthetic.conformer_ensemble(np.random.default_rng(3), 4, 30)[1]
    b = synthetic.conformer_ensemble(np.random.default_rng(3), 4, 30)[1]
    assert np.array_equal(a, b) and abs(a.reshape(-1, 3).mean(axis=0)).max() < 1e-9
    xf = synthetic.sweep_poses(np.random.default_rng(1), a[0], a[1], 10)
    assert np.array_equal(xf, xf.astype(np.float32).astype(np.float64))


def test_vectorised_bimolecular_group_table_equals_loop_restatement():
    """embeds.cyclical_groups (numpy, all groups at once) against the loop restatement of
    embeds.py:596-641 in oracle.port, with and without a user pairing."""
    from firecode_b200 import embeds, problem
    from firecode_b200.utils import polygonize
    from oracle import port
    from synth_embedder import make_embedder

    for kw in (dict(n_conf=3, n_atoms=24, seed=11, n_reactive=2, n_orb=2),
               dict(n_conf=[4, 2], n_atoms=20, seed=13, n_reactive=1, n_orb=2)):
        emb = make_embedder("cyclical", **kw)
        prob = problem.cyclical_problem(emb)
        first = [tuple(int(x) for x in port.cyclical_groups_bimol(prob)[1]["ids"][0])]
        for pairings in ([], first):
            prob.pairings = pairings
            table = embeds.cyclical_groups(prob)
            ref = port.cyclical_groups_bimol(prob)
            assert len(ref) == len(table["conf"]) > 0
            for i, r in enumerate(ref):
                assert tuple(table["conf"][i]) == r["conf"]
                assert [tuple(x) for x in table["ids"][i]] == [tuple(x) for x in r["ids"]]
                assert np.array_equal(table["vecs"][i], polygonize(r["norms"])[r["v"]])
                for m in range(2):
                    assert np.array_equal(table["pivot"][i][m], prob.pivot_vec[m][r["conf"][m]][r["piv"][m]])
                    assert np.array_equal(table["mean"][i][m], prob.pivot_mean[m][r["conf"][m]][r["piv"][m]])


def test_take_rows_matches_boolean_indexing():
    """fc_take_rows (threaded gather behind the `structures[mask]` the pruning entry points return) is host-only."""
    import ctypes as C

    from firecode_b200 import _lib
    from firecode_b200.pruner import _take

    lib = _lib.load(require_device=False)
    rng = np.random.default_rng(0)
    for n, shape in ((0, (5, 3)), (1, (5, 3)), (1000, (7, 3)), (4099, (1, 3))):
        x = rng.normal(size=(n,) + shape)
        mask = rng.random(n) < 0.4
        assert np.array_equal(_take(lib, x, mask), x[mask])
    x = rng.normal(size=(10, 4, 3))
    out = np.empty((3, 4, 3))
    m = np.ones(10, dtype=np.uint8)
    rc = lib.fc_take_rows(x.ctypes.data_as(C.c_void_p), 96, m.ctypes.data_as(C.c_void_p), 10,
                          out.ctypes.data_as(C.c_void_p), 3)
    assert rc != 0 and b"selects 10 rows" in lib.fc_last_error()


def _ref_write_xyz(atoms, coords, title):
    """The reference's formatting, restated for boxes without /root/reference (utils.py:105-116)."""
    s = str(len(coords)) + f"\n{title}\n"
    for atom, c in zip(atoms, coords):
        s += "%s     % .6f % .6f % .6f\n" % (atom, c[0], c[1], c[2])
    return s


def test_write_xyz_is_byte_identical():
    import io

    from firecode_b200 import utils

    rng = np.random.default_rng(4)
    atoms = np.array(["C", "H", "Cl", "Br", "N", "O", "H"])
    structures = rng.normal(size=(700, 7, 3)) * np.array([1.0, 100.0, 1e-4])
    structures[0, 0] = [0.0, -0.0, 1e-7]
    structures[1, 1] = [123456.7890125, -0.0000005, 0.0000005]
    structures[2, 2] = [1e15, -2.5e-7, 9.9999995]
    titles = [f"structure {i} E = {i * 0.37:.3f}" for i in range(len(structures))]
    want = "".join(_ref_write_xyz(atoms, s, t) for s, t in zip(structures, titles))
    assert utils.xyz_text(atoms, structures, titles) == want
    buf = io.StringIO()
    utils.write_xyz(atoms, structures[5], buf, title="temp")
    assert buf.getvalue() == _ref_write_xyz(atoms, structures[5], "temp")
    buf = io.StringIO()
    utils.write_xyz_batch(atoms, structures[:3], buf)
    assert buf.getvalue() == "".join(_ref_write_xyz(atoms, s, "temp") for s in structures[:3])
    assert utils.xyz_text(atoms, structures[:0]) == ""


@pytest.mark.reference
def test_write_xyz_equals_live_reference():
    import io

    from firecode_b200 import utils
    from oracle import loader

    loader.install()
    from firecode.utils import write_xyz as ref_write_xyz

    rng = np.random.default_rng(5)
    atoms = np.array(["C", "H", "Si", "O"])
    for i in range(20):
        coords = rng.normal(size=(4, 3)) * 10.0 ** rng.integers(-3, 6)
        a, b = io.StringIO(), io.StringIO()
        ref_write_xyz(atoms, coords, a, title=f"t{i}")
        utils.write_xyz(atoms, coords, b, title=f"t{i}")
        assert a.getvalue() == b.getvalue()


def _plan(mask, k, prev_k, world=1, rank=0, n_sms=4, over_active=False):
    import ctypes as C

    from firecode_b200 import _lib

    lib = _lib.load(require_device=False)
    n = len(mask)
    m8 = np.ascontiguousarray(mask, dtype=np.uint8)
    spos = np.zeros(2 * n + 16 * k + 1024, dtype=np.int32)
    work = np.zeros((n * 4 + 4096, 4), dtype=np.int32)
    n_spos, n_work = C.c_int64(0), C.c_int64(0)
    counts = np.zeros(2, dtype=np.int64)
    rc = lib.fc_prune_plan(m8.ctypes.data, n, k, prev_k, world, rank, n_sms, 1 if over_active else 0, spos.ctypes.data, len(spos),
                           C.byref(n_spos), work.ctypes.data, len(work), C.byref(n_work), counts.ctypes.data)
    assert rc == 0, lib.fc_last_error()
    return spos[: n_spos.value], work[: n_work.value], counts


@pytest.mark.parametrize("over_active", [False, True])
@pytest.mark.parametrize("n,k,prev_k,world", [(700, 1, 0, 1), (700, 1, 2, 1), (3000, 5, 10, 3), (3000, 2, 5, 2),
                                              (2500, 10, 20, 1), (1000, 1, 2, 8), (333, 3, 0, 2)])
def test_prune_planner_covers_every_unknown_pair_once(n, k, prev_k, world, over_active):
    """The work items of the tensor-core screen (row block x column-tile range, dealt to the ranks) must cover every
    pair of active structures of a chunk exactly once, except pairs whose two members shared a chunk of the
    previous pass (known dissimilar), which may be left out -- and nothing else may be left out."""
    rng = np.random.default_rng(n + k)
    mask = rng.random(n) < 0.6
    def membership(kk):
        """chunk of every structure: kk chunks of n // kk structures, or of n_active // kk ACTIVE structures"""
        if not over_active:
            return np.minimum(np.arange(n) // (n // kk), kk - 1)
        act_idx = np.flatnonzero(mask)
        sz = max(1, len(act_idx) // kk)
        starts = np.array([0] + [act_idx[c * sz] if c * sz < len(act_idx) else n for c in range(1, kk)])
        return np.searchsorted(starts, np.arange(n), side="right") - 1

    chunk_of = membership(k)
    prev_of = membership(prev_k) if prev_k else None
    covered = {}
    tiled_total = 0
    for rank in range(world):
        spos, work, counts = _plan(mask, k, prev_k, world, rank, over_active=over_active)
        tiled_total += int(counts[0])
        # position list: chunks start at multiples of 16, hold their active structures in order, -1 elsewhere
        pos_active = np.flatnonzero(spos >= 0)
        assert np.array_equal(spos[pos_active], np.flatnonzero(mask))
        assert len(spos) % 16 == 0 and np.all(spos[-128:] == -1)
        for c in range(k):
            members = pos_active[chunk_of[spos[pos_active]] == c]
            if len(members):
                assert members[0] % 16 == 0 and np.array_equal(members, np.arange(members[0], members[0] + len(members)))
        for row0, ct0, nt, pend in work:
            assert row0 % 8 == 0 and nt >= 1 and spos[row0] >= 0
            for r in range(row0, min(row0 + 128, pend)):
                lo, hi = max(16 * ct0, r + 1), min(16 * (ct0 + nt), pend)
                for c in range(lo, hi):
                    key = (int(spos[r]), int(spos[c]))
                    assert key[0] >= 0 and key[1] >= 0 and chunk_of[key[0]] == chunk_of[key[1]]
                    assert key not in covered
                    covered[key] = rank
    assert tiled_total == len(covered)
    act = np.flatnonzero(mask)
    missing = 0
    for c in range(k):
        mem = act[chunk_of[act] == c]
        for a_i in range(len(mem)):
            for b_i in range(a_i + 1, len(mem)):
                if (int(mem[a_i]), int(mem[b_i])) not in covered:
                    missing += 1
                    assert prev_of is not None and prev_of[mem[a_i]] == prev_of[mem[b_i]]
    assert missing == int(counts[1])
    if world > 1:
        assert len(set(covered.values())) >= 2          # the items are dealt to the ranks


@pytest.mark.parametrize("over_active", [False, True])
@pytest.mark.parametrize("n,k,prev_k", [(700, 1, 0), (700, 1, 2), (3000, 5, 10), (3000, 2, 5), (2500, 10, 20), (333, 3, 0)])
def test_prune_segments_may_be_reordered(n, k, prev_k, over_active):
    """The tile culling of the tensor-core screen sorts the positions of every SEGMENT by shape on the device.  That is
    only legal if (1) a segment is a contiguous run of positions of one chunk whose structures shared a chunk of the
    previous pass, (2) sorting by (segment, anything) leaves the padding where it is, and (3) the work items still cover
    every unknown pair exactly once after an arbitrary permutation inside the segments."""
    import ctypes as C

    from firecode_b200 import _lib

    rng = np.random.default_rng(7 * n + k)
    mask = rng.random(n) < 0.6
    spos, work, counts = _plan(mask, k, prev_k, over_active=over_active)
    lib = _lib.load(require_device=False)
    seg = np.zeros(len(spos) + 64, dtype=np.int32)
    n_seg = C.c_int64(0)
    m8 = np.ascontiguousarray(mask, dtype=np.uint8)
    assert lib.fc_prune_plan_segments(m8.ctypes.data, n, k, prev_k, 1 if over_active else 0, seg.ctypes.data, len(seg),
                                      C.byref(n_seg)) == 0, lib.fc_last_error()
    assert n_seg.value == len(spos)
    seg = seg[: len(spos)]
    assert np.all(np.diff(seg) >= 0)                                   # (2) a stable sort by segment keeps the runs in place ...
    real = spos >= 0
    for sid in np.unique(seg):
        run = np.flatnonzero(seg == sid)
        assert np.array_equal(run, np.arange(run[0], run[-1] + 1))     # (1) contiguous
        r = real[run]
        assert not np.any(r[np.argmin(r):]) or r.all()                 # ... and padding sits behind the structures of its run

    def membership(kk):
        if not over_active:
            return np.minimum(np.arange(n) // (n // kk), kk - 1)
        act_idx = np.flatnonzero(mask)
        sz = max(1, len(act_idx) // kk)
        starts = np.array([0] + [act_idx[c * sz] if c * sz < len(act_idx) else n for c in range(1, kk)])
        return np.searchsorted(starts, np.arange(n), side="right") - 1

    chunk_of = membership(k)
    prev_of = membership(prev_k) if prev_k else None
    for sid in np.unique(seg[real]):
        members = spos[(seg == sid) & real]
        assert len(set(chunk_of[members])) == 1
        if prev_of is not None:
            assert len(set(prev_of[members])) == 1                     # (1) one chunk of the previous pass
    # (3) shuffle the structures inside every segment and count the pairs the work items visit
    shuffled = spos.copy()
    for sid in np.unique(seg[real]):
        at = np.flatnonzero((seg == sid) & real)
        shuffled[at] = rng.permutation(spos[at])
    covered = set()
    for row0, ct0, nt, pend in work:
        for r in range(row0, min(row0 + 128, pend)):
            for c in range(max(16 * ct0, r + 1), min(16 * (ct0 + nt), pend)):
                a_, b_ = int(shuffled[r]), int(shuffled[c])
                key = (min(a_, b_), max(a_, b_))
                assert a_ >= 0 and b_ >= 0 and chunk_of[a_] == chunk_of[b_] and key not in covered
                covered.add(key)
    act = np.flatnonzero(mask)
    for c in range(k):
        mem = act[chunk_of[act] == c]
        for a_i in range(len(mem)):
            for b_i in range(a_i + 1, len(mem)):
                if (int(mem[a_i]), int(mem[b_i])) not in covered:
                    assert prev_of is not None and prev_of[mem[a_i]] == prev_of[mem[b_i]]


def test_prune_sharded_replicates_small_ensembles(monkeypatch):
    """dist.prune_sharded shards the pair work only above PRUNE_SHARD_MIN_PAIRS pairs (or when forced); below,
    every rank prunes the whole ensemble with the single-GPU call (same result, no collectives)."""
    from firecode_b200 import dist as fdist
    from firecode_b200 import pruner

    calls = []

    def fake(structures, atoms, **kw):
        calls.append(kw)
        return structures, np.ones(len(structures), dtype=bool)

    monkeypatch.setattr(pruner, "prune_by_rmsd", fake)
    monkeypatch.setattr(pruner, "prune_by_moment_of_inertia", fake)
    monkeypatch.setattr(fdist, "world_info", lambda group=None: (1, 4))
    x = np.zeros((100, 3, 3))
    fdist.prune_sharded(x, ["C"] * 3, "rmsd", max_rmsd=0.3)
    assert "shard" not in calls[-1] and calls[-1]["max_rmsd"] == 0.3
    fdist.prune_sharded(x, ["C"] * 3, "rmsd", force_shard=True, max_rmsd=0.3)
    rank, world, gather, gather_dev = calls[-1]["shard"]   # default: the exchange steps run on the devices
    assert (rank, world) == (1, 4) and gather is None and callable(gather_dev)
    fdist.prune_sharded(x, ["C"] * 3, "rmsd", force_shard=True, host_staged=True, max_rmsd=0.3)
    rank, world, gather = calls[-1]["shard"]               # host-staged lists (round 1)
    assert (rank, world) == (1, 4) and callable(gather)
    monkeypatch.setattr(fdist, "PRUNE_SHARD_MIN_PAIRS", 1000.0)
    fdist.prune_sharded(x, ["C"] * 3, "moi")
    assert calls[-1]["shard"][:2] == (1, 4)
    monkeypatch.setattr(fdist, "world_info", lambda group=None: (0, 1))
    fdist.prune_sharded(x, ["C"] * 3, "rmsd", force_shard=True)
    assert "shard" not in calls[-1]


def test_kabsch_rotation_is_defined_for_rank_deficient_covariances():
    """three collinear atoms (rank-1 covariance) and a single atom (zero covariance) used to give an
    all-NaN rotation, so exact duplicates of linear species were never pruned.  The rotation must be proper, finite
    and optimal (trace(R^T H) = sum of the singular values) for every rank."""
    import ctypes as C

    from firecode_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(4)

    def kabsch(h):
        h = np.ascontiguousarray(h, dtype=np.float64)
        r = np.zeros(9)
        sig = np.zeros(3)
        assert lib.fc_kabsch_host(h.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), sig.ctypes.data_as(C.c_void_p)) == 0
        return r.reshape(3, 3), sig

    cases = []
    line = np.array([[-1.0, 0, 0], [0, 0, 0], [1.0, 0, 0]])  # three collinear axis-aligned atoms
    cases.append(line.T @ line)
    d = rng.normal(size=3)
    pts = np.outer(np.array([-1.3, 0.2, 1.1]), d / np.linalg.norm(d))
    q = pts @ synth_rot(rng).T
    cases.append(pts.T @ q)                                    # rank 1, generic directions
    cases.append(np.zeros((3, 3)))                             # single atom after centring
    p2 = rng.normal(size=(5, 3)); p2[:, 2] = 0                 # planar: rank 2
    cases.append(p2.T @ (p2 @ synth_rot(rng).T))
    p3 = rng.normal(size=(7, 3))
    cases.append(p3.T @ (p3 @ synth_rot(rng).T))               # full rank
    cases.append(-(p3.T @ p3))                                 # reflection-like (det < 0)
    for h in cases:
        r, sig = kabsch(h)
        assert np.all(np.isfinite(r)) and np.all(np.isfinite(sig))
        assert np.abs(r @ r.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(r) - 1.0) < 1e-12
        s = np.linalg.svd(h, compute_uv=False)
        best = s[0] + s[1] + (s[2] if np.linalg.det(h) >= 0 else -s[2])
        assert abs(np.trace(r.T @ h) - best) <= 1e-10 * max(1.0, s[0])
        assert np.allclose(np.abs(sig), s, atol=1e-7 * max(1.0, s[0]))  # sqrt of eigenvalues of H^T H: ~1e-8 sigma_1


def synth_rot(rng):
    from firecode_b200 import synthetic

    return synthetic.random_rotations(rng, 1)[0]


def test_vectorised_oracle_pruner_equals_the_loop_version():
    """oracle.prism_pruner.pruner.prune_by_rmsd_vectorised (used for the 20 k C4 subset on the GPU box) returns the
    mask of the plain loop restatement, near-threshold pairs included."""
    from firecode_b200 import synthetic
    from oracle.prism_pruner import pruner as ref_pruner

    for seed, n, n_atoms, basins, jitter in ((1, 700, 30, 20, (0.02, 0.5)), (2, 900, 24, 6, (0.05, 0.3)),
                                             (3, 450, 40, 450, (0.0, 0.0))):
        atoms, structures, _ = synthetic.pruning_ensemble(np.random.default_rng(seed), n, n_atoms, basins, jitter=jitter)
        _, slow = ref_pruner.prune_by_rmsd(structures, atoms, 0.5)
        _, fast = ref_pruner.prune_by_rmsd_vectorised(structures, atoms, 0.5, row_block=64)
        assert np.array_equal(slow, fast), seed
        assert 0 < fast.sum() <= n


def test_bimolecular_group_table_equals_the_oracle_enumeration():
    """C-ABI fc_cyclical_groups (host code) against oracle.port.cyclical_groups_bimol: same groups in the same order,
    with and without user pairings, with an internal constraint given as a list (matches) or as an ndarray (quirk N10:
    never matches), and with the norm filter biting."""
    from firecode_b200 import embeds, problem
    from firecode_b200.synthetic_embedder import make_embedder
    from oracle import port

    emb = make_embedder("cyclical", [3, 2], [14, 11], seed=5, n_reactive=2, n_orb=2)
    base = problem.cyclical_problem(emb)
    ref0 = port.cyclical_groups_bimol(base)
    assert len(ref0) > 8
    couple = tuple(int(x) for x in ref0[1]["ids"][0])
    other = tuple(int(x) for x in ref0[1]["ids"][1])
    cases = [({}, np.array([]), 5.0), ({"a": couple}, np.array([]), 5.0), ({"a": couple, "b": (0, 1)}, [(0, 1)], 5.0),
             ({"a": couple, "b": (0, 1)}, np.array([(0, 1)]), 5.0), ({"a": couple, "b": other}, np.array([]), 5.0),
             ({}, np.array([]), 1e-3)]
    for table, internal, delta in cases:
        emb.pairings_table, emb.internal_constraints = table, internal
        prob = problem.cyclical_problem(emb, max_norm_delta=delta)
        got = embeds.cyclical_groups(prob)
        ref = port.cyclical_groups_bimol(prob)
        assert len(got["conf"]) == len(ref)
        for g, r in enumerate(ref):
            assert tuple(got["conf"][g]) == r["conf"]
            assert [tuple(x) for x in got["ids"][g]] == [tuple(int(v) for v in c) for c in r["ids"]]
            pv = [prob.pivot_vec[m][r["conf"][m]][r["piv"][m]] for m in range(2)]
            pm = [prob.pivot_mean[m][r["conf"][m]][r["piv"][m]] for m in range(2)]
            assert np.array_equal(got["pivot"][g], np.array(pv)) and np.array_equal(got["mean"][g], np.array(pm))
            n0, n1 = r["norms"]
            want = np.array([[[-n0 / 2, 0, 0], [n0 / 2, 0, 0]], [[-n1 / 2, 0, 0], [n1 / 2, 0, 0]]])
            if r["v"] == 1:
                want[1] *= -1
            assert np.array_equal(got["vecs"][g], want)
            assert np.array_equal(got["dirs"][g], [[0.0, 1.0, 0.0], [0.0, -1.0, 0.0]])
    assert len(embeds.cyclical_groups(prob)["conf"]) < len(ref0)   # the last case: the norm filter dropped groups
