"""GPU parity: CUDA clash screen (C-ABI fc_clash_batch) vs the CPU oracle, bit-exact masks."""

import numpy as np
import pytest

from firecode_b200 import synthetic
from firecode_b200.clash import NEAR_EPS, STATUS_NEAR, compenetration_check_batch
from oracle import port

pytestmark = pytest.mark.gpu


def _frags(rng, n_a, n_b):
    _, a, _, _ = synthetic.molecule_cloud(rng, n_a)
    _, b, _, _ = synthetic.molecule_cloud(rng, n_b)
    return a, b


@pytest.mark.parametrize("n_a,n_b,n_poses", [(150, 150, 8000), (30, 30, 5000), (5, 5, 300),
                                             (61, 47, 3000), (1, 1, 64), (150, 33, 2000),
                                             (7, 150, 2000)])
def test_mask_matches_oracle(gpu, n_a, n_b, n_poses):
    rng = np.random.default_rng(synthetic.SEED + n_a * 1000 + n_b)
    a, b = _frags(rng, max(n_a, 2), max(n_b, 2))
    a, b = a[:n_a], b[:n_b]
    xf = synthetic.sweep_poses(rng, a, b, n_poses)
    res = compenetration_check_batch(a, b, xf, thresh=1.5, want_min_dist=True)
    mask, dmin, closest = port.clash_batch(a, b, xf, thresh=1.5)
    assert res.mask.dtype == bool and res.mask.shape == (n_poses,)
    near = np.abs(dmin - 1.5) <= NEAR_EPS
    # every near-threshold pose is listed; everything else is bit-exact
    assert np.array_equal(res.mask[~near], mask[~near])
    assert set(np.flatnonzero(near)) <= set(res.near_idx.tolist())
    assert np.allclose(res.min_dist, dmin, atol=5e-3)
    # clash rate sanity: the sweep is a mixed set
    assert 0.02 < mask.mean() < 0.98 or n_a * n_b < 100


def test_empty_and_single(gpu):
    rng = np.random.default_rng(1)
    a, b = _frags(rng, 20, 20)
    res = compenetration_check_batch(a, b, np.zeros((0, 12)), thresh=1.5)
    assert res.mask.shape == (0,)
    xf = np.array([[1, 0, 0, 0, 1, 0, 0, 0, 1, 100.0, 0, 0]], dtype=float)
    assert compenetration_check_batch(a, b, xf, thresh=1.5).mask.tolist() == [True]
    xf[0, 9] = 0.0
    assert compenetration_check_batch(a, a, xf, thresh=1.5).mask.tolist() == [False]


def test_conformer_tiles(gpu):
    rng = np.random.default_rng(7)
    _, ens_a, _, _ = synthetic.conformer_ensemble(rng, 6, 40)
    _, ens_b, _, _ = synthetic.conformer_ensemble(rng, 5, 33)
    n = 7000
    ca = np.sort(rng.integers(0, 6, size=n))
    cb = rng.integers(0, 5, size=n)
    order = np.lexsort((cb, ca))
    ca, cb = ca[order], cb[order]
    xf = synthetic.sweep_poses(rng, ens_a[0], ens_b[0], n)
    res = compenetration_check_batch(ens_a, ens_b, xf, thresh=1.5, conf_a=ca, conf_b=cb)
    mask, dmin, _ = port.clash_batch(ens_a, ens_b, xf, thresh=1.5, conf_a=ca, conf_b=cb)
    ok = np.abs(dmin - 1.5) > NEAR_EPS
    assert np.array_equal(res.mask[ok], mask[ok])


@pytest.mark.parametrize("max_clashes,strict", [(0, False), (3, True), (10, False)])
def test_count_mode_and_nonstrict(gpu, max_clashes, strict):
    rng = np.random.default_rng(11 + max_clashes)
    a, b = _frags(rng, 60, 60)
    xf = synthetic.sweep_poses(rng, a, b, 4000, shell=(-3.0, 2.0))
    res = compenetration_check_batch(a, b, xf, thresh=1.5, max_clashes=max_clashes, strict=strict)
    mask, _, closest = port.clash_batch(a, b, xf, thresh=1.5, max_clashes=max_clashes, strict=strict)
    ok = closest > NEAR_EPS
    assert np.array_equal(res.mask[ok], mask[ok])


def test_adversarial_near_threshold(gpu):
    """Poses constructed so the closest pair sits at thresh + delta for tiny deltas: the FP64
    recheck must decide them like the oracle and list those within 1e-6 A."""
    rng = np.random.default_rng(2026)
    a, b = _frags(rng, 150, 150)
    thr = 1.5
    base = synthetic.sweep_poses(rng, a, b, 3000, shell=(2.0, 6.0), dtype=np.float64)
    _, dmin, _ = port.clash_batch(a, b, base, thresh=thr)
    out = []
    deltas = [0.0, 1e-7, -1e-7, 5e-7, -5e-7, 1e-5, -1e-5, 3e-4, -3e-4]
    k = 0
    for p in np.flatnonzero(dmin > thr + 0.2)[:600]:
        placed = port.place(b, base[p])
        d = np.linalg.norm(placed[:, None, :] - a[None, :, :], axis=-1)
        j, i = np.unravel_index(np.argmin(d), d.shape)
        u = (a[i] - placed[j]) / d[j, i]
        xf = base[p].copy()
        xf[9:12] += u * (d[j, i] - (thr + deltas[k % len(deltas)]))
        k += 1
        out.append(xf)
    xf = np.array(out)
    res = compenetration_check_batch(a, b, xf, thresh=thr, want_min_dist=True)
    mask, dmin2, closest = port.clash_batch(a, b, xf, thresh=thr)
    gap = np.abs(dmin2 - thr)
    far = gap > 1e-9  # beyond double rounding noise the decision must agree
    assert np.array_equal(res.mask[far], mask[far])
    near = gap <= 0.9 * NEAR_EPS
    assert near.sum() > 100
    assert set(np.flatnonzero(near)) <= set(res.near_idx.tolist())
    assert np.all(res.status[np.flatnonzero(near)] & STATUS_NEAR)
    # listed distances are the FP64 minimum distances
    lut = dict(zip(res.near_idx.tolist(), res.near_dist.tolist()))
    for p in np.flatnonzero(near)[:50]:
        assert abs(lut[p] - dmin2[p]) < 1e-9


def test_single_structure_api(gpu):
    """utils.compenetration_check keeps the reference signature (utils.py:507-513)."""
    from firecode_b200.utils import compenetration_check

    rng = np.random.default_rng(3)
    for _ in range(20):
        coords = rng.normal(scale=2.5, size=(30, 3))
        for ids in ((12, 18), (10, 10, 10)):
            assert compenetration_check(coords, ids=ids, thresh=1.2) == port.compenetration_check(coords, ids=ids, thresh=1.2)
        assert compenetration_check(coords, ids=(12, 18), thresh=1.2, max_clashes=2) == \
            port.compenetration_check(coords, ids=(12, 18), thresh=1.2, max_clashes=2)


def test_nonfragment_branch_and_rmsd_similarity(gpu):
    """utils.compenetration_check(ids=None) (utils.py:523-542), algebra.count_clashes and
    utils.rmsd_similarity (utils.py:494-504) against plain numpy restatements."""
    import networkx as nx
    from scipy.spatial.distance import cdist

    from firecode_b200 import algebra, synthetic
    from firecode_b200.utils import compenetration_check, rmsd_similarity
    from oracle.prism_pruner.rmsd import rmsd_and_max

    rng = np.random.default_rng(11)
    for trial in range(12):
        atoms, coords, bonds, _ = synthetic.molecule_cloud(rng, 25)
        coords = coords + rng.normal(scale=0.25 * (trial % 3), size=coords.shape)
        g = nx.Graph()
        g.add_nodes_from(range(len(coords)))
        g.add_edges_from(bonds)
        d = cdist(coords, coords)
        n_close = int(np.count_nonzero((d < 0.5) & (d > 0)))
        assert algebra.count_clashes(coords) == n_close
        for mc in (0, 3):
            # reference semantics restated: running count tested at the top of each argwhere iteration
            def ref_check(graph):
                if n_close > mc:
                    return False
                if graph is None:
                    return True
                clashes = 0
                for i1, i2 in np.argwhere(d < 1.3):
                    if clashes > mc:
                        return False
                    if i1 != i2 and (i1, i2) not in graph.edges:
                        clashes += 1
                return True
            assert compenetration_check(coords, thresh=1.3, max_clashes=mc) == ref_check(None)
            assert compenetration_check(coords, graph=g, thresh=1.3, max_clashes=mc) == ref_check(g)
    ref = rng.normal(size=(30, 3)) * 3
    structs = np.array([ref + rng.normal(scale=s, size=ref.shape) for s in (0.9, 0.6, 0.2, 1.5)])
    r, m = algebra.rmsd_and_max_batch(ref, structs)
    for k in range(len(structs)):
        rr, mm = rmsd_and_max(ref, structs[k], center=False)
        assert abs(r[k] - rr) < 1e-9 and abs(m[k] - mm) < 1e-9
    rc, mc_ = algebra.rmsd_and_max_batch(ref, structs + 5.0, center=True)
    assert np.allclose(rc, [rmsd_and_max(ref, s + 5.0, center=True)[0] for s in structs], atol=1e-9)
    for thr in (0.1, 0.5, 1.0):
        want = any(rmsd_and_max(ref, s)[0] < thr and rmsd_and_max(ref, s)[1] < 2 * thr for s in structs)
        assert rmsd_similarity(ref, structs, rmsd_thr=thr) == want
    assert rmsd_similarity(ref, np.zeros((0, 30, 3))) is False


@pytest.mark.parametrize("n_a,n_b,thresh", [(150, 150, 1.5), (60, 33, 1.2), (254, 40, 2.0), (17, 150, 0.9)])
def test_cell_list_path_equals_all_pairs_and_oracle(gpu, monkeypatch, n_a, n_b, thresh):
    """The cell-list screen (FC_CLASH_MODE=1) and the all-pairs kernel (FC_CLASH_MODE=0) write identical
    status bytes; both equal the oracle."""
    from firecode_b200 import synthetic

    rng = np.random.default_rng(n_a * 1000 + n_b)
    _, a, _, _ = synthetic.molecule_cloud(rng, n_a)
    _, b, _, _ = synthetic.molecule_cloud(rng, n_b)
    xf = synthetic.sweep_poses(rng, a, b, 40000)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FC_CLASH_MODE", mode)
        for mc in (0, 2):
            out[(mode, mc)] = compenetration_check_batch(a, b, xf, thresh=thresh, max_clashes=mc)
    for mc in (0, 2):
        assert np.array_equal(out[("0", mc)].status & 1, out[("1", mc)].status & 1)
        ref_mask, dmin, closest = port.clash_batch(a, b, xf[:6000], thresh=thresh, max_clashes=mc)
        safe = closest > 1e-6
        assert np.array_equal(out[("1", mc)].mask[:6000][safe], ref_mask[safe])
    assert 0.05 < out[("1", 0)].mask.mean() < 0.95


def test_cell_list_path_with_conformer_tiles(gpu, monkeypatch):
    """Explicit tile lists (several conformers of both fragments) through the cell-list path."""
    from firecode_b200 import synthetic

    rng = np.random.default_rng(5)
    _, ca, _, _ = synthetic.conformer_ensemble(rng, 3, 40, n_torsions=4)
    _, cb, _, _ = synthetic.conformer_ensemble(rng, 2, 35, n_torsions=4)
    xf = synthetic.sweep_poses(rng, ca[0], cb[0], 3000)
    conf_a = rng.integers(0, 3, size=3000)
    conf_b = rng.integers(0, 2, size=3000)
    order = np.lexsort((conf_b, conf_a))
    conf_a, conf_b = conf_a[order], conf_b[order]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FC_CLASH_MODE", mode)
        res[mode] = compenetration_check_batch(ca, cb, xf, thresh=1.4, conf_a=conf_a, conf_b=conf_b)
    assert np.array_equal(res["0"].status & 1, res["1"].status & 1)
    ref_mask, _, closest = port.clash_batch(ca, cb, xf, thresh=1.4, conf_a=conf_a, conf_b=conf_b)
    assert np.array_equal(res["1"].mask[closest > 1e-6], ref_mask[closest > 1e-6])


def test_cell_list_path_edge_geometries(gpu, monkeypatch):
    """Cell-list path corner cases: many conformers of A (32^3 grids), a fragment B larger than one flag
    block, a fragment A too large for byte indices (falls back to all pairs), very small / large thresholds."""
    from firecode_b200 import synthetic

    rng = np.random.default_rng(123)
    monkeypatch.setenv("FC_CLASH_MODE", "1")
    # 130 conformers of A -> the 64^3 grids would exceed the budget -> 32^3
    _, ca, _, _ = synthetic.conformer_ensemble(rng, 130, 20, n_torsions=3)
    _, b, _, _ = synthetic.molecule_cloud(rng, 70)
    xf = synthetic.sweep_poses(rng, ca[0], b, 5200)
    conf_a = np.sort(rng.integers(0, 130, size=5200))
    res = compenetration_check_batch(ca, b, xf, thresh=1.5, conf_a=conf_a)
    ref_mask, _, closest = port.clash_batch(ca, b, xf, thresh=1.5, conf_a=conf_a)
    assert np.array_equal(res.mask[closest > 1e-6], ref_mask[closest > 1e-6])
    # B with 300 atoms (ten flag blocks), A with 300 atoms (all-pairs fallback: indices do not fit a byte)
    _, a300, _, _ = synthetic.molecule_cloud(rng, 300)
    _, b300, _, _ = synthetic.molecule_cloud(rng, 300)
    _, a40, _, _ = synthetic.molecule_cloud(rng, 40)
    for fa, fb in ((a40, b300), (a300, a40)):
        xf = synthetic.sweep_poses(rng, fa, fb, 3000)
        res = compenetration_check_batch(fa, fb, xf, thresh=1.5)
        ref_mask, _, closest = port.clash_batch(fa, fb, xf, thresh=1.5, chunk=256)
        assert np.array_equal(res.mask[closest > 1e-6], ref_mask[closest > 1e-6])
    # thresholds far from the usual 1.5 A
    xf = synthetic.sweep_poses(rng, a40, b, 4000)
    for thr in (0.05, 0.7, 4.0, 9.0):
        res = compenetration_check_batch(a40, b, xf, thresh=thr)
        ref_mask, _, closest = port.clash_batch(a40, b, xf, thresh=thr)
        assert np.array_equal(res.mask[closest > 1e-6], ref_mask[closest > 1e-6]), thr
    # poses far away from A (everything outside the grid) and exactly on top of it
    far = xf.copy()
    far[:, 9:] += 500.0
    assert compenetration_check_batch(a40, b, far, thresh=1.5).mask.all()
    on_top = np.tile(np.array([[1.0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]]), (64, 1))
    assert not compenetration_check_batch(a40, a40, on_top, thresh=1.5).mask.any()


def test_structure_form_batch(gpu):
    """utils.compenetration_check_structures == the reference loop of compenetration_refining
    (embedder.py:1954-1975) over complete structures, two and three fragments, with max_clashes."""
    from firecode_b200.utils import compenetration_check_structures

    rng = np.random.default_rng(17)
    structures = rng.normal(scale=2.6, size=(400, 33, 3))
    for ids in ((15, 18), (10, 12, 11)):
        for mc in (0, 1, 4):
            mask, closest = compenetration_check_structures(structures, ids, thresh=1.3, max_clashes=mc, return_closest=True)
            ref = np.array([port.compenetration_check(s, ids=ids, thresh=1.3, max_clashes=mc) for s in structures])
            safe = closest > 1e-6
            assert np.array_equal(mask[safe], ref[safe])
            assert 0 < mask.sum() < len(mask) or mc == 4
    assert compenetration_check_structures(structures[:0], (15, 18)).shape == (0,)


# ------------------------------------------------------------------------------------------------
# round 2: the default (cell-list) screen against adversarial inputs, compact poses, bitmask output
# ------------------------------------------------------------------------------------------------
def _near_threshold_poses(a, b, base, thr, deltas, count):
    """Translate poses of `base` along their closest pair so that the minimum distance becomes thr + delta."""
    _, dmin, _ = port.clash_batch(a, b, base, thresh=thr)
    out = []
    k = 0
    for p in np.flatnonzero(dmin > thr + 0.2)[:count]:
        placed = port.place(b, base[p])
        d = np.linalg.norm(placed[:, None, :] - a[None, :, :], axis=-1)
        j, i = np.unravel_index(np.argmin(d), d.shape)
        u = (a[i] - placed[j]) / d[j, i]
        xf = base[p].copy()
        xf[9:12] += u * (d[j, i] - (thr + deltas[k % len(deltas)]))
        k += 1
        out.append(xf)
    return np.array(out)


def test_adversarial_near_threshold_cell_list(gpu, monkeypatch):
    """the adversarial set (thresh +- {0, 1e-7, 1e-6, 1e-5, 3e-4}) at >= 20 k poses through the
    DEFAULT cell-list kernel (FC_CLASH_MODE=1), bit-exact against the oracle and against the all-pairs kernel."""
    rng = np.random.default_rng(20262)
    a, b = _frags(rng, 150, 150)
    thr = 1.5
    base = synthetic.sweep_poses(rng, a, b, 70000, shell=(2.0, 6.0), dtype=np.float64)
    deltas = [0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 3e-4, -3e-4]
    xf = _near_threshold_poses(a, b, base, thr, deltas, 20700)
    assert len(xf) >= 20000
    monkeypatch.setenv("FC_CLASH_MODE", "1")
    res = compenetration_check_batch(a, b, xf, thresh=thr, near_cap=len(xf))
    monkeypatch.setenv("FC_CLASH_MODE", "0")
    res_ap = compenetration_check_batch(a, b, xf, thresh=thr)
    mask, dmin2, _ = port.clash_batch(a, b, xf, thresh=thr)
    gap = np.abs(dmin2 - thr)
    far = gap > 1e-9  # beyond double rounding noise the decision must agree
    assert np.array_equal(res.mask[far], mask[far])
    assert np.array_equal(res.mask[far], res_ap.mask[far])
    near = gap <= 0.9 * NEAR_EPS
    assert near.sum() > 5000
    assert set(np.flatnonzero(near)) <= set(res.near_idx.tolist())
    assert np.all(res.status[np.flatnonzero(near)] & STATUS_NEAR)
    # every pose of this set sits inside the FP32 band: all of them were decided in FP64
    assert res.n_rechecked >= int(0.99 * len(xf))


def test_cell_faces_and_crowded_cells(gpu, monkeypatch):
    """Atoms of B placed within 1e-4 A of a cell FACE of the grid (any axis, either side) at a distance thresh + delta
    from an atom of A; and a fragment A with a dense blob (cells with more than 15 candidates: count byte 255)."""
    from firecode_b200.clash import cell_grid_meta

    rng = np.random.default_rng(77)
    a, b = _frags(rng, 150, 150)
    thr = 1.5
    meta = cell_grid_meta(a, thr)
    assert meta["g"] == 64 and 0.1 < meta["h"] < 1.0
    o, h = meta["origin"], meta["h"]
    base = synthetic.sweep_poses(rng, a, b, 30000, shell=(1.0, 5.0), dtype=np.float64)
    deltas = [0.0, 1e-6, -1e-6, 1e-4, -1e-4, 3e-3, -3e-3, 2e-2, -2e-2]
    centre = a.mean(axis=0)
    out = []
    for p in range(len(base)):
        placed = port.place(b, base[p])
        d = np.linalg.norm(placed[:, None, :] - a[None, :, :], axis=-1)
        j, i = np.unravel_index(np.argmin(d), d.shape)
        dist = thr + deltas[p % len(deltas)]
        # target position of atom j: on the sphere |x - a_i| = dist, pointing away from A, one coordinate on a face
        u = a[i] - centre + rng.normal(scale=0.3, size=3)
        u /= np.linalg.norm(u)
        axis = p % 3
        want = a[i][axis] + dist * u[axis]
        face = o[axis] + (np.round((want - o[axis]) / h - 0.5) + 0.5) * h + rng.uniform(-1e-4, 1e-4)
        c = (face - a[i][axis]) / dist
        if abs(c) >= 0.999:
            continue
        rest = np.delete(u, axis)
        rest *= np.sqrt(1 - c * c) / np.linalg.norm(rest)
        u = np.insert(rest, axis, c)
        target = a[i] + dist * u
        xf = base[p].copy()
        xf[9:12] += target - placed[j]
        out.append(xf)
    xf = np.array(out)
    assert len(xf) > 25000
    monkeypatch.setenv("FC_CLASH_MODE", "1")
    res = compenetration_check_batch(a, b, xf, thresh=thr)
    mask, dmin, _ = port.clash_batch(a, b, xf, thresh=thr)
    safe = np.abs(dmin - thr) > 1e-9
    assert np.array_equal(res.mask[safe], mask[safe])
    assert (np.abs(dmin - thr) < 4e-3).sum() > 500  # the constructed pair decides a good share of the poses
    assert 0.02 < mask.mean() < 0.98

    # ---- crowded cells: 40 atoms of A inside a 0.6 A ball
    a2 = a.copy()
    a2[100:140] = a[20] + rng.normal(scale=0.25, size=(40, 3))
    xf2 = synthetic.sweep_poses(rng, a2, b, 24000, shell=(-3.0, 3.0))
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FC_CLASH_MODE", mode)
        out[mode] = compenetration_check_batch(a2, b, xf2, thresh=thr)
    assert np.array_equal(out["0"].status & 1, out["1"].status & 1)
    mask2, dmin2, _ = port.clash_batch(a2, b, xf2[:8000], thresh=thr)
    safe = np.abs(dmin2 - thr) > 1e-6
    assert np.array_equal(out["1"].mask[:8000][safe], mask2[safe])
    # poses that drive atoms of B straight into the blob
    target = a2[100:140].mean(axis=0)
    xf3 = xf2[:20000].copy()
    for p in range(len(xf3)):
        j = p % len(b)
        placed_j = xf3[p, :9].reshape(3, 3) @ b[j] + xf3[p, 9:]
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        xf3[p, 9:] += target + u * rng.uniform(0.8, 2.6) - placed_j
    for mode in ("0", "1"):
        monkeypatch.setenv("FC_CLASH_MODE", mode)
        out[mode] = compenetration_check_batch(a2, b, xf3, thresh=thr)
    assert np.array_equal(out["0"].status & 1, out["1"].status & 1)
    mask3, dmin3, _ = port.clash_batch(a2, b, xf3[:6000], thresh=thr)
    safe = np.abs(dmin3 - thr) > 1e-6
    assert np.array_equal(out["1"].mask[:6000][safe], mask3[safe])


def _random_pose7(rng, a, b, n, shell=(-2.0, 4.0)):
    q = rng.normal(size=(n, 4)) * rng.uniform(0.5, 2.0, size=(n, 1))  # not unit: the expansion normalises
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    base = synthetic.radius_of_gyration(a) + synthetic.radius_of_gyration(b)
    t = d * rng.uniform(base + shell[0], base + shell[1], size=(n, 1))
    from firecode_b200.clash import pack_poses7

    return pack_poses7(q, t)


@pytest.mark.parametrize("n_a,n_b,n_poses,mode", [(150, 150, 40000, "1"), (150, 150, 3000, "0"), (40, 70, 20000, None),
                                                  (254, 33, 17000, "1"), (300, 20, 5000, None), (20, 300, 18000, "1")])
def test_compact_poses_match_oracle(gpu, monkeypatch, n_a, n_b, n_poses, mode):
    """FC_POSE_Q7 (28 bytes per pose in, one bit per pose out): mask == oracle on the documented FP64 expansion,
    for both kernels behind the entry point; bitmask == status bytes; counts consistent."""
    from firecode_b200.clash import compenetration_check_batch_pose7, unpack_bits

    if mode is None:
        monkeypatch.delenv("FC_CLASH_MODE", raising=False)
    else:
        monkeypatch.setenv("FC_CLASH_MODE", mode)
    rng = np.random.default_rng(n_a * 7 + n_b)
    a, b = _frags(rng, n_a, n_b)
    pose7 = _random_pose7(rng, a, b, n_poses)
    res = compenetration_check_batch_pose7(a, b, pose7, thresh=1.5, want_status=True)
    xf = port.pose7_to_xf(pose7)
    # the expansion is a rotation
    r = xf[:, :9].reshape(-1, 3, 3)
    assert np.abs(r @ r.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12
    n_ref = min(n_poses, 8000)
    mask, dmin, _ = port.clash_batch(a, b, xf[:n_ref], thresh=1.5, chunk=256)
    safe = np.abs(dmin - 1.5) > 1e-6
    assert np.array_equal(res.mask[:n_ref][safe], mask[safe])
    assert np.array_equal(unpack_bits(res.bits, n_poses), (res.status & 1).astype(bool))
    assert res.n_pass == int(res.mask.sum())
    assert res.n_rechecked == int(((res.status & 2) != 0).sum())
    # the f64 entry point on the expanded transforms gives the same status bytes
    res64 = compenetration_check_batch(a, b, xf, thresh=1.5)
    same = ((res.status | res64.status) & 2) == 0  # both decided in FP32: identical; FP64 rechecks may differ in membership
    assert np.array_equal(res.status[same], res64.status[same])
    assert np.array_equal(res.mask[safe_all(dmin, n_ref, n_poses)], res64.mask[safe_all(dmin, n_ref, n_poses)])
    assert 0.02 < res.mask.mean() < 0.98


def safe_all(dmin, n_ref, n_poses):
    ok = np.ones(n_poses, dtype=bool)
    ok[:n_ref] = np.abs(dmin - 1.5) > 1e-6
    return ok


def test_compact_poses_near_threshold_and_ragged(gpu, monkeypatch):
    """Compact poses whose minimum distance lands within float32 resolution of the threshold (translations are
    float32: the constructed offsets are rounded, so the distances scatter within ~1e-6 A of it), count mode, `<=`,
    conformer tiles, sizes that are not multiples of 32 / of the chunk, empty input."""
    from firecode_b200.clash import compenetration_check_batch_pose7, pack_poses7

    rng = np.random.default_rng(99)
    a, b = _frags(rng, 150, 150)
    thr = 1.5
    pose7 = _random_pose7(rng, a, b, 60000, shell=(2.0, 6.0))
    xf = port.pose7_to_xf(pose7)
    _, dmin, _ = port.clash_batch(a, b, xf, thresh=thr)
    idx = np.flatnonzero(dmin > thr + 0.2)[:20001]
    adj = pose7[idx].copy()
    for k, p in enumerate(idx):
        placed = port.place(b, xf[p])
        d = np.linalg.norm(placed[:, None, :] - a[None, :, :], axis=-1)
        j, i = np.unravel_index(np.argmin(d), d.shape)
        u = (a[i] - placed[j]) / d[j, i]
        adj[k, 4:] = (xf[p, 9:12] + u * (d[j, i] - thr)).astype(np.float32)
    monkeypatch.setenv("FC_CLASH_MODE", "1")
    monkeypatch.setenv("FC_CLASH_CHUNK", "7000")  # several ragged chunks
    res = compenetration_check_batch_pose7(a, b, adj, thresh=thr, near_cap=30000)
    monkeypatch.delenv("FC_CLASH_CHUNK")
    mask, dmin2, _ = port.clash_batch(a, b, port.pose7_to_xf(adj), thresh=thr)
    gap = np.abs(dmin2 - thr)
    assert (gap < 1e-5).sum() > 15000
    far = gap > 1e-9
    assert np.array_equal(res.mask[far], mask[far])
    listed = set(res.near_idx.tolist())
    assert set(np.flatnonzero(gap <= 0.9 * NEAR_EPS)) <= listed
    assert res.n_near == len(res.near_idx)
    # count mode and the non-strict comparison
    p2 = _random_pose7(rng, a, b, 20000, shell=(-3.0, 2.0))
    for mc, strict in ((3, True), (0, False), (12, False)):
        r2 = compenetration_check_batch_pose7(a, b, p2, thresh=thr, max_clashes=mc, strict=strict)
        m2, _, closest = port.clash_batch(a, b, port.pose7_to_xf(p2)[:5000], thresh=thr, max_clashes=mc, strict=strict)
        ok = closest > 1e-6
        assert np.array_equal(r2.mask[:5000][ok], m2[ok])
    # conformer tiles
    _, ens_a, _, _ = synthetic.conformer_ensemble(rng, 4, 40)
    _, ens_b, _, _ = synthetic.conformer_ensemble(rng, 3, 33)
    n = 17011
    ca = np.sort(rng.integers(0, 4, size=n))
    cb = rng.integers(0, 3, size=n)
    order = np.lexsort((cb, ca))
    ca, cb = ca[order], cb[order]
    p3 = _random_pose7(rng, ens_a[0], ens_b[0], n)
    r3 = compenetration_check_batch_pose7(ens_a, ens_b, p3, thresh=1.4, conf_a=ca, conf_b=cb)
    m3, d3, _ = port.clash_batch(ens_a, ens_b, port.pose7_to_xf(p3), thresh=1.4, conf_a=ca, conf_b=cb)
    ok = np.abs(d3 - 1.4) > 1e-6
    assert np.array_equal(r3.mask[ok], m3[ok])
    # empty and tiny
    r0 = compenetration_check_batch_pose7(a, b, np.zeros((0, 7), dtype=np.float32), thresh=thr)
    assert r0.mask.shape == (0,) and r0.n_pass == 0
    one = pack_poses7([[0, 0, 0, 1.0]], [[100.0, 0, 0]])
    assert compenetration_check_batch_pose7(a, b, one, thresh=thr).mask.tolist() == [True]
    one[0, 4] = 0.0
    assert compenetration_check_batch_pose7(a, a, one, thresh=thr).mask.tolist() == [False]


def test_level_boundaries_do_not_change_the_result(gpu, monkeypatch):
    """The cell-list screen runs in levels over the (re-ordered) atoms of B; any set of boundaries gives the
    same status bytes as the single-level run and as the all-pairs kernel."""
    rng = np.random.default_rng(5150)
    a, b = _frags(rng, 150, 150)
    xf = synthetic.sweep_poses(rng, a, b, 50000)
    monkeypatch.setenv("FC_CLASH_MODE", "0")
    ref = compenetration_check_batch(a, b, xf, thresh=1.5)
    monkeypatch.setenv("FC_CLASH_MODE", "1")
    for levels in ("", "149", "1", "8,16,40,100", "32,64,96,128", "5,6,7"):
        monkeypatch.setenv("FC_CLASH_LEVELS", levels)
        for mc in (0, 2):
            if mc and levels not in ("", "8,16,40,100"):
                continue
            out = compenetration_check_batch(a, b, xf, thresh=1.5, max_clashes=mc)
            want = ref if mc == 0 else None
            if want is None:
                monkeypatch.setenv("FC_CLASH_MODE", "0")
                want = compenetration_check_batch(a, b, xf, thresh=1.5, max_clashes=mc)
                monkeypatch.setenv("FC_CLASH_MODE", "1")
            assert np.array_equal(out.status & 1, want.status & 1), levels


def test_device_bitmask_epilogue(gpu):
    """fc_clash_screen_ex_dev writes the survivor bitmask itself (ballot / atomicOr): equal to the packed status
    bytes, for both pose formats, including a pose count that is not a multiple of 32."""
    import torch

    from firecode_b200 import clash

    rng = np.random.default_rng(31)
    a, b = _frags(rng, 150, 150)
    n = 50003
    pose7 = _random_pose7(rng, a, b, n)
    xf = port.pose7_to_xf(pose7)
    dev = torch.device("cuda", 0)
    a_dev = torch.from_numpy(a).to(dev)[None].contiguous()
    b_dev = torch.from_numpy(b).to(dev)[None].contiguous()
    prep = clash.DevicePrep(a_dev, 1.5)
    masks = []
    for fmt, poses in ((clash.POSE_Q7, torch.from_numpy(pose7).to(dev)), (clash.POSE_XF64, torch.from_numpy(xf).to(dev))):
        status = torch.zeros(n, dtype=torch.uint8, device=dev)
        bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        clash.screen_device_ex(prep, b_dev, poses, fmt, status_out=status, bits_out=bits, recheck_count=cnt)
        only_bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device=dev)
        clash.screen_device_ex(prep, b_dev, poses, fmt, bits_out=only_bits)
        torch.cuda.synchronize()
        st = status.cpu().numpy()
        m = clash.unpack_bits(bits.cpu().numpy().view(np.uint32), n)
        assert np.array_equal(m, (st & 1).astype(bool))
        assert np.array_equal(only_bits.cpu().numpy(), bits.cpu().numpy())
        assert int(cnt.item()) == int(((st & 2) != 0).sum())
        masks.append(m)
    prep.free()
    mask, dmin, _ = port.clash_batch(a, b, xf[:6000], thresh=1.5)
    safe = np.abs(dmin - 1.5) > 1e-6
    assert np.array_equal(masks[0][:6000][safe], mask[safe])
    assert np.array_equal(masks[1][:6000][safe], mask[safe])
