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
