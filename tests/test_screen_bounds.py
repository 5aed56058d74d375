"""Mathematics of the tensor-core pruning screen (firecode_b200/csrc/fc_gram_tc.cuh), checked in numpy on the CPU:
every quantity the epilogue compares against is an UPPER bound on S = sum of the singular values of the 3x3
covariance H (sign of det H on the smallest), so `e0 - 2 S_upper` is a lower bound on the summed squared deviation
and a pair that is ruled out cannot be similar; and the TF32 rounding of the operands stays inside the band the
kernel subtracts (kGramTf32Eps).  The CUDA kernel itself is compared with the FP64 path on the GPU
(tests/test_prune_gpu.py::test_prune_screen_flavours_agree, tools/gram_tc_test.cu)."""

import numpy as np

TF32_EPS = 1.0 / 512.0          # kGramTf32Eps
SCREEN_BAND = 0.05              # kScreenBand (Angstrom)


def _signed_sigma_sum(h):
    s = np.linalg.svd(h, compute_uv=False)
    return s.sum(axis=-1) - 2.0 * s[..., 2] * (np.linalg.det(h) < 0)


def _cof_norm2_and_det(h):
    c = np.empty_like(h)
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]
            q = [k for k in range(3) if k != j]
            c[..., i, j] = (-1) ** (i + j) * (h[..., r[0], q[0]] * h[..., r[1], q[1]] - h[..., r[0], q[1]] * h[..., r[1], q[0]])
    return (c * c).sum(axis=(-1, -2)), np.linalg.det(h)


def _random_covariances(rng, n):
    """3x3 matrices with singular values over several orders of magnitude, random orientation and both signs of det."""
    u, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    v, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    sig = np.sort(10.0 ** rng.uniform(-2, 3, size=(n, 3)), axis=1)[:, ::-1]
    sig[rng.random(n) < 0.2, 2] = 0.0                        # rank-deficient cases
    eq = rng.random(n) < 0.1
    sig[eq] = sig[eq][:, :1]                                 # three equal singular values (the loosest case of pass 1)
    return np.einsum("nij,nj,nkj->nik", u, sig, v)


def test_every_epilogue_bound_is_an_upper_bound():
    rng = np.random.default_rng(0)
    h = _random_covariances(rng, 20000)
    s = _signed_sigma_sum(h)
    f2 = (h * h).sum(axis=(-1, -2))
    cc, det = _cof_norm2_and_det(h)
    tol = 1e-9 * np.sqrt(f2)
    assert np.all(np.sqrt(3.0 * f2) >= s - tol)                                  # pass 1: sqrt(3) |H|_F
    lam0 = np.sqrt(f2 + 2.0 * np.sqrt(3.0 * cc))                                 # pass 2: S^2 = f2 + 2 e2, e2 <= sqrt(3 cc)
    assert np.all(lam0 >= s - tol)
    assert np.all(lam0 <= np.sqrt(3.0 * f2) * (1 + 1e-12))                       # ... and it is never looser than pass 1
    # Newton from above on P(x) = (x^2 - f2)^2 - 8 det x - 4 cc: the iterates decrease monotonically onto S
    lam = lam0.copy()
    for _ in range(60):
        t = lam * lam - f2
        p = t * t - 8.0 * det * lam - 4.0 * cc
        dp = 4.0 * lam * t - 8.0 * det
        step = np.where((dp > 0) & (p > 0), p / np.where(dp > 0, dp, 1.0), 0.0)
        new = lam - step
        assert np.all(new >= s - 1e-7 * np.maximum(s, 1.0))
        assert np.all(new <= lam + 1e-12)
        lam = new
    well = f2 > 0
    assert np.max(np.abs(lam - s)[well] / np.sqrt(f2[well])) < 1e-6              # ... and converge to it


def test_square_root_free_form_of_pass_two():
    """u - 2 sqrt(f2 + 2 sqrt(3 cc)) > 0  <=>  u > 0 and v = u^2 - 4 f2 > 0 and v^2 > 192 cc."""
    rng = np.random.default_rng(1)
    f2 = 10.0 ** rng.uniform(-2, 6, 100000)
    cc = f2 * f2 / 3.0 * rng.random(100000)                                       # cc <= f2^2 / 3 always
    u = np.sqrt(f2) * 10.0 ** rng.uniform(-1, 1, 100000) * np.where(rng.random(100000) < 0.1, -1, 1)
    direct = u - 2.0 * np.sqrt(f2 + 2.0 * np.sqrt(3.0 * cc)) > 0
    v = u * u - 4.0 * f2
    free = (u > 0) & (v > 0) & (v * v > 192.0 * cc)
    margin = np.abs(u - 2.0 * np.sqrt(f2 + 2.0 * np.sqrt(3.0 * cc))) > 1e-9 * np.abs(u)
    assert np.array_equal(direct[margin], free[margin])


def _round_tf32(x):
    """Round-to-nearest (ties away) to 10 explicit mantissa bits, as cvt.rna.tf32.f32 does."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & ~np.uint64(0x1FFF)).astype(np.uint32)
    return b.view(np.float32)


def test_tf32_rounding_stays_inside_the_band_and_no_similar_pair_is_ruled_out():
    rng = np.random.default_rng(2)
    n, nh, max_rmsd = 600, 58, 0.5
    base = rng.normal(size=(30, nh, 3)) * 3.0
    x = base[rng.integers(0, 30, n)] + rng.normal(size=(n, nh, 3)) * rng.uniform(0.02, 0.4, size=(n, 1, 1))
    x -= x.mean(axis=1, keepdims=True)
    g = (x * x).sum(axis=(1, 2))
    xt = _round_tf32(x).astype(np.float64)
    i, j = np.triu_indices(n, 1)
    h = np.einsum("pka,pkb->pab", x[i], x[j])
    h_tc = np.einsum("pka,pkb->pab", xt[i], xt[j]).astype(np.float32).astype(np.float64)   # FP32 accumulators
    rel = np.sqrt(((h_tc - h) ** 2).sum(axis=(1, 2))) / np.sqrt(g[i] * g[j])
    assert rel.max() < TF32_EPS / 2                         # rigorous bound for nearest rounding: 2^-10
    # the screen's decision (all bounds collapse to the exact S here) never rules out a similar pair
    e0 = g[i] + g[j]
    true_msd = (e0 - 2.0 * _signed_sigma_sum(h)) / nh
    u = e0 * (1.0 - np.sqrt(3.0) * TF32_EPS) - (max_rmsd + SCREEN_BAND) ** 2 * nh
    ruled_out = u - 2.0 * _signed_sigma_sum(h_tc) > 0
    similar = true_msd < max_rmsd ** 2
    assert similar.sum() > 100 and ruled_out.sum() > 1000
    assert not np.any(ruled_out & similar)
    # and the band costs little: the pairs kept for the FP64 stage are within ~0.2 A of the threshold
    kept = ~ruled_out
    assert np.sqrt(np.maximum(true_msd[kept], 0)).max() < max_rmsd + 0.25


def test_tile_culling_bound_von_neumann():
    """Tile culling of the screen: for centred coordinate matrices P, Q the best superposition leaves
    E = |P|^2 + |Q|^2 - 2 S >= sum_k (sigma_k(P) - sigma_k(Q))^2 (sigma = singular values, descending), whatever the
    relative orientation -- also with the proper-rotation constraint (S carries the sign of det on the smallest singular
    value, which only lowers it) and for planar / linear species."""
    rng = np.random.default_rng(11)
    for n_atoms in (3, 5, 20, 72):
        p = rng.normal(size=(400, n_atoms, 3)) * rng.uniform(0.2, 3.0, size=(400, 1, 3))
        q = rng.normal(size=(400, n_atoms, 3)) * rng.uniform(0.2, 3.0, size=(400, 1, 3))
        q[:50] = p[:50] + rng.normal(size=(50, n_atoms, 3)) * 0.05              # near duplicates, rotated below
        rot, _ = np.linalg.qr(rng.normal(size=(400, 3, 3)))
        q = np.einsum("nij,nkj->nki", rot, q)                                     # random (im)proper orthogonal maps
        p[100:120, :, 2] = 0.0                                                    # planar structures
        p[120:130, :, 1:] = 0.0                                                   # linear structures
        p -= p.mean(axis=1, keepdims=True)
        q -= q.mean(axis=1, keepdims=True)
        h = np.einsum("nka,nkb->nab", p, q)
        e = (p * p).sum(axis=(1, 2)) + (q * q).sum(axis=(1, 2)) - 2.0 * _signed_sigma_sum(h)
        sp = np.linalg.svd(p, compute_uv=False)
        sq = np.linalg.svd(q, compute_uv=False)
        bound = ((sp - sq) ** 2).sum(axis=1)
        assert np.all(e >= bound - 1e-9 * (1.0 + bound)), (n_atoms, (bound - e).max())
        # the one-number version (norms only) is weaker but also a bound
        assert np.all(bound >= (np.linalg.norm(p, axis=(1, 2)) - np.linalg.norm(q, axis=(1, 2))) ** 2 - 1e-9)
