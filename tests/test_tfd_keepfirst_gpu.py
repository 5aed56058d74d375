"""Keep-first torsion-fingerprint sweep (C-ABI fc_tfd_keepfirst; embeds.py:59-84 + torsion_module.py:1056-1067 in the
reference: a pose is accepted iff no EARLIER ACCEPTED pose is similar) against a sequential numpy restatement:
several blocks of the bit-matrix sweep, long chains of dependent decisions, empty and one-row inputs."""

import ctypes as C

import numpy as np
import pytest

from firecode_b200 import _lib
from firecode_b200.embeds import _ptr

pytestmark = pytest.mark.gpu


def keepfirst_numpy(fp, thr=10.0):
    """Sequential keep-first; also returns the smallest distance of any evaluated sum to the threshold."""
    n = len(fp)
    keep = np.zeros(n, dtype=bool)
    acc = np.zeros((0, fp.shape[1]))
    margin = np.inf
    for t in range(n):
        if len(acc):
            d = np.abs(acc - fp[t])
            d = np.abs(d - np.where(d > 180.0, 360.0, 0.0))
            s = d.sum(axis=1)
            margin = min(margin, np.abs(s - thr).min())
            if (s < thr).any():
                continue
        keep[t] = True
        acc = np.vstack([acc, fp[t][None]])
    return keep, margin


def keepfirst_gpu(fp, thr=10.0):
    lib = _lib.load(require_device=True)
    fp = np.ascontiguousarray(fp, dtype=np.float64)
    n, q = fp.shape
    labels = np.arange(n, dtype=np.int64)
    out = np.zeros(n, dtype=np.uint8)
    n_ties = C.c_int64(0)
    _lib.check(lib.fc_tfd_keepfirst(_ptr(fp), _ptr(labels), n, q, thr, _ptr(out), None, 0, C.byref(n_ties)), "fc_tfd_keepfirst")
    return out.astype(bool)


@pytest.mark.parametrize("n,q,centres", [(1, 3, 1), (2, 3, 1), (700, 4, 40), (20000, 5, 900), (40000, 3, 2500), (5000, 37, 200),
                                         (3000, 300, 60)])
def test_keepfirst_matches_sequential_sweep(gpu, n, q, centres):
    rng = np.random.default_rng(n + q)
    base = rng.uniform(-180.0, 180.0, size=(centres, q))
    # copies of a centre differ by about the threshold (10) in the wrapped L1 norm: both verdicts occur; q = 300 makes
    # the pair kernel stage fewer rows per CTA, q = 37 is an odd fingerprint length
    fp = base[rng.integers(0, centres, size=n)] + rng.normal(scale=8.0 / q, size=(n, q))
    fp = (fp + 180.0) % 360.0 - 180.0   # values on both sides of the +-180 wrap
    ref, margin = keepfirst_numpy(fp)
    assert margin > 1e-6            # no decision of this input sits on the threshold
    got = keepfirst_gpu(fp)
    assert np.array_equal(got, ref)
    assert 0 < ref.sum() < max(n, 2)


def test_keepfirst_long_dependency_chain(gpu):
    """Row t is similar to row t - 1 only: the verdict of every row depends on the one before it (alternating keep /
    drop), i.e. as many dependent rounds as rows; plus a second block of rows behind the chain."""
    n_chain = 960
    fp = np.zeros((n_chain + 18000, 2))
    fp[:n_chain, 0] = (np.arange(n_chain) * 6.0) % 360.0 - 180.0   # neighbours 6 apart (similar), next-nearest 12 (not)
    fp[:n_chain, 1] = np.arange(n_chain) // 60 * 11.0              # the circle is walked 16 times: 11 per lap keeps them apart
    rng = np.random.default_rng(3)
    fp[n_chain:] = np.column_stack([rng.uniform(-180, 180, 18000), rng.uniform(2000.0, 2300.0, 18000)])
    ref, margin = keepfirst_numpy(fp)
    assert margin > 1e-6
    assert ref[:8].tolist() == [True, False, True, False, True, False, True, False]
    assert np.array_equal(keepfirst_gpu(fp), ref)


def test_keepfirst_empty(gpu):
    lib = _lib.load(require_device=True)
    n_ties = C.c_int64(-1)
    _lib.check(lib.fc_tfd_keepfirst(None, None, 0, 4, 10.0, None, None, 0, C.byref(n_ties)), "fc_tfd_keepfirst")
    assert n_ties.value == 0
