"""Steps either side of the embedding screen with the reference's names and semantics:

* ``scramble_check`` / ``molecule_check`` (firecode/utils.py:341-400), the bond-graph post-filters the optimisation
  loops call per structure (embedder.py:2181-2195, 2430-2444; optimization_methods.py:139-147; operators.py:502;
  interfaces/goat.py:309) -- here for a whole batch of structures in one GPU call (C-ABI ``fc_bond_delta_batch``: one
  CTA per structure walks the atom pairs and compares the bond bits with the expected ones);
* ``get_pivots`` / ``set_pivots`` (``Embedder._get_pivots`` / ``_set_pivots``, embedder.py:902-987), the pivot tables of
  the cyclical embeds, built for all conformers at once from the orbital centres (plain arrays in, plain arrays out:
  set-up scale, stays on the host).

The bond criterion is prism_pruner's ``graphize`` (atoms closer than ``d_min_bond`` = 1.2 x the sum of the covalent
radii); prism_pruner is not part of the reference tree, so the radii / factor are the ones restated in
``oracle/prism_pruner`` unless the host application's own table can be imported (PARITY UNPINNED at that boundary).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .pt import COVALENT_RADII
from .utils import cartesian_product

BOND_FACTOR = 1.2   # prism_pruner.graph_manipulations.d_min_bond: factor * (r_cov[e1] + r_cov[e2])
NEAR_EPS = 1e-6


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def covalent_radii(atoms):
    """Covalent radius of every atom: the host application's table (prism_pruner.periodic_table) when it is
    importable, the built-in one (firecode_b200.pt) otherwise."""
    table = COVALENT_RADII
    try:  # inside FIRECODE the installed prism_pruner is the authority
        from prism_pruner.periodic_table import RADII_TABLE as table  # type: ignore
    except Exception:
        pass
    return np.array([float(table[str(a)]) for a in np.asarray(atoms)], dtype=np.float64)


def n_words(n_atoms):
    return (int(n_atoms) + 31) // 32


def bits_from_edges(n_atoms, edges):
    """Symmetric bond-bit matrix (n_atoms, ceil(n_atoms / 32)) uint32 from an iterable of (a, b) pairs; self loops
    are dropped as the reference does (``if a != b``)."""
    bits = np.zeros((n_atoms, n_words(n_atoms)), dtype=np.uint32)
    for a, b in edges:
        a, b = int(a), int(b)
        if a == b:
            continue
        bits[a, b >> 5] |= np.uint32(1 << (b & 31))
        bits[b, a >> 5] |= np.uint32(1 << (a & 31))
    return bits


def edges_from_bits(bits):
    """Sorted (a, b), a < b, bonds of one bond-bit matrix."""
    n = bits.shape[0]
    dense = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8).reshape(n, -1), axis=1, bitorder="little")[:, :n]
    a, b = np.nonzero(np.triu(dense, 1))
    return [(int(i), int(j)) for i, j in zip(a, b)]


def assembly_bits(mols_graphs):
    """Expected bonds of a multimolecular assembly: the union of the fragments' graphs, each shifted by the number of
    atoms before it (utils.py:371-377)."""
    n_tot = sum(len(g.nodes) for g in mols_graphs)
    edges, pos = [], 0
    for g in mols_graphs:
        edges += [(a + pos, b + pos) for a, b in g.edges if a != b]
        pos += len(g.nodes)
    return bits_from_edges(n_tot, edges)


def bond_graph_batch(atoms, structures, factor=BOND_FACTOR, want_near=False):
    """Bond bits of every structure (P, n, W) uint32 on the GPU (C-ABI fc_bond_graph_batch)."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    if x.ndim == 2:
        x = x[None]
    p, n = x.shape[:2]
    radii = covalent_radii(atoms)
    assert len(radii) == n
    adj = np.zeros((p, n, n_words(n)), dtype=np.uint32)
    near = np.zeros(p, dtype=np.int32)
    if p:
        _lib.check(lib.fc_bond_graph_batch(_ptr(x), p, n, _ptr(radii), float(factor), _ptr(adj), _ptr(near)),
                   "fc_bond_graph_batch")
    return (adj, near) if want_near else adj


def bond_delta_batch(atoms, structures, expected_bits, excluded_atoms=(), factor=BOND_FACTOR):
    """Number of bonds present in exactly one of {structure, expected} per structure, bonds touching an excluded atom not
    counted.  ``expected_bits`` is (n, W) for all structures or (P, n, W) per structure.  Returns (delta, near)."""
    lib = _lib.load(require_device=True)
    x = np.ascontiguousarray(np.asarray(structures, dtype=np.float64))
    if x.ndim == 2:
        x = x[None]
    p, n = x.shape[:2]
    radii = covalent_radii(atoms)
    assert len(radii) == n, "atoms and structures disagree"
    exp = np.ascontiguousarray(expected_bits, dtype=np.uint32)
    per_structure = exp.ndim == 3
    assert exp.shape[-2:] == (n, n_words(n)) and (not per_structure or len(exp) == p)
    excl = np.zeros(n, dtype=np.uint8)
    for a in excluded_atoms:
        excl[int(a)] = 1
    delta = np.zeros(p, dtype=np.int32)
    near = np.zeros(p, dtype=np.int32)
    if p:
        _lib.check(lib.fc_bond_delta_batch(_ptr(x), p, n, _ptr(radii), float(factor), _ptr(exp), 1 if per_structure else 0,
                                           _ptr(excl), _ptr(delta), _ptr(near)), "fc_bond_delta_batch")
    return delta, near


def scramble_check_batch(embedded_atoms, structures, excluded_atoms, mols_graphs, max_newbonds=0, return_counts=False):
    """``scramble_check`` (utils.py:356-400) for a batch of structures of the same assembly: True where at most
    ``max_newbonds`` bonds formed or broke with respect to the fragments' graphs."""
    x = np.asarray(structures, dtype=np.float64)
    if x.ndim == 2:
        x = x[None]
    assert x.shape[1] == sum(len(g.nodes) for g in mols_graphs)
    delta, near = bond_delta_batch(embedded_atoms, x, assembly_bits(mols_graphs), excluded_atoms)
    ok = delta <= int(max_newbonds)
    return (ok, delta, near) if return_counts else ok


def scramble_check(embedded_atoms, embedded_structure, excluded_atoms, mols_graphs, max_newbonds=0, logfunction=None,
                   title=None):
    """Single-structure form with the reference's signature and log line (utils.py:356-400)."""
    excluded_atoms = list(excluded_atoms)
    ok, delta, _ = scramble_check_batch(embedded_atoms, embedded_structure, excluded_atoms, mols_graphs, max_newbonds,
                                        return_counts=True)
    if not ok[0] and logfunction is not None:
        bonds = set(edges_from_bits(assembly_bits(mols_graphs)))
        new_bonds = set(edges_from_bits(bond_graph_batch(embedded_atoms, embedded_structure)[0]))
        delta_bonds = {b for b in (bonds | new_bonds) - (bonds & new_bonds) if not any(a in b for a in excluded_atoms)}
        logfunction(f"{title}, scramble_check - found {len(delta_bonds)} extra bonds: {delta_bonds}")
    return bool(ok[0])


def molecule_check_batch(atoms, old_coords, new_structures, max_newbonds=0, return_counts=False):
    """``molecule_check`` (utils.py:341-353) for a batch: ``old_coords`` is one structure (n, 3) -- every new structure
    is compared with its graph -- or one per new structure (P, n, 3)."""
    new = np.asarray(new_structures, dtype=np.float64)
    if new.ndim == 2:
        new = new[None]
    old = np.asarray(old_coords, dtype=np.float64)
    expected = bond_graph_batch(atoms, old)
    expected = expected[0] if old.ndim == 2 else expected
    delta, near = bond_delta_batch(atoms, new, expected)
    ok = delta <= int(max_newbonds)
    return (ok, delta, near) if return_counts else ok


def molecule_check(atoms, old_coords, new_coords, max_newbonds=0):
    """Single-structure form with the reference's signature (utils.py:341-353)."""
    return bool(molecule_check_batch(atoms, old_coords, np.asarray(new_coords)[None], max_newbonds)[0])


# ------------------------------------------------------------------------------------------------
# pivots of the cyclical embeds
# ------------------------------------------------------------------------------------------------
def get_pivots(centers):
    """``Embedder._get_pivots`` (embedder.py:936-987) for all conformers at once.

    ``centers``: one array per reactive atom of the molecule, (n_conf, K_atom, 3) -- the orbital centres
    ``reactive_atom.center`` of every conformer.  Two reactive atoms: one pivot per pair of centres in
    ``cartesian_product`` order (start on the first atom's centre, end on the second's).  One reactive atom
    (chelotropic): one pivot per unordered pair of its own centres.  Returns a dict of arrays over conformers:
    start, end, pivot (= start - end), meanpoint (n_conf, P, 3) and index (P, 2); P = 0 for any other atom count."""
    centers = [np.asarray(c, dtype=np.float64) for c in centers]
    if len(centers) == 2:
        c1, c2 = centers
        idx = cartesian_product(range(c1.shape[1]), range(c2.shape[1])).astype(np.int64)
        start, end = c1[:, idx[:, 0]], c2[:, idx[:, 1]]
    elif len(centers) == 1:
        c1 = centers[0]
        k = c1.shape[1]
        idx = cartesian_product(range(k), range(k)).astype(np.int64)
        idx = idx[(idx[:, 0] != idx[:, 1]) & (idx[:, 0] <= idx[:, 1])].reshape(-1, 2)
        start, end = c1[:, idx[:, 0]], c1[:, idx[:, 1]]
    else:
        n_conf = len(centers[0]) if centers else 0
        z = np.zeros((n_conf, 0, 3))
        return {"start": z, "end": z.copy(), "pivot": z.copy(), "meanpoint": z.copy(), "index": np.zeros((0, 2), dtype=np.int64)}
    return {"start": start, "end": end, "pivot": start - end, "meanpoint": np.mean((start, end), axis=0), "index": idx}


def set_pivots(centers, suprafacial=False, sp3_sigmastar=False):
    """``Embedder._set_pivots`` (embedder.py:902-934): the pivots of ``get_pivots`` after the two filters, per conformer
    (the filters can keep a different subset in every conformer): a list of index arrays into the pivot table plus the
    table itself.  suprafacial: with four pivots only the two shortest stay (when the shortest two are unambiguous, as
    the reference's loop decides it); sp3_sigmastar: only the pivots within 1e-5 of the shortest stay."""
    table = get_pivots(centers)
    n_conf, n_piv = table["pivot"].shape[:2]
    keep = []
    for c in range(n_conf):
        sel = np.arange(n_piv)
        if suprafacial and len(sel) == 4:
            norms = np.linalg.norm(table["pivot"][c][sel], axis=1)
            for sample in norms:
                to_keep = norms[sample >= norms]
                if len(to_keep) == 2:
                    sel = sel[np.isin(norms, to_keep)]
                    break
        if sp3_sigmastar and len(sel):
            lengths = np.linalg.norm(table["pivot"][c][sel], axis=1)
            sel = sel[(lengths - lengths.min()) < 1e-5]
        keep.append(sel)
    return table, keep
