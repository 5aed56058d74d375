"""Orbital centres of the reactive atoms for ALL conformers of a molecule at once: the producer in front of the
embedding screen.

The reference's ``Hypermolecule.compute_orbitals`` (hypermolecule_class.py:166-183) calls, conformer by conformer and
atom by atom, the ``init(mol, index, update=True, conf=c)`` of a reactive-atom class (reactive_atoms_classes.py) which
mixes two kinds of work: graph look-ups that do not depend on the conformer (neighbours, the bonded reactive partner,
which neighbour is the leaving group ...) and a short closed-form construction on that conformer's coordinates.  Here
the look-ups run once per atom and the constructions are evaluated for the whole (n_conf, ...) coordinate array:
``orbital_centers(kind, ...)`` returns ``center`` (n_conf, K, 3), the array ``problem.py`` stacks into the pivots /
orbital tables of the embed kernels.  Set-up scale (O(conformers x reactive atoms)): host numpy, no GPU work.

Covered: "Single Bond", "sp2", "sp3", "Ether", "Ketone" (ketene / carbonyl / trilobe sub-types), "Imine" -- with the
sigma-star (two bonded sp3 / single-bond centres) and sigmatropic variants.  "sp" / carbene centres draw a random
vector in the reference (np.random.rand, reactive_atoms_classes.py:579) and metals / lone atoms need the pairing
partners; those stay with the host application (NotImplementedError).

``orb_dim`` (half the transition-state bonding distance, firecode/parameters.py:23-48) is the caller's: pass the
number, or let ``default_orb_dim`` read the host application's table when firecode is importable.
"""

from __future__ import annotations

import numpy as np

from . import conventions

KINDS = ("Single Bond", "sp2", "sp3", "Ether", "Ketone", "Imine")
BOND_LENGTH = "bond length"   # orb_dim of a single-bond centre without parameters (reactive_atoms_classes.py:114-117)


def default_orb_dim(symbol, kind):
    """``orb_dim_dict`` look-up as the classes do it (key "<symbol> <kind>", "Fallback" otherwise); None for a single
    bond without parameters (the class then uses the bond length)."""
    try:
        from firecode.parameters import orb_dim_dict
    except Exception as exc:  # pragma: no cover - outside the host application the caller passes orb_dim
        raise ValueError("orb_dim must be given when firecode.parameters is not importable") from exc
    value = orb_dim_dict.get(f"{symbol} {kind}")
    if value is None and kind != "Single Bond":
        value = orb_dim_dict["Fallback"]
    return value


def _normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rot_mats(pointers, angle_deg):
    """rot_mat_from_pointer (prism_pruner.algebra contract, see firecode_b200.utils) for an array of axes (n, 3)."""
    p = _normalize(np.asarray(pointers, dtype=np.float64))
    half = conventions.ROT_HANDEDNESS * float(angle_deg) * np.pi / 360.0
    x, y, z = (p * np.sin(half)).T
    w = np.full(len(p), np.cos(half))
    return np.stack([
        np.stack([x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w], -1),
    ], -2)


def _apply(mats, vecs):
    """mats (n, 3, 3) @ vecs (n, 3)."""
    return np.einsum("nij,nj->ni", mats, vecs)


def _dot(a, b):
    return np.sum(a * b, axis=-1, keepdims=True)


def _neighbors(graph, i):
    return list(graph.neighbors(i))


def _partner(graph, i, reactive_indices):
    """The other reactive atom bonded to i (sigma-star pairs, reactive_atoms_classes.py:84-90, 240-246)."""
    nb = _neighbors(graph, i)
    for index in [r for r in reactive_indices if r != i]:
        if index in nb:
            return int(index)
    raise ValueError(f"atom {i} has no bonded reactive partner")


def _staggered(pivot, orb_vec, scale_unit):
    """Three directions staggered with respect to the substituents: orb_vec made orthogonal to pivot, rotated by
    60 / 180 / 300 degrees about it (reactive_atoms_classes.py:98-108, 252-263)."""
    orb_vec = orb_vec - _dot(orb_vec, pivot) * pivot
    vecs = np.stack([_apply(_rot_mats(pivot, angle + 60), orb_vec) for angle in range(0, 360, 120)], axis=1)
    return _normalize(vecs) if scale_unit else vecs


def orbital_centers(kind, coords, atoms, graph, index, orb_dim=None, reactive_indices=(), sp3_sigmastar=False,
                    sigmatropic=None):
    """``center`` of the reactive atom ``index`` for every conformer: (n_conf, K, 3).

    kind: the class name as ``repr`` gives it ("Ketone (sp2)" -> "Ketone"); coords (n_conf, n_atoms, 3); atoms: symbols;
    graph: the molecule's networkx graph; sigmatropic: bool per conformer (``mol.sigmatropic``); sp3_sigmastar:
    ``mol.sp3_sigmastar``.  Also returns the sub-type string where the class has one."""
    kind = str(kind).split(" (")[0]
    x = np.asarray(coords, dtype=np.float64)
    if x.ndim == 2:
        x = x[None]
    n_conf = len(x)
    atoms = [str(a) for a in atoms]
    i = int(index)
    nb = _neighbors(graph, i)
    coord = x[:, i]
    sigmatropic = np.zeros(n_conf, dtype=bool) if sigmatropic is None else np.asarray(sigmatropic, dtype=bool)
    if orb_dim is None:
        orb_dim = default_orb_dim(atoms[i], kind)
    if orb_dim is None or (isinstance(orb_dim, str) and orb_dim == BOND_LENGTH):
        if kind != "Single Bond":
            raise ValueError(f"{kind}: orb_dim must be a number")
        orb_dim = None

    if kind == "Single Bond":
        other = x[:, nb[0]]
        if not sp3_sigmastar:
            vecs = _normalize(coord - other)[:, None]
        else:
            partner = _partner(graph, i, reactive_indices)
            pivot = _normalize(x[:, partner] - coord)
            nop = [k for k in _neighbors(graph, partner) if k != i]
            vecs = _staggered(pivot, _normalize(x[:, nop[0]] - x[:, partner]), scale_unit=False)
        dim = np.linalg.norm(coord - other, axis=-1)[:, None, None] if orb_dim is None else orb_dim
        return dim * vecs + coord[:, None], None

    if kind == "sp2":
        v = _normalize(x[:, nb] - coord[:, None])
        normal = _normalize(np.mean(np.stack([np.cross(v[:, 0], v[:, 1]), np.cross(v[:, 1], v[:, 2]),
                                              np.cross(v[:, 2], v[:, 0])]), axis=0))
        return np.stack([normal, -normal], axis=1) * orb_dim + coord[:, None], None

    if kind == "sp3":
        if not sp3_sigmastar:
            symbols = [atoms[k] for k in nb]
            if len([s for s in symbols if s in ("O", "N", "Cl", "Br", "I")]) == 1:
                # (the reference tests five elements but looks the atom up among four: a lone N neighbour raises there)
                leaving = nb[symbols.index([s for s in symbols if s in ("O", "Cl", "Br", "I")][0])]
            elif len([s for s in symbols if s != "H"]) == 1:
                leaving = nb[symbols.index([s for s in symbols if s != "H"][0])]
            else:
                leaving = nb[0]
            vecs = _normalize(coord - x[:, leaving])[:, None]
        else:
            partner = _partner(graph, i, reactive_indices)
            pivot = _normalize(x[:, partner] - coord)
            others = [k for k in nb if k != partner]
            vecs = _staggered(pivot, _normalize(x[:, others[0]] - coord), scale_unit=True)
        return orb_dim * vecs + coord[:, None], None

    if kind == "Ether":
        v = orb_dim * _normalize(x[:, nb] - coord[:, None])
        mats = np.einsum("nij,njk->nik", _rot_mats(np.mean(v, axis=1), 90), _rot_mats(np.cross(v[:, 0], v[:, 1]), 180))
        return np.einsum("nij,nkj->nki", mats, v) + coord[:, None], None

    if kind == "Ketone":
        vector = _normalize(x[:, nb[0]] - coord) * orb_dim
        non = [k for k in _neighbors(graph, nb[0]) if k != i]
        if len(non) == 1:      # ketene
            subs = [k for k in _neighbors(graph, non[0]) if k != nb[0]]
            v = x[:, subs[0]] - x[:, non[0]]
            pointer = v - _dot(v, _normalize(vector)) * vector
            pointer = _normalize(pointer) * orb_dim
            center = np.stack([_apply(_rot_mats(vector, 90 * step), pointer) for step in range(4)], axis=1)
            subtype = "p+p"
        elif len(non) == 2:    # carbonyl / enolate: p lobes for sigmatropic conformers, n lobes otherwise
            pivot = _normalize(np.cross(x[:, non[0]] - coord, x[:, non[1]] - coord))
            p_lobes = np.stack([pivot * orb_dim, -pivot * orb_dim], axis=1)
            n_lobes = np.stack([_apply(_rot_mats(pivot, angle), vector) for angle in (120, 240)], axis=1)
            center = np.where(sigmatropic[:, None, None], p_lobes, n_lobes)
            subtype = "p" if sigmatropic.all() else ("sp2" if not sigmatropic.any() else "mixed")
        elif len(non) == 3:    # alkoxide, sulfonamide
            v = _normalize(x[:, non] - coord[:, None]) * orb_dim
            flip = _rot_mats(_normalize(np.cross(vector, v[:, 0])), 180)
            center = np.einsum("nij,nkj->nki", flip, v)
            subtype = "trilobe"
        else:
            raise ValueError("Ketone: the carbonyl carbon must carry one to three more neighbours")
        return center + coord[:, None], subtype

    if kind == "Imine":
        v = x[:, nb] - coord[:, None]
        p_lobe = _normalize(np.cross(v[:, 0], v[:, 1])) * orb_dim
        lone = -_normalize(np.mean(_normalize(v), axis=1)) * orb_dim
        if sigmatropic.all():
            return np.stack([p_lobe, -p_lobe], axis=1) + coord[:, None], None
        if not sigmatropic.any():
            return lone[:, None] + coord[:, None], None
        raise ValueError("Imine: conformers disagree on sigmatropicity (one lone-pair lobe against two p lobes): the "
                         "centre array would be ragged")

    raise NotImplementedError(f"orbital_centers: the '{kind}' centres stay with the host application (see module docstring)")


def pairing_orb_dims(embedder, mol_index):
    """{reactive atom index: orb_dim} imposed on molecule ``mol_index`` by the embedder's pairing distances
    (``DIST(a=2.5)`` -> half the distance on both partners, embedder.py:856-883)."""
    dims = {}
    for letter, dist in getattr(embedder, "pairing_dists", {}).items():
        r_index = embedder.pairings_dict[mol_index].get(letter)
        if r_index is None or isinstance(r_index, tuple):
            continue   # (tuples are internal constraints: no orbital spacing)
        for r_i in ([r_index] if isinstance(r_index, (int, np.integer)) else r_index):
            dims[int(r_i)] = dist / 2
    return dims


def compute_orbitals_batch(mol, write_back=False, orb_dims=None):
    """What ``Hypermolecule.compute_orbitals`` leaves in ``reactive_atom.center`` (hypermolecule_class.py:166-183), for all
    conformers at once: {reactive atom index: centres (n_conf, K, 3)}.  ``mol`` is the host application's molecule object
    after ``_inspect_reactive_atoms`` (it supplies ``coords``, ``atoms``, ``graph``, ``reactive_indices``,
    ``reactive_atoms_classes_dict`` for the class of every reactive atom, ``sp3_sigmastar`` and ``sigmatropic``).
    ``orb_dims`` = {atom index: orb_dim} overrides the parameter table (``pairing_orb_dims`` builds it from an embedder).
    ``write_back=True`` stores the rows into the per-conformer reactive-atom objects, as the reference's loop does."""
    coords = np.asarray(mol.coords, dtype=np.float64)
    orb_dims = orb_dims or {}
    out = {}
    for index, atom in mol.reactive_atoms_classes_dict[0].items():
        kind = repr(atom)
        symbol = str(mol.atoms[int(index)])
        base = kind.split(" (")[0]
        dim = orb_dims[int(index)] if int(index) in orb_dims else default_orb_dim(symbol, base)
        centers, _ = orbital_centers(kind, coords, mol.atoms, mol.graph, int(index),
                                     orb_dim=BOND_LENGTH if dim is None else dim,
                                     reactive_indices=[int(r) for r in mol.reactive_indices],
                                     sp3_sigmastar=bool(getattr(mol, "sp3_sigmastar", False)),
                                     sigmatropic=getattr(mol, "sigmatropic", None))
        out[int(index)] = centers
        if write_back:
            for c in range(len(coords)):
                mol.reactive_atoms_classes_dict[c][index].center = centers[c].copy()
    return out
