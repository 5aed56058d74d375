"""Duck-typed synthetic ``embedder`` objects (no reference code involved) with the attributes the
embed functions read.  Used to feed identical inputs to the CUDA path, to the
oracle port and -- in the CPU container -- to the UNMODIFIED reference embed functions."""

from __future__ import annotations

from types import SimpleNamespace

import numpy as np
from networkx import Graph, set_node_attributes

from firecode_b200 import synthetic
from firecode_b200.utils import cartesian_product


class _RAtom:
    def __init__(self, center, orb_vecs, cumnum, index):
        self.center = center
        self.orb_vecs = orb_vecs
        self.cumnum = cumnum
        self.index = index


class _Pivot:
    def __init__(self, start, end, a1, a2, i1, i2):
        self.start, self.end = start, end
        self.start_atom, self.end_atom = a1, a2
        self.index1, self.index2 = i1, i2
        self.pivot = start - end
        self.meanpoint = np.mean((start, end), axis=0)


class _Mol:
    def __init__(self, atoms, coords, bonds, reactive_indices, orb_len, n_orb, rng, offset):
        self.atoms = atoms
        self.coords = coords
        self.reactive_indices = np.array(reactive_indices)
        self.rotation = np.identity(3)
        self.position = np.zeros(3)
        g = Graph()
        g.add_nodes_from(range(len(atoms)))
        g.add_edges_from(bonds)
        set_node_attributes(g, dict(enumerate(str(a) for a in atoms)), "atoms")
        self.graph = g
        self.reactive_atoms_classes_dict = {}
        for c in range(len(coords)):
            d = {}
            for idx in reactive_indices:
                nb = list(g.neighbors(idx))
                base = coords[c][idx] - coords[c][nb[0]]
                base /= np.linalg.norm(base)
                # lobes are tilted off the bond axis: an orbital collinear with the bond makes the
                # torsion across the new bond (neighbour, r1, r2, ...) degenerate, and the reference
                # then compares atan2(rounding noise) values -- unreproducible by construction
                perp = np.cross(base, np.array([0.3, -0.5, 0.8]))
                perp /= np.linalg.norm(perp)
                vecs = [(base + 0.6 * perp) / np.linalg.norm(base + 0.6 * perp)]
                if n_orb >= 2:
                    vecs = [(base + 0.8 * perp) / np.linalg.norm(base + 0.8 * perp),
                            (base - 0.8 * perp) / np.linalg.norm(base - 0.8 * perp)]
                vecs = np.array(vecs)
                d[int(idx)] = _RAtom(coords[c][idx] + orb_len * vecs, vecs, int(idx) + offset, int(idx))
            self.reactive_atoms_classes_dict[c] = d
        self.pivots = self._pivots()

    def get_r_atoms(self, c):
        return list(self.reactive_atoms_classes_dict[c].values())

    def get_centers(self, c):
        return np.array([[v for v in atom.center] for atom in self.get_r_atoms(c)])

    def _pivots(self):
        out = []
        for c in range(len(self.coords)):
            atoms = self.get_r_atoms(c)
            plist = []
            if len(atoms) == 2:
                a1, a2 = atoms
                for i, j in cartesian_product(range(len(a1.center)), range(len(a2.center))):
                    plist.append(_Pivot(a1.center[i], a2.center[j], a1, a2, int(i), int(j)))
            elif len(atoms) == 1 and len(atoms[0].center) >= 2:
                a1 = atoms[0]
                plist.append(_Pivot(a1.center[0], a1.center[1], a1, a1, 0, 1))
            out.append(plist)
        return out


def _pick_reactive(atoms, bonds, n, rng, coords=None):
    """Reactive atoms on the periphery of the molecule (far from the centroid), so that embeds have
    clash-free poses: the outermost heavy atom and, for n = 2, a heavy atom bonded to it."""
    heavy = [i for i, a in enumerate(atoms) if a != "H"]
    nbrs = {i: [] for i in range(len(atoms))}
    for a, b in bonds:
        nbrs[a].append(b)
        nbrs[b].append(a)
    dist = np.linalg.norm(coords - coords.mean(axis=0), axis=1)
    order = sorted(heavy, key=lambda i: -dist[i])
    for first in order:
        partners = [j for j in nbrs[first] if atoms[j] != "H"]
        if n == 1:
            return [int(first)]
        if partners:
            second = max(partners, key=lambda j: dist[j])
            return sorted([int(first), int(second)])
    raise RuntimeError("no reactive atoms found")


def make_embedder(embed, n_conf, n_atoms, seed, n_mols=2, n_orb=2, n_reactive=1, orb_len=1.25,
                  angles=None, thresh=1.5):
    """Synthetic embedder for a string (n_reactive=1) or cyclical (n_reactive=2) embed."""
    rng = np.random.default_rng(seed)
    n_conf = [n_conf] * n_mols if np.isscalar(n_conf) else list(n_conf)
    n_atoms = [n_atoms] * n_mols if np.isscalar(n_atoms) else list(n_atoms)
    mols, offset = [], 0
    for m in range(n_mols):
        atoms, coords, bonds, _ = synthetic.conformer_ensemble(rng, n_conf[m], n_atoms[m], n_torsions=4)
        reactive = _pick_reactive(atoms, bonds, n_reactive, rng, coords[0])
        mols.append(_Mol(atoms, coords, bonds, reactive, orb_len, n_orb, rng, offset))
        offset += n_atoms[m]
    if angles is None:
        if embed == "string":
            angles = [n * 360 / 36 for n in range(36)]
        else:
            steps, rng_deg = 5, 45
            angles = list(cartesian_product(*[range(steps + 1) for _ in mols]) * 2 * rng_deg / steps - rng_deg)
    logs = []
    emb = SimpleNamespace(
        objects=mols, ids=np.array(n_atoms), embed=embed, candidates=0,
        options=SimpleNamespace(clash_thresh=thresh, debug=False, max_clashes=0, rmsd=0.5),
        systematic_angles=angles, pairings_table={}, internal_constraints=np.array([]),
        log=lambda s="", p=True: logs.append(s), logs=logs)
    return emb
