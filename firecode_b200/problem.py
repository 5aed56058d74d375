"""Plain-array views of the data the embed functions read from an ``embedder`` (duck-typed
``firecode.embedder.Embedder``: SURVEY.md 8b lists the attributes).  Extraction is host-side and
O(conformers): everything per-pose happens on the GPU.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import graphs
from .utils import cartesian_product


@dataclass
class StringProblem:
    """Inputs of the string embed (embeds.py:51-158)."""

    coords: list            # [ (C1,N1,3), (C2,N2,3) ] float64, Hypermolecule.coords
    centers: list           # [ (C1,K1,3), (C2,K2,3) ] orbital centres of the first reactive atom
    vecs: list              # [ (C1,K1,3), (C2,K2,3) ] orbital vectors of the first reactive atom
    angles: np.ndarray      # (A,) degrees, embedder.systematic_angles
    quadruplets: np.ndarray  # (Q,4) int64, torsions of the sum graph (cumulative numbering)
    thresh: float
    constrained: np.ndarray  # (1,2) int: reactive atom pair (cumulative numbering)
    conf_pairs: np.ndarray = field(default=None)    # (C1*C2, 2) enumeration order (N1)
    center_pairs: np.ndarray = field(default=None)  # (K1*K2, 2)

    def __post_init__(self):
        c1, c2 = (len(c) for c in self.coords)
        k1, k2 = (c.shape[1] for c in self.centers)
        self.conf_pairs = cartesian_product(np.arange(c1), np.arange(c2)).astype(np.int64)
        self.center_pairs = cartesian_product(np.arange(k1), np.arange(k2)).astype(np.int64)

    @property
    def n_poses(self) -> int:
        return len(self.conf_pairs) * len(self.center_pairs) * len(self.angles)

    def decode(self, pose):
        """pose index -> (c1, c2, ai1, ai2, angle) in the reference's loop order."""
        n_ang, n_cen = len(self.angles), len(self.center_pairs)
        ci, rest = divmod(int(pose), n_cen * n_ang)
        ki, ai = divmod(rest, n_ang)
        c1, c2 = self.conf_pairs[ci]
        a1, a2 = self.center_pairs[ki]
        return int(c1), int(c2), int(a1), int(a2), float(self.angles[ai])


def string_problem(embedder) -> StringProblem:
    """Read a string-embed problem from an embedder (attributes read: embeds.py:86-156)."""
    mols = embedder.objects
    assert len(mols) == 2
    coords = [np.ascontiguousarray(np.asarray(m.coords, dtype=np.float64)) for m in mols]
    # number of orbital centres is taken from conformer 0 (embeds.py:94-96, quirk N3)
    n_centers = [len(m.get_centers(0)[0]) for m in mols]
    centers, vecs = [], []
    for m, k in zip(mols, n_centers):
        cen = np.empty((len(m.coords), k, 3))
        vec = np.empty((len(m.coords), k, 3))
        for c in range(len(m.coords)):
            ra = m.get_r_atoms(c)[0]
            cen[c] = np.asarray(ra.center, dtype=np.float64)[:k]
            vec[c] = np.asarray(ra.orb_vecs, dtype=np.float64)[:k]
        centers.append(cen)
        vecs.append(vec)
    constrained = np.array([[int(mols[0].reactive_indices[0]),
                             int(mols[1].reactive_indices[0] + embedder.ids[0])]], dtype=np.int64)
    quads = graphs.quadruplets(graphs.sum_graph((mols[0].graph, mols[1].graph), constrained))
    return StringProblem(coords=coords, centers=centers, vecs=vecs,
                         angles=np.asarray(embedder.systematic_angles, dtype=np.float64),
                         quadruplets=quads, thresh=float(embedder.options.clash_thresh),
                         constrained=constrained)


@dataclass
class CyclicalProblem:
    """Inputs of the bi-/tri-molecular cyclical embed (embeds.py:180-750)."""

    coords: list              # per molecule (C,N,3) float64
    reactive: list            # per molecule (n_reactive,) int
    pivot_vec: list           # per molecule list over conformers of (P,3) arrays: Pivot.pivot
    pivot_mean: list          # per molecule list over conformers of (P,3) arrays: Pivot.meanpoint
    pivot_ids: list           # per molecule list over conformers of (P,2) int: start/end cumnum
    angles: np.ndarray        # (A, M) degrees
    thresh: float
    pairings: list            # embedder.pairings_table.values() as list of tuples
    internal_constraints_is_array: bool
    internal_constraints: list
    max_norm_delta: float = 5.0
    ratoms0: list = None      # per molecule (K,2) int: [index, cumnum] of reactive_atoms_classes_dict[0]
                              # in dict order (read by _adjust_directions, embeds.py:330-337)
    ids: list = None          # embedder.ids (atoms per molecule)

    def __post_init__(self):
        if self.ids is None:
            self.ids = [int(c.shape[1]) for c in self.coords]
        if self.ratoms0 is None:
            off = np.concatenate([[0], np.cumsum(self.ids)[:-1]])
            self.ratoms0 = [np.array([[int(i), int(i) + int(off[m])] for i in self.reactive[m]],
                                     dtype=np.int64).reshape(-1, 2) for m in range(len(self.coords))]

    @property
    def n_mols(self) -> int:
        return len(self.coords)


def cyclical_problem(embedder, max_norm_delta: float = 5.0) -> CyclicalProblem:
    mols = embedder.objects
    coords = [np.ascontiguousarray(np.asarray(m.coords, dtype=np.float64)) for m in mols]
    pv, pm, pi = [], [], []
    for m in mols:
        pv.append([np.array([p.pivot for p in plist], dtype=np.float64).reshape(-1, 3) for plist in m.pivots])
        pm.append([np.array([p.meanpoint for p in plist], dtype=np.float64).reshape(-1, 3) for plist in m.pivots])
        pi.append([np.array([[p.start_atom.cumnum, p.end_atom.cumnum] for p in plist],
                            dtype=np.int64).reshape(-1, 2) for plist in m.pivots])
    table = getattr(embedder, "pairings_table", None) or {}
    ic = getattr(embedder, "internal_constraints", [])
    ratoms0 = None
    if all(hasattr(m, "reactive_atoms_classes_dict") for m in mols):
        ratoms0 = [np.array([[int(idx), int(ra.cumnum)] for idx, ra in m.reactive_atoms_classes_dict[0].items()],
                            dtype=np.int64).reshape(-1, 2) for m in mols]
    return CyclicalProblem(
        coords=coords, reactive=[np.asarray(m.reactive_indices, dtype=np.int64) for m in mols],
        pivot_vec=pv, pivot_mean=pm, pivot_ids=pi,
        angles=np.asarray(embedder.systematic_angles, dtype=np.float64).reshape(-1, len(mols)),
        thresh=float(embedder.options.clash_thresh),
        pairings=[tuple(int(x) for x in pair) for pair in table.values()],
        internal_constraints_is_array=isinstance(ic, np.ndarray),
        internal_constraints=np.asarray(ic).tolist() if len(ic) else [],
        max_norm_delta=float(max_norm_delta), ratoms0=ratoms0,
        ids=[int(x) for x in embedder.ids])
