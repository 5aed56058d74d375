"""prism_pruner.graph_manipulations restated (TEST INFRASTRUCTURE).

Contracts inferred from firecode/hypermolecule_class.py:157, firecode/utils.py:345-346,578-588,
firecode/torsion_module.py:116-135,183-187,220, firecode/graph_manipulations.py:108-113.
"""

from __future__ import annotations

import numpy as np
from networkx import Graph, cycle_basis, set_node_attributes

from .periodic_table import RADII_TABLE


def d_min_bond(e1, e2, factor=1.2):
    """Largest distance at which e1-e2 count as bonded: factor * (r_cov[e1] + r_cov[e2])."""
    return factor * (RADII_TABLE[str(e1)] + RADII_TABLE[str(e2)])


def graphize(atoms, coords, mask=None):
    """Molecular graph: node attr "atoms" = symbol, edge iff distance < d_min_bond. No self loops."""
    atoms = np.asarray(atoms)
    coords = np.asarray(coords, dtype=float)
    n = len(atoms)
    mask = np.ones(n, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    graph = Graph()
    graph.add_nodes_from(range(n))
    set_node_attributes(graph, {i: str(a) for i, a in enumerate(atoms)}, "atoms")
    diff = coords[:, None, :] - coords[None, :, :]
    dist = np.sqrt((diff * diff).sum(axis=-1))
    for i in range(n):
        if not mask[i]:
            continue
        for j in range(i + 1, n):
            if mask[j] and dist[i, j] < d_min_bond(atoms[i], atoms[j]):
                graph.add_edge(i, j)
    return graph


def _neighbors(graph, index):
    return [n for n in graph.neighbors(index) if n != index]


def get_sp_n(index, graph):
    """Hybridisation exponent n of sp^n for C/N/O-like atoms from the neighbour count, else None."""
    element = graph.nodes[index]["atoms"]
    nb = len(_neighbors(graph, index))
    table = {
        "C": {4: 3, 3: 2, 2: 1},
        "N": {3: 3, 2: 2, 1: 1, 4: 3},
        "O": {2: 3, 1: 2},
        "S": {2: 3, 1: 2},
        "P": {3: 3, 4: 3},
        "B": {3: 2, 4: 3},
        "Si": {4: 3},
    }
    return table.get(element, {}).get(nb)


def is_amide_n(index, graph, mode=-1):
    """True for a nitrogen bonded to a carbonyl carbon. mode 0/1/2 selects primary/secondary/
    tertiary amides only; -1 accepts any."""
    if graph.nodes[index]["atoms"] != "N":
        return False
    nb = _neighbors(graph, index)
    n_h = sum(1 for n in nb if graph.nodes[n]["atoms"] == "H")
    if mode != -1 and (2 - n_h) != mode:
        return False
    for n in nb:
        if graph.nodes[n]["atoms"] == "C":
            for m in _neighbors(graph, n):
                if graph.nodes[m]["atoms"] == "O" and len(_neighbors(graph, m)) == 1:
                    return True
    return False


def is_ester_o(index, graph):
    """True for the bridging oxygen of an ester/acid (O bonded to a carbonyl carbon)."""
    if graph.nodes[index]["atoms"] != "O":
        return False
    nb = _neighbors(graph, index)
    if len(nb) != 2:
        return False
    for n in nb:
        if graph.nodes[n]["atoms"] == "C":
            for m in _neighbors(graph, n):
                if m != index and graph.nodes[m]["atoms"] == "O" and len(_neighbors(graph, m)) == 1:
                    return True
    return False


def get_phenyl_ids(index, graph):
    """Indices of a six-membered all-carbon ring containing ``index`` (else None)."""
    for cycle in cycle_basis(graph):
        if len(cycle) == 6 and index in cycle:
            if all(graph.nodes[i]["atoms"] == "C" for i in cycle):
                return list(cycle)
    return None
