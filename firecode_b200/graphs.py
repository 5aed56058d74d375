"""Host-side graph logic feeding the torsion-fingerprint filter of the string embed.

Restates firecode/graph_manipulations.py:116-143 (get_sum_graph) and
firecode/torsion_module.py:385-408 (get_quadruplets).  Graph chemistry stays on the host
(SURVEY.md 2, row 4); only the per-pose dihedral arithmetic runs on the GPU.  networkx is used the
same way the reference uses it because spanning-tree edge order and neighbour order decide which
quadruplets are produced.
"""

from __future__ import annotations

import numpy as np
from networkx import Graph, minimum_spanning_tree


def sum_graph(graphs, extra_edges=()):
    """Union of the molecular graphs with cumulative atom numbering plus ``extra_edges``."""
    out = Graph()
    offset = 0
    for g in graphs:
        n = g.number_of_nodes()
        if offset == 0:
            out.add_nodes_from(g.nodes(data=True))
            out.add_edges_from(g.edges())
        else:
            for e1, e2 in g.edges():
                out.add_edge(e1 + offset, e2 + offset)
        offset += n
    for e1, e2 in extra_edges:
        out.add_edge(int(e1), int(e2))
    return out


def quadruplets(graph) -> np.ndarray:
    """(Q, 4) int: for every spanning-tree edge (i, j) with a further neighbour on both sides,
    (first neighbour of i, i, j, first neighbour of j)."""
    out = []
    for i, j in minimum_spanning_tree(graph).edges():
        ni = [n for n in graph.neighbors(i) if n not in (i, j)]
        nj = [n for n in graph.neighbors(j) if n not in (i, j)]
        if ni and nj:
            out.append((ni[0], i, j, nj[0]))
    return np.array(out, dtype=np.int64).reshape(-1, 4)
