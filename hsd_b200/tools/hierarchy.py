"""Drop-in for ``tools/hierarchy.py``: k-hop ring ("hierarchy") construction and
its ``.layers`` text format.

The BFS itself (tools/hierarchy.py:25-38) runs on the GPU (hsd_bfs_rings: CSR
frontier expansion with bitmap visited sets); this module converts to and from
the reference's Python-dict / text forms.  Ring order inside a layer is
ascending node index — the reference's is set-iteration order, i.e. unspecified.
"""
from __future__ import annotations

import os

import networkx as nx
import numpy as np

__all__ = ["get_hierarchical_representation", "get_node_hierarchical_structure",
           "save_hierarchical_representation", "read_hierarchical_representation", "read_hierarchy"]

# The reference hard-codes an absolute path of its author's machine (tools/const.py:12);
# here the directory comes from the environment, with the same file naming.
HierarchyDirEnv = "HSD_HIERARCHY_DIR"


def _ringset(graph: nx.Graph, maxHop: int, sources=None):
    import torch
    from .. import engine
    from ..graph import CSRGraph
    from ..rings import RingSet
    g = CSRGraph.from_networkx(graph)
    dg = engine.DeviceGraph.upload(g)
    if sources is None:
        return g, RingSet.bfs(dg, maxHop)
    idx = {v: i for i, v in enumerate(g.nodes)}
    rows = torch.tensor([idx[s] for s in sources], dtype=torch.int32, device=dg.rowptr.device)
    _, sizes, bm, _ = engine.ring_signature_degree(dg, maxHop, rows=rows, want_sig=False, want_bitmaps=True)
    return g, RingSet(bitmaps=bm, sizes=sizes, hops=maxHop, n=g.n, orig_of=dg.orig_of)


def get_hierarchical_representation(graph: nx.Graph, maxHop):
    """tools/hierarchy.py:16-22: {node: [[node], ring_1, ..., ring_maxHop]} for every node."""
    g, rs = _ringset(graph, maxHop)
    hierarchy = rs.to_hierarchy(g.nodes)
    print(f"done, number of nodes: {len(hierarchy)}")
    return hierarchy


def get_node_hierarchical_structure(graph: nx.Graph, node: str, maxHop: int):
    """tools/hierarchy.py:25-38 for one source; empty rings are kept as []."""
    g, rs = _ringset(graph, maxHop, [node])
    bm = rs.bitmaps.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(bm.view(np.uint8), axis=-1, bitorder="little")[0, :, :g.n]
    orig = rs.orig_of.cpu().numpy()
    return [[g.nodes[j] for j in np.sort(orig[np.nonzero(bits[h])[0]])] for h in range(maxHop + 1)]


def save_hierarchical_representation(graph: nx.Graph, file_path: str, hop=7):
    """tools/hierarchy.py:41-63: one line per node, ``node#a,b,...#c,...#``; stops at the
    first empty ring."""
    g, rs = _ringset(graph, hop)
    mem = rs.members_host()
    with open(file_path, encoding="utf-8", mode="w+") as fout:
        for i in range(g.n):
            record = ""
            for layer in mem[i]:
                if len(layer) == 0:
                    break
                record += ",".join(str(g.nodes[j]) for j in layer) + "#"
            fout.write(record + "\n")


def read_hierarchical_representation(graphName: str, maxHop=3) -> dict:
    """tools/hierarchy.py:66-73 with the directory taken from $HSD_HIERARCHY_DIR."""
    base = os.environ.get(HierarchyDirEnv, os.path.join(os.getcwd(), "data", "hierarchy"))
    return read_hierarchy(os.path.join(base, "{}.layers".format(graphName)), maxHop)


def read_hierarchy(file_path: str, maxHop: int) -> dict:
    """tools/hierarchy.py:76-99: exactly maxHop+1 layers of *strings* per node; the first
    missing layer comes back as [''] (split of the trailing '#'), later ones as []."""
    if not os.path.exists(file_path):
        raise FileNotFoundError(f"path:{file_path}, hierarchy file not exist")
    hierarchy = {}
    with open(file_path, mode="r", encoding="utf-8") as fin:
        for raw in fin:
            line = raw.strip()
            if not line:
                break
            parts = line.split("#")
            layers = [parts[h].strip().split(",") if h < len(parts) else [] for h in range(maxHop + 1)]
            hierarchy[layers[0][0]] = layers
    print(f"done, number of nodes: {len(hierarchy)}")
    return hierarchy
