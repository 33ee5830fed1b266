"""The pieces of the reference's tools/util.py the hot path depends on."""
import networkx as nx
import numpy as np

__all__ = ["build_node_idx_map", "recommend_scale_range", "scale_boundary"]


def build_node_idx_map(graph) -> (dict, dict):
    """Node <-> index maps in order of first appearance (tools/util.py:11-24)."""
    nodes = list(nx.nodes(graph))
    idx2node = dict(enumerate(nodes))
    node2idx = {node: idx for idx, node in idx2node.items()}
    return idx2node, node2idx


def scale_boundary(e1, eN, eta=0.85, gamma=0.95):
    """GraphWave's scale range (tools/util.py:120-125)."""
    t = np.sqrt(e1 * eN)
    return -np.log(gamma) / t, -np.log(eta) / t


def recommend_scale_range(eignvalues) -> (float, float):
    """tools/util.py:100-108: first eigenvalue above 1e-3 and the largest one."""
    ev = sorted(eignvalues)
    e1 = next((e for e in ev if e > 0.001), ev[0])
    return scale_boundary(e1, ev[-1])
