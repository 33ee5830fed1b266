"""The on-disk formats either side of the hot path (reference: tools/rw.py): embedding
vectors as CSV (``%.8f``, node id in the first column — the format of the reference's golden
vector tests/robust_test/robust.csv) and distance matrices as ``u v dist`` edge lists."""
import os

import numpy as np
import pandas as pd

__all__ = ["save_vectors", "save_vectors_dict", "read_vectors", "read_distance", "save_distance_edgelist"]


def save_vectors(nodes: list, vectors: list, path: str):
    """tools/rw.py:13-22."""
    pd.DataFrame(data=vectors, index=nodes, columns=None, dtype=float).to_csv(path, header=False, float_format="%.8f")


def save_vectors_dict(vectors: dict, path: str):
    """tools/rw.py:25-31."""
    save_vectors(list(vectors.keys()), list(vectors.values()), path)


def read_vectors(path: str) -> dict:
    """tools/rw.py:34-48: {str(int(node)): [floats]}."""
    if not os.path.exists(path):
        raise FileNotFoundError
    df = pd.read_csv(path, header=None)
    return {str(int(df.iloc[i, 0])): list(df.iloc[i, 1:]) for i in range(df.shape[0])}


def read_distance(path: str, n_nodes: int) -> np.ndarray:
    """tools/rw.py:51-66: symmetric matrix from ``u v dist`` lines (integer node indices)."""
    mat = np.zeros((n_nodes, n_nodes), dtype=float)
    with open(path, mode="r", encoding="utf-8") as fin:
        for line in fin:
            if not line.strip():
                break
            u, v, dist = line.strip().split(" ")
            mat[int(u), int(v)] = mat[int(v), int(u)] = float(dist)
    return mat


def save_distance_edgelist(path: str, nodes: list, mat: np.ndarray):
    """tools/rw.py:84-98: the strict upper triangle, one ``node1 node2 distance`` line per pair."""
    n = len(mat)
    iu, ju = np.triu_indices(n, 1)
    with open(path, mode="w+", encoding="utf-8") as fout:
        fout.writelines(f"{nodes[i]} {nodes[j]} {mat[i, j]}\n" for i, j in zip(iu, ju))
