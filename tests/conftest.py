import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def nx_graph(golden_graphs, name):
    """networkx graph with the reference's node order and string labels."""
    import networkx as nx
    nodes = [str(v) for v in golden_graphs[f"{name}_nodes"]]
    g = nx.Graph()
    g.add_nodes_from(nodes)
    g.add_edges_from((nodes[u], nodes[v]) for u, v in golden_graphs[f"{name}_edges"])
    return g


@pytest.fixture(scope="session")
def golden_graphs():
    return np.load(os.path.join(GOLDEN, "graphs.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_runs():
    return np.load(os.path.join(GOLDEN, "reference_runs.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def robust_csv():
    return np.load(os.path.join(GOLDEN, "robust_csv.npz"), allow_pickle=False)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    # the .so travels with the repo snapshot; build it when missing (nvcc cross-compiles without a GPU)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_hsd_build", os.path.join(ROOT, "hsd_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()
