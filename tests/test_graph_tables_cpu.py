"""The native graph-table builder (`ccn_graph_tables_*`, graphflow_b200/csrc/graph_tables.cu) against the numpy restatement in
graphflow_b200/graph.py -- which tests/test_model_cpu.py pins to the receptive fields of the unmodified reference models --
on random molecular and Erdos-Renyi graphs, for both model families, including ties in the WL ranking and limited fields."""
import numpy as np
import pytest

from graphflow_b200.graph import GraphTables
from tests.util import molecular_adjacency


def _same(a, b):
    assert np.array_equal(a.features, b.features)
    assert (a.rank is None) == (b.rank is None) and (a.rank is None or np.array_equal(a.rank, b.rank))
    assert a.phi == b.phi
    assert len(a.levels) == len(b.levels)
    for la, lb in zip(a.levels, b.levels):
        for va, vb in zip(la, lb):
            assert va["n"] == vb["n"] and va["src"] == vb["src"] and va["m"] == vb["m"]
            assert np.array_equal(va["adj"], vb["adj"]) and va["adj"].dtype == vb["adj"].dtype
            assert np.array_equal(va["pos"], vb["pos"])


@pytest.mark.parametrize("seed", range(6))
def test_native_tables_equal_the_numpy_restatement_beta(seed):
    rng = np.random.default_rng(seed)
    V, F = int(rng.integers(3, 14)), int(rng.integers(1, 5))
    if seed % 2:
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
    else:
        up = np.triu(rng.integers(0, 2, (V, V)), 1)
        adj = (up + up.T).astype(np.int32)                      # possibly disconnected: INF distances
    feat = np.eye(F)[rng.integers(0, F, V)]                     # few atom types: many ties in the WL ranking
    L, D = int(rng.integers(1, 4)), int(rng.integers(0, 4))
    _same(GraphTables(adj, feat, L, D, native=True), GraphTables(adj, feat, L, D, native=False))


@pytest.mark.parametrize("seed", range(6))
def test_native_tables_equal_the_numpy_restatement_omega(seed):
    rng = np.random.default_rng(100 + seed)
    V, F = int(rng.integers(4, 16)), 3
    adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
    feat = rng.uniform(-1, 1, (V, F))
    L, mf = int(rng.integers(1, 5)), int(rng.integers(2, V + 1))
    _same(GraphTables(adj, feat, L, kind="omega", max_field=mf, native=True),
          GraphTables(adj, feat, L, kind="omega", max_field=mf, native=False))


def test_bad_arguments_are_rejected():
    from graphflow_b200 import _lib

    with pytest.raises(_lib.CCNError):
        GraphTables(np.zeros((3, 3), np.int32), np.zeros((3, 2)), -1, 1)
