"""dg partitioner: the oracle restatement and the native implementation (pg_partition_dg, host code — runs without a
GPU) against golden vectors produced by the real reference dg.py, and against each other on random graphs."""
import os

import numpy as np
import pytest
import scipy.sparse as spsp

from conftest import GOLDEN
from oracle import dg_oracle
from pagraph_b200.partition import dg as dg_mod


def _golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    V, P, hops = int(g["V"]), int(g["P"]), int(g["hops"])
    return g, V, P, hops


@pytest.mark.parametrize("name", ["dg_p2_h1", "dg_p4_h2"])
def test_oracle_matches_reference_golden(name):
    g, V, P, hops = _golden(name)
    belongs, member = dg_oracle.dg(P, g["row"], g["col"], V, g["train"], hops)
    for p in range(P):
        np.testing.assert_array_equal(np.where(belongs == p)[0], g["sub_trainv_%d" % p])
        np.testing.assert_array_equal(np.where(member[p])[0], g["sub_v_%d" % p])


@pytest.mark.parametrize("name", ["dg_p2_h1", "dg_p4_h2"])
def test_native_matches_reference_golden(name):
    g, V, P, hops = _golden(name)
    adj = spsp.coo_matrix((np.ones(len(g["row"]), np.int64), (g["row"], g["col"])), shape=(V, V))
    sub_v, sub_trainv = dg_mod.dg(P, adj, g["train"], hops)
    for p in range(P):
        np.testing.assert_array_equal(sub_trainv[p], g["sub_trainv_%d" % p])
        np.testing.assert_array_equal(sub_v[p], g["sub_v_%d" % p])


@pytest.mark.parametrize("P,hops,seed", [(2, 1, 0), (3, 2, 1), (5, 3, 2), (8, 2, 3), (2, 3, 4)])
def test_native_matches_oracle_random(P, hops, seed):
    rng = np.random.default_rng(seed)
    V, nnz = 400, 1600
    row, col = rng.integers(0, V, nnz), rng.integers(0, V, nnz)
    train = np.sort(rng.choice(V, 260, replace=False)).astype(np.int64)
    adj = spsp.coo_matrix((np.ones(nnz, np.int64), (row, col)), shape=(V, V))
    belongs, member = dg_oracle.dg(P, row, col, V, train, hops)
    sub_v, sub_trainv = dg_mod.dg(P, adj, train, hops)
    assert sum(len(t) for t in sub_trainv) == len(train)
    for p in range(P):
        np.testing.assert_array_equal(sub_trainv[p], np.where(belongs == p)[0])
        np.testing.assert_array_equal(sub_v[p], np.where(member[p])[0])


def test_native_rejects_bad_arguments():
    from pagraph_b200 import _lib
    adj = spsp.coo_matrix((np.ones(3), ([0, 1, 2], [1, 2, 0])), shape=(3, 3))
    with pytest.raises(_lib.PGError):
        dg_mod.dg(1, adj, np.array([0, 1]), 1)           # the reference's argsort(...)[-2:] needs >= 2 partitions
    with pytest.raises(_lib.PGError):
        dg_mod.dg(2, adj, np.array([0, 7]), 1)           # train id out of range
