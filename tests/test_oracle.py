"""The oracle against the reference's golden vectors and known answers (CPU only)."""
import os

import numpy as np
import pytest
import scipy.sparse as spsp

import oracle
from conftest import GOLDEN, random_in_csr


# ------------------------------------------------------------------ RNG contract
def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32_10
    assert oracle.philox4x32_10([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_is_keyed_and_in_range():
    seen = set()
    for t in range(200):
        p = oracle.draw(7, 1, 2, 12345, 1, t, 37)
        assert 0 <= p < 37
        seen.add(p)
    assert len(seen) == 37                       # every position is reachable
    assert oracle.draw(7, 1, 2, 12345, 1, 0, 1000) != oracle.draw(8, 1, 2, 12345, 1, 0, 1000) or \
        oracle.draw(7, 1, 2, 12345, 1, 1, 1000) != oracle.draw(8, 1, 2, 12345, 1, 1, 1000)


# ------------------------------------------------------------------ sampling: structure (SURVEY Appendix A invariants)
def _check_structure(nf, indptr, indices, eids, seeds, fanouts):
    L = len(fanouts)
    assert nf.num_layers == L + 1
    # seed layer: dedup, order preserved
    _, first = np.unique(seeds, return_index=True)
    np.testing.assert_array_equal(nf.layer_parent_nid(L), np.asarray(seeds)[np.sort(first)])
    for j in range(L):
        lay = nf.layer_parent_nid(j)
        assert np.all(np.diff(lay) > 0), "non-seed layers are sorted ascending and unique"
    assert nf.indptr[nf.layer_offsets[1]] == 0 and np.all(nf.indptr[:nf.layer_offsets[1]] == 0)
    for j in range(1, L + 1):
        fan = fanouts[L - j]
        rows = nf.layer_parent_nid(j)
        src_layer = nf.layer_parent_nid(j - 1)
        used = np.zeros(len(src_layer), bool)
        for r, v in enumerate(rows):
            a, b = nf.indptr[nf.layer_offsets[j] + r], nf.indptr[nf.layer_offsets[j] + r + 1]
            deg = indptr[v + 1] - indptr[v]
            assert b - a == min(deg, fan)
            cols = nf.indices[a:b] - nf.layer_offsets[j - 1]
            used[cols] = True
            srcs = src_layer[cols]
            e = nf.edge_mapping[a:b]
            row_src = indices[indptr[v]:indptr[v + 1]]
            row_eid = eids[indptr[v]:indptr[v + 1]] if eids is not None else np.arange(indptr[v], indptr[v + 1])
            # sampled entries are a subsequence of the row: positions strictly increasing
            pos = np.searchsorted(row_eid, e) if np.all(np.diff(row_eid) > 0) else None
            if pos is not None:
                assert np.all(np.diff(pos) > 0)
                np.testing.assert_array_equal(row_src[pos], srcs)
                np.testing.assert_array_equal(row_eid[pos], e)
            if deg <= fan:
                np.testing.assert_array_equal(srcs, row_src)
        assert used.all(), "every node of a layer is the source of some sampled edge"
        assert nf.flow_offsets[j] == nf.indptr[nf.layer_offsets[j + 1]]


@pytest.mark.parametrize("fanouts", [[2, 2], [5, 3], [3], [4, 2, 3]])
def test_sampler_structure(fanouts):
    indptr, indices, eids, _ = random_in_csr(300, 3000, seed=1)
    seeds = np.random.default_rng(2).choice(300, 40, replace=False)
    nf = oracle.sample(indptr, indices, eids, seeds, fanouts, seed=9, epoch=1, batch=3)
    _check_structure(nf, indptr, indices, eids, seeds, fanouts)


def test_sampler_duplicate_seeds_and_empty():
    indptr, indices, eids, _ = random_in_csr(100, 600, seed=3)
    seeds = np.array([5, 9, 5, 7, 9, 9, 1])
    nf = oracle.sample(indptr, indices, eids, seeds, [3, 3])
    np.testing.assert_array_equal(nf.layer_parent_nid(-1), [5, 9, 7, 1])
    nf0 = oracle.sample(indptr, indices, eids, np.zeros(0, np.int64), [3, 3])
    assert nf0.layer_offsets.tolist() == [0, 0, 0, 0] and nf0.flow_offsets.tolist() == [0, 0, 0]


def test_sampler_full_fanout_known_answer():
    """fanout >= max degree: the NodeFlow is RNG-free — the reference's own usage
    (PaGraph/partition/utils.py:11-18, examples/eval.py:20-25). Checked against a plain BFS."""
    indptr, indices, eids, coo = random_in_csr(200, 1500, seed=4)
    csc = coo.tocsc()
    seeds = np.array([3, 17, 42, 99, 150])
    for key in [(0, 0, 0), (5, 2, 7)]:      # must not depend on the RNG key
        nf = oracle.sample(indptr, indices, eids, seeds, [200, 200], seed=key[0], epoch=key[1], batch=key[2])
        hop1 = np.unique(np.concatenate([csc.indices[csc.indptr[v]:csc.indptr[v + 1]] for v in seeds]))
        hop2 = np.unique(np.concatenate([csc.indices[csc.indptr[v]:csc.indptr[v + 1]] for v in hop1]))
        np.testing.assert_array_equal(nf.layer_parent_nid(2), seeds)
        np.testing.assert_array_equal(nf.layer_parent_nid(1), hop1)
        np.testing.assert_array_equal(nf.layer_parent_nid(0), hop2)
        assert nf.flow_offsets[-1] == sum(indptr[v + 1] - indptr[v] for v in hop1) + \
            sum(indptr[v + 1] - indptr[v] for v in seeds)
        # block edges mapped back to parent ids == all in-edges of the dst layer (utils.py:25-30)
        src, dst = [], []
        for j in (1, 2):
            for r, v in enumerate(nf.layer_parent_nid(j)):
                a, b = nf.indptr[nf.layer_offsets[j] + r], nf.indptr[nf.layer_offsets[j] + r + 1]
                src += nf.node_mapping[nf.indices[a:b]].tolist()
                dst += [v] * (b - a)
        got = set(zip(src, dst))
        want = set()
        for v in np.concatenate([seeds, hop1]):
            want |= {(int(u), int(v)) for u in csc.indices[csc.indptr[v]:csc.indptr[v + 1]]}
        assert got == want


def test_sampler_uniformity():
    """Each in-neighbour of a high-degree vertex is picked with probability k/deg (both the direct
    and the complement branch of GetUniformSample)."""
    deg = 30
    indptr = np.array([0, deg] + [deg] * deg, dtype=np.int64)
    indices = np.arange(1, deg + 1, dtype=np.int64)
    for k in (4, 20):                               # deg > 2k  and  k < deg <= 2k
        hits = np.zeros(deg + 1)
        trials = 3000
        for b in range(trials):
            nf = oracle.sample(indptr, indices, None, [0], [k], seed=1, epoch=0, batch=b)
            assert nf.flow_offsets[-1] == k
            hits[nf.layer_parent_nid(0)] += 1
        p = hits[1:] / trials
        assert abs(p.mean() - k / deg) < 1e-9
        assert np.all(np.abs(p - k / deg) < 5 * np.sqrt(k / deg * (1 - k / deg) / trials))


# ------------------------------------------------------------------ sampling: DGL's own generator (libstdc++ minstd_rand0, sequential)
def test_stdlib_engine_is_minstd_rand0():
    """std::default_random_engine of libstdc++ is minstd_rand0 (x <- 16807 x mod 2^31 - 1): with a range of 2^31 - 2
    values uniform_int_distribution passes the engine's output through (minus its minimum, 1), so a vertex of that
    degree would pick position engine() - 1. Checked on a two-vertex graph small enough to build: deg = 3 > 2k = 2,
    k = 1 -> one draw; libstdc++ down-scales by rejection with scaling = (2^31 - 2) / 3."""
    indptr = np.array([0, 3, 3, 3, 3], dtype=np.int64)
    indices = np.array([1, 2, 3], dtype=np.int64)
    for seed in (1, 2, 12345):
        x, rng, scaling = seed, 2 ** 31 - 2, (2 ** 31 - 2) // 3
        while True:                                   # libstdc++ uniform_int_distribution, urngrange > urange branch
            x = (16807 * x) % (2 ** 31 - 1)
            r = x - 1
            if r < 3 * scaling:
                break
        nf = oracle.sample_stdlib(indptr, indices, None, [0], [1], seed=seed)
        assert nf.layer_parent_nid(0).tolist() == [1 + r // scaling]


@pytest.mark.parametrize("fanouts", [[2, 2], [5, 3], [3], [4, 2, 3]])
def test_stdlib_sampler_structure(fanouts):
    indptr, indices, eids, _ = random_in_csr(300, 3000, seed=1)
    seeds = np.random.default_rng(2).choice(300, 40, replace=False)
    nf = oracle.sample_stdlib(indptr, indices, eids, seeds, fanouts, seed=7)
    _check_structure(nf, indptr, indices, eids, seeds, fanouts)
    # sequential stream: the same seed reproduces the minibatch, another seed draws another one
    same = oracle.sample_stdlib(indptr, indices, eids, seeds, fanouts, seed=7)
    other = oracle.sample_stdlib(indptr, indices, eids, seeds, fanouts, seed=8)
    np.testing.assert_array_equal(nf.edge_mapping, same.edge_mapping)
    assert not np.array_equal(nf.edge_mapping, other.edge_mapping)


def test_stdlib_sampler_equals_counter_based_sampler_when_rng_free():
    """fanout >= max degree (how the reference calls the sampler for partitioning and evaluation): both generators must
    give the same NodeFlow, array for array — discovery-order expansion and sorted expansion meet in ConstructNodeFlow."""
    indptr, indices, eids, _ = random_in_csr(200, 1500, seed=4)
    seeds = np.array([3, 17, 42, 99, 150, 17])
    a = oracle.sample(indptr, indices, eids, seeds, [200, 200, 200], seed=3, epoch=1, batch=2)
    b = oracle.sample_stdlib(indptr, indices, eids, seeds, [200, 200, 200], seed=5)
    for name in ("node_mapping", "layer_offsets", "indptr", "indices", "edge_mapping", "flow_offsets"):
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=name)


def test_stdlib_sampler_uniformity_matches_counter_based():
    """Inclusion frequencies under DGL's generator agree with k / deg (direct and complement branch). One sequential
    stream over many vertices that share a neighbour list (consecutive small seeds would only show minstd_rand0's
    first output, which is 16807 * seed)."""
    deg, trials = 30, 3000
    indptr = np.concatenate([np.arange(trials + 1) * deg, np.full(deg, trials * deg)]).astype(np.int64)
    indices = np.tile(np.arange(trials, trials + deg, dtype=np.int64), trials)
    for k in (4, 20):
        nf = oracle.sample_stdlib(indptr, indices, None, np.arange(trials), [k], seed=12345)
        assert nf.flow_offsets[-1] == k * trials
        src = nf.layer_parent_nid(0)[nf.indices - nf.layer_offsets[0]]
        p = np.bincount(src - trials, minlength=deg) / trials
        assert abs(p.mean() - k / deg) < 1e-9
        assert np.all(np.abs(p - k / deg) < 5 * np.sqrt(k / deg * (1 - k / deg) / trials))


def test_sampler_depends_on_epoch_and_batch_only_when_sampling():
    indptr, indices, eids, _ = random_in_csr(300, 6000, seed=5)
    seeds = np.arange(0, 300, 7)
    a = oracle.sample(indptr, indices, eids, seeds, [3, 3], seed=1, epoch=0, batch=0)
    b = oracle.sample(indptr, indices, eids, seeds, [3, 3], seed=1, epoch=0, batch=0)
    c = oracle.sample(indptr, indices, eids, seeds, [3, 3], seed=1, epoch=1, batch=0)
    np.testing.assert_array_equal(a.node_mapping, b.node_mapping)
    np.testing.assert_array_equal(a.edge_mapping, b.edge_mapping)
    assert not np.array_equal(a.edge_mapping, c.edge_mapping)


# ------------------------------------------------------------------ gather: golden vectors from the real storage.py
STORAGE_CASES = ["storage_gcn_partial", "storage_ties_odd", "storage_full", "storage_sage4"]


class _NF:
    def __init__(self, g):
        self.node_mapping, self.layer_offsets = g["node_mapping"], g["layer_offsets"]
        self.num_layers = len(self.layer_offsets) - 1

    def layer_parent_nid(self, i):
        return self.node_mapping[self.layer_offsets[i]:self.layer_offsets[i + 1]]


@pytest.mark.parametrize("case", STORAGE_CASES)
def test_oracle_cache_matches_reference_storage(case):
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    names = [str(x) for x in g["field_names"]]
    host = {n: g["host_" + n] for n in names}
    oc = oracle.OracleCache(host, len(g["nid_map"]), g["nid_map"])
    oc.init_field(names)
    assert oc.total_dim == int(g["total_dim"])
    nf = _NF(g)
    cold = oc.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in names:
            np.testing.assert_array_equal(cold[i][0][n], g["cold_l%d_%s" % (i, n)])
    assert oc.miss_num / oc.try_num == float(g["cold_miss_rate"]) == 1.0
    oc.try_num = oc.miss_num = 0
    oc.auto_cache(g["out_deg"], names, int(g["capability"]))
    assert oc.full_cached == bool(g["full_cached"]) and oc.cached_num == int(g["cached_num"])
    np.testing.assert_array_equal(oc.gpu_flag, g["gpu_flag"])            # the cache hit SET
    for n in names:                                                     # same rows cached (order-free)
        ref = g["cache_" + n][g["l2c_on_cached"][g["gpu_flag"]]]
        mine = oc.gpu_fix_cache[n][oc.localid2cacheid[oc.gpu_flag]]
        np.testing.assert_array_equal(mine, ref)
    warm = oc.fetch_data(nf)
    for i in range(nf.num_layers):
        for n in names:
            np.testing.assert_array_equal(warm[i][0][n], g["warm_l%d_%s" % (i, n)])
    if not oc.full_cached:
        assert oc.miss_num == int(g["warm_miss_num"]) and oc.try_num == int(g["warm_try_num"])


def test_fetch_c_matches_numpy_restatement():
    g = np.load(os.path.join(GOLDEN, "storage_gcn_partial.npz"))
    host = {"features": g["host_features"]}
    oc = oracle.OracleCache(host, len(g["nid_map"]), g["nid_map"])
    oc.init_field(["features"])
    oc.auto_cache(g["out_deg"], ["features"], 60)
    frame, mask = oc.fetch_layer(g["node_mapping"])
    out, cmask, miss = oracle.fetch_c(g["node_mapping"], oc.gpu_flag, oc.localid2cacheid, oc.nid_map,
                                      oc.gpu_fix_cache["features"], host["features"], threads=2)
    np.testing.assert_array_equal(out, frame["features"])
    np.testing.assert_array_equal(cmask, mask)
    assert miss == int((~mask).sum())


# ------------------------------------------------------------------ aggregation
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_aggregate_matches_scipy_float64(mode):
    indptr, indices, eids, _ = random_in_csr(150, 900, seed=6)
    nf = oracle.sample(indptr, indices, eids, np.arange(0, 150, 5), [4, 3], seed=3)
    rng = np.random.default_rng(0)
    for i in range(nf.num_blocks):
        ip, cols, base = nf.block(i)
        n_src = len(nf.layer_parent_nid(i))
        x = rng.standard_normal((n_src, 13)).astype(np.float32)
        A = spsp.csr_matrix((np.ones(ip[-1] - ip[0]), cols[ip[0]:ip[-1]] - base, ip - ip[0]),
                            shape=(len(ip) - 1, n_src))
        want = A @ x.astype(np.float64)
        if mode == "mean":
            want /= np.maximum(np.diff(ip), 1)[:, None]
        got = oracle.aggregate(ip, cols, base, x, mode)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
        gd = rng.standard_normal((len(ip) - 1, 13)).astype(np.float32)
        gw = gd.astype(np.float64)
        if mode == "mean":
            gw = gw / np.maximum(np.diff(ip), 1)[:, None]
        np.testing.assert_allclose(oracle.aggregate_bwd(ip, cols, base, gd, n_src, mode), A.T @ gw,
                                   rtol=1e-6, atol=1e-6)


def test_aggregate_zero_degree_rows_are_zero():
    indptr = np.array([0, 0, 2, 2], np.int64)
    cols = np.array([10, 11], np.int64)
    x = np.ones((2, 4), np.float32) * 3
    out = oracle.aggregate(indptr, cols, 10, x, "mean")
    np.testing.assert_array_equal(out, [[0] * 4, [3] * 4, [0] * 4])


@pytest.mark.parametrize("seed,n,dim,p", [(0, 7, 600, 0.2), (0xDEADBEEF12345, 300, 602, 0.5), (2 ** 64 - 3, 50, 64, 0.2),
                                          (41, 33, 13, 0.9), (5, 10, 32, 0.0)])
def test_dropout_mask_contract_library_vs_oracle(seed, n, dim, p):
    """The kernels' mask function (pg_common.cuh drop_hash, evaluated on the host by pg_dropout_keep_mask: no GPU work)
    and the oracle's numpy restatement are the same function."""
    import ctypes
    from pagraph_b200 import _lib
    out = np.zeros((n, dim), dtype=np.uint8)
    _lib.check(_lib.lib().pg_dropout_keep_mask(seed, n, dim, p, out.ctypes.data_as(ctypes.c_void_p)), "pg_dropout_keep_mask")
    want = oracle.dropout_keep_mask(seed, n, dim, p)
    assert np.array_equal(out.astype(bool), want)
    if p == 0.0:
        assert out.all()


def test_dropout_mask_statistics():
    """Bernoulli(1 - p) marginals per 16-bit lane, no correlation between neighbouring columns / rows / steps, and the
    step is hashed into the key (consecutive steps do not shift the mask)."""
    keep = oracle.dropout_keep_mask(12345, 4000, 600, 0.2)
    assert abs(keep.mean() - 0.8) < 2e-3
    for lane in range(4):
        assert abs(keep[:, lane::4].mean() - 0.8) < 4e-3
    x = keep.astype(np.float64) - keep.mean()
    var = x.var()
    assert abs((x[:, :-1] * x[:, 1:]).mean() / var) < 5e-3
    assert abs((x[:, :-4] * x[:, 4:]).mean() / var) < 5e-3
    assert abs((x[:-1] * x[1:]).mean() / var) < 5e-3
    nxt = oracle.dropout_keep_mask(12346, 4000, 600, 0.2).astype(np.float64) - 0.8
    for a, b in ((x, nxt), (x[:, 4:], nxt[:, :-4]), (x[1:], nxt[:-1]), (x[:, :-4], nxt[:, 4:])):
        assert abs((a * b).mean() / var) < 5e-3
    # a golden word pins the constants (computed with Python integers)
    M = 2 ** 64 - 1

    def sm(v):
        v = (v + 0x9E3779B97F4A7C15) & M
        v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & M
        v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & M
        return v ^ (v >> 31)
    seed, j, c = 12345, 17, 42
    h = ((sm((sm(seed) + j) & M) ^ sm(0xD1B54A32D192ED03 + c // 4)) * 0x9E3779B97F4A7C15) & M
    h ^= h >> 32
    assert bool(keep[j, c]) == (((h >> (16 * (c % 4))) & 0xFFFF) >= int(np.float32(0.2) * np.float32(65536.0) + np.float32(0.5)))
