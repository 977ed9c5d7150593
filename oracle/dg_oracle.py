"""Plain-Python restatement of the reference's streaming partitioner — TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows PaGraph/partition/dg.py: in_neighbors_hop :18-27, dg_max_score :30-35, dg_ind :38-56, dg :59-103.
PINNED by tests/golden/dg_*.npz, which tests/golden/make_golden.py produced by running the real dg.py.
Small inputs only (pure Python loops)."""
import numpy as np


def _in_lists(row, col, V):
    """in-neighbour lists, ascending and de-duplicated (what adj.tocsc() yields, dg.py:60)."""
    lists = [[] for _ in range(V)]
    for s, d in sorted(set(zip(row.tolist(), col.tolist())), key=lambda e: (e[1], e[0])):
        lists[d].append(s)
    return lists


def _neighbourhood(lists, nid, hops):
    if hops == 1:                                            # dg.py:19-20
        return list(lists[nid])
    appended = []                                            # the list `nids` of dg.py:22
    for _ in range(hops):                                    # dg.py:23-26
        frontier = appended[-1] if appended else [nid]       # NB: only the list appended LAST (reference quirk)
        for n in list(frontier):
            appended.append(lists[n])
    return sorted(set(x for lst in appended for x in lst))   # np.unique(np.hstack(nids)), dg.py:27


def dg(P, row, col, V, train, hops):
    """Returns (belongs int array [V] (-1 = unassigned), member bool [P, V])."""
    lists = _in_lists(np.asarray(row), np.asarray(col), V)
    belongs = -np.ones(V, dtype=np.int64)
    member = np.zeros((P, V), dtype=bool)
    p_vnum = np.zeros(P, dtype=np.int64)
    r_vnum = np.zeros(P, dtype=np.int64)
    avg = V * 0.65 / P                                        # dg.py:54
    for nid in np.asarray(train).tolist():
        neigh = _neighbourhood(lists, nid, hops)
        common = np.ones(P, dtype=np.int64)                   # dg.py:47
        for w in neigh:
            if belongs[w] != -1:
                common[belongs[w]] += 1
        score = common * (-p_vnum + avg) / (r_vnum + 1)       # dg.py:55, float64
        order = np.argsort(score, kind="stable")[-2:]         # dg.py:31 (tie order: see DESIGN.md)
        lo, hi = int(order[0]), int(order[1])
        if score[lo] != score[hi]:
            ind = hi
        else:
            ind = lo if p_vnum[lo] < p_vnum[hi] else hi       # dg.py:32-35
        if belongs[nid] == -1:                                # dg.py:76-83
            belongs[nid] = ind
            p_vnum[ind] += 1
            for w in neigh + [nid]:
                if not member[ind, w]:
                    member[ind, w] = True
                    r_vnum[ind] += 1
    return belongs, member
