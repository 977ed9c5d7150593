"""Dataset / partition file loaders (same names and on-disk formats as the reference
PaGraph/data/get_data.py:8-103 — SURVEY.md Appendix C) and the synthetic-data recipe of
PaGraph/data/preprocess.py:50-114 (random features / labels / 65-10-25 split) plus an R-MAT
generator standing in for the PaRMAT binary (README.md:36-40), which is unavailable offline.
"""
import os

import numpy as np
import scipy.sparse


# ---------------------------------------------------------------- loaders (reference names)
def get_graph_data(dataname):
    """adj.npz (COO, row=src, col=dst) and feat.npy; random [V, 600] features if feat.npy is absent
    (get_data.py:8-29)."""
    adj = scipy.sparse.load_npz(os.path.join(dataname, 'adj.npz'))
    try:
        feat = np.load(os.path.join(dataname, 'feat.npy'))
    except FileNotFoundError:
        print('random generate feat...')
        import torch
        feat = torch.rand((adj.shape[0], 600))
    return adj, feat


def get_struct(dataname):
    return scipy.sparse.load_npz(os.path.join(dataname, 'adj.npz'))


def get_masks(dataname):
    return tuple(np.load(os.path.join(dataname, f + '.npy')) for f in ('train', 'val', 'test'))


def get_labels(dataname):
    return np.load(os.path.join(dataname, 'labels.npy'))


def _part_dir(dataname, partitions):
    return os.path.join(dataname, '{}naive'.format(partitions))


def get_sub_train_graph(dataname, idx, partitions):
    """(subadj CSR in sub-graph ids, sub id -> full id) of partition `idx` (get_data.py:32-47)."""
    d = _part_dir(dataname, partitions)
    adj = scipy.sparse.load_npz(os.path.join(d, 'subadj_{}.npz'.format(idx)))
    train2fullid = np.load(os.path.join(d, 'sub_train2fullid_{}.npy'.format(idx)))
    return adj, train2fullid


def get_sub_train_nid(dataname, idx, partitions):
    return np.load(os.path.join(_part_dir(dataname, partitions), 'sub_trainid_{}.npy'.format(idx)))


def get_sub_train_labels(dataname, idx, partitions):
    return np.load(os.path.join(_part_dir(dataname, partitions), 'sub_label_{}.npy'.format(idx)))


def get_feat_from_server(g, nids, embed_name):
    """CPU rows of field `embed_name` for full-graph ids `nids` (get_data.py:105-116)."""
    return g._node_frame._frame[embed_name].data[nids]


# ---------------------------------------------------------------- synthetic recipe (preprocess.py)
def random_feature(vnum, feat_size, seed=2):
    return np.random.default_rng(seed).random((vnum, feat_size), dtype=np.float32)


def random_label(vnum, class_num, seed=3):
    return np.random.default_rng(seed).integers(0, class_num, size=vnum).astype(np.int64)


def split_dataset(vnum, seed=4):
    """train:val:test = 6.5:1:2.5 masks (preprocess.py:83-114)."""
    nids = np.random.default_rng(seed).permutation(vnum)
    train_len, val_len = int(vnum * 0.65), int(vnum * 0.1)
    masks = [np.zeros(vnum, dtype=np.int64) for _ in range(3)]
    masks[0][nids[:train_len]] = 1
    masks[1][nids[train_len:train_len + val_len]] = 1
    masks[2][nids[train_len + val_len:]] = 1
    return tuple(masks)


def rmat_pairs_numpy(vnum, n_pairs, seed=1, a=0.45, b=0.22, c=0.22):
    """n_pairs distinct undirected pairs (u<v) from an R-MAT process; ids >= vnum are rejected."""
    rng = np.random.default_rng(seed)
    scale = max(1, int(np.ceil(np.log2(max(vnum, 2)))))
    have = np.zeros(0, dtype=np.int64)
    while len(have) < n_pairs:
        m = int((n_pairs - len(have)) * 1.5) + 1024
        src = np.zeros(m, np.int64)
        dst = np.zeros(m, np.int64)
        for _ in range(scale):
            r = rng.random(m)
            src = (src << 1) | (r > a + b)
            dst = (dst << 1) | (((r > a) & (r <= a + b)) | (r > a + b + c))
        ok = (src < vnum) & (dst < vnum) & (src != dst)
        lo, hi = np.minimum(src[ok], dst[ok]), np.maximum(src[ok], dst[ok])
        have = np.unique(np.concatenate([have, lo * vnum + hi]))
    if len(have) > n_pairs:
        have = np.sort(rng.choice(have, n_pairs, replace=False))
    return have // vnum, have % vnum


def rmat_adj(vnum, nnz, seed=1):
    """Symmetrised COO adjacency with exactly `nnz` (even) entries, as pp2adj(is_direct=False) builds it
    (preprocess.py:36-43): (src ++ dst, dst ++ src)."""
    u, v = rmat_pairs_numpy(vnum, nnz // 2, seed)
    row, col = np.concatenate([u, v]), np.concatenate([v, u])
    return scipy.sparse.coo_matrix((np.ones(len(row), np.int64), (row, col)), shape=(vnum, vnum))


def rmat_in_csr_cuda(vnum, nnz, seed=1, device="cuda", a=0.45, b=0.22, c=0.22):
    """Same process at benchmark scale, generated and sorted on the GPU (setup only, not the hot
    path). Returns the in-CSR (indptr, indices) as CUDA int64 tensors; edge id = CSR position."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_pairs = nnz // 2
    scale = max(1, int(np.ceil(np.log2(max(vnum, 2)))))
    have = torch.zeros(0, dtype=torch.int64, device=device)
    while have.numel() < n_pairs:
        m = int((n_pairs - have.numel()) * 1.6) + 4096
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            src = (src << 1) | (r > a + b)
            dst = (dst << 1) | (((r > a) & (r <= a + b)) | (r > a + b + c))
        ok = (src < vnum) & (dst < vnum) & (src != dst)
        src, dst = src[ok], dst[ok]
        key = torch.minimum(src, dst) * vnum + torch.maximum(src, dst)
        del src, dst, ok
        have = torch.unique(torch.cat([have, key]))
        del key
    if have.numel() > n_pairs:
        keep = torch.randperm(have.numel(), device=device, generator=gen)[:n_pairs]
        have = have[keep]
    u, v = have // vnum, have % vnum
    del have
    dst = torch.cat([v, u])
    src = torch.cat([u, v])
    del u, v
    key, _ = torch.sort(dst * vnum + src)     # rows by dst, sources ascending inside a row
    del dst, src
    indices = key % vnum
    counts = torch.bincount(key // vnum, minlength=vnum)
    del key
    indptr = torch.zeros(vnum + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, indices.contiguous()


def write_dataset(dataname, adj, feat=None, labels=None, masks=None):
    """Write the reference on-disk layout (Appendix C)."""
    os.makedirs(dataname, exist_ok=True)
    scipy.sparse.save_npz(os.path.join(dataname, 'adj.npz'), adj.tocoo())
    if feat is not None:
        np.save(os.path.join(dataname, 'feat.npy'), feat)
    if labels is not None:
        np.save(os.path.join(dataname, 'labels.npy'), labels)
    if masks is not None:
        for name, m in zip(('train', 'val', 'test'), masks):
            np.save(os.path.join(dataname, name + '.npy'), m)
