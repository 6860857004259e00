"""Shared helpers for the parity tests and bench.py: synthetic inputs of the reference's shapes."""
import numpy as np


def molecular_adjacency(n, rng, self_loops=True):
    """adj = I + S, S symmetric 0/1 'molecular-like': random spanning tree with max degree 4 plus floor(n/8)
    ring-closing edges (the reference builds I + adjacency, SMP_beta.h:505-526)."""
    A = np.zeros((n, n), np.float32)
    deg = np.zeros(n, np.int64)
    order = rng.permutation(n)
    for k in range(1, n):
        v = order[k]
        cands = [u for u in order[:k] if deg[u] < 4]
        u = cands[rng.integers(len(cands))] if cands else order[rng.integers(k)]
        A[u, v] = A[v, u] = 1
        deg[u] += 1
        deg[v] += 1
    extra, tries = n // 8, 0
    while extra > 0 and tries < 1000:
        tries += 1
        u, v = rng.integers(n), rng.integers(n)
        if u != v and A[u, v] == 0 and deg[u] < 4 and deg[v] < 4:
            A[u, v] = A[v, u] = 1
            deg[u] += 1
            deg[v] += 1
            extra -= 1
    if self_loops:
        A += np.eye(n, dtype=np.float32)
    return A


def random_instance(n, C, rng, signed_adj=False):
    T = rng.uniform(-1, 1, (n, n, n, C)).astype(np.float32)
    adj = rng.uniform(-1, 1, (n, n)).astype(np.float32) if signed_adj else molecular_adjacency(n, rng)
    gout = rng.uniform(-1, 1, (n, n, 18 * C)).astype(np.float32)
    return T, adj, gout


def per_slab_errors(x, ref, C):
    """[18] array of max|x-ref| / max|ref_k| (the parity metric)."""
    x = np.asarray(x, np.float64).reshape(-1, 18, C)
    r = np.asarray(ref, np.float64).reshape(-1, 18, C)
    out = np.zeros(18)
    for k in range(18):
        den = np.abs(r[:, k]).max()
        num = np.abs(x[:, k] - r[:, k]).max()
        out[k] = num / den if den > 0 else num
    return out
