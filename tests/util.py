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


# ---- one CCN level from the level l-1 tensors (promotion -> stack -> contraction -> mix): tables and oracle chain ----------
def level_tables(fields_prev, fields_cur, C, n_max, f_base=0):
    """Promotion tables of one graph (SMP_beta.h:446-459, 588-594 restated as index tables).  fields_prev[w] / fields_cur[v]:
    ordered member lists of phi_{l-1}(w) / phi_l(v).  The level l-1 tensors f[w] ([m_w, m_w, C]) are packed back to back from
    element f_base.  Returns (f_off [V*n_max] int64, m [V*n_max] int32, pos [V*n_max*n_max] int32, n [V] int32, f_size)."""
    V = len(fields_cur)
    starts, off = [], f_base
    for w in range(len(fields_prev)):
        starts.append(off)
        off += len(fields_prev[w]) ** 2 * C
    f_off = np.zeros((V, n_max), np.int64)
    m = np.ones((V, n_max), np.int32)
    pos = -np.ones((V, n_max, n_max), np.int32)
    n = np.zeros(V, np.int32)
    for v in range(V):
        phi = list(fields_cur[v])
        n[v] = len(phi)
        for a, w in enumerate(phi):
            prev = list(fields_prev[w])
            where = {u: i for i, u in enumerate(prev)}
            f_off[v, a] = starts[w]
            m[v, a] = len(prev)
            for i, u in enumerate(phi):
                pos[v, a, i] = where.get(u, -1)
    return f_off.reshape(-1), m.reshape(-1), pos.reshape(-1), n, off - f_base


def oracle_gather_level(f, f_off, m, pos, n, adj, K, bias, gZ, n_max, C, alpha=0.01):
    """fp64 oracle of ccn_gather_level_forward + _backward for a batch of instances: promotion (the plain-C restatement of
    MatTensorMul + TensorMatMul with 0/1 selection matrices), StackTensor3D, RisiContraction_18 (einsum statement pinned to the
    compiled reference in tests/test_oracle_cpu.py), MatMul, VectorAddTensor, LeakyReLU3D and all their backward passes.
    f: flat level l-1 buffer; adj [B, n_max*n_max] (compact per instance); gZ [B, n_max*n_max, C_out] (compact rows).
    Returns per-instance lists X, Z and the sums gf (flat), gK, gbias."""
    from oracle import pyoracle

    orc = pyoracle.COracle("f64")
    B = len(n)
    Co = K.shape[1]
    f = np.asarray(f, np.float64)
    gf = np.zeros_like(f)
    gK = np.zeros(K.shape, np.float64)
    gb = np.zeros(Co, np.float64)
    Xs, Zs = [], []
    for i in range(B):
        ni = int(n[i])
        if ni == 0:
            Xs.append(None)
            Zs.append(None)
            continue
        A = np.asarray(adj[i].reshape(-1)[:ni * ni], np.float64).reshape(ni, ni)
        T = np.zeros((ni, ni, ni, C))
        srcs = []
        for a in range(ni):
            o, mm = int(f_off[i * n_max + a]), int(m[i * n_max + a])
            p = pos[(i * n_max + a) * n_max:(i * n_max + a) * n_max + ni]
            srcs.append((o, mm, p))
            T[a] = orc.promote_forward(f[o:o + mm * mm * C].reshape(mm, mm, C), p)
        X = pyoracle.einsum18_forward(T, A).reshape(ni * ni, 18 * C)
        Y = orc.matmul_forward(X, K)
        Z = orc.bias_lrelu_forward(Y, bias, alpha)
        Xs.append(X)
        Zs.append(Z)
        gz = np.asarray(gZ[i].reshape(-1, Co)[:ni * ni], np.float64)
        gY, gbi = orc.bias_lrelu_backward(Y, bias, gz, alpha)
        gX, gKi = orc.matmul_backward(X, K, gY)
        gK += gKi
        gb += gbi
        gT = pyoracle.einsum18_backward(gX.reshape(ni, ni, 18 * C), A)
        for a, (o, mm, p) in enumerate(srcs):
            gf[o:o + mm * mm * C] += orc.promote_backward(gT[a], p, mm).reshape(-1)
    return Xs, Zs, gf, gK, gb


def line_graph(adj, feat):
    """The line graph of (adj [V,V], feat [V,F]): one vertex per undirected edge {i < j}, two of them adjacent when the edges share
    an endpoint; an edge's feature = the sum of its endpoints' features.  What the callers of SMP_omega_pairgraphs pass as
    molecule_2."""
    import numpy as np

    V = adj.shape[0]
    edges = [(i, j) for i in range(V) for j in range(i + 1, V) if adj[i, j] or adj[j, i]]
    E = len(edges)
    a2 = np.zeros((E, E), np.int32)
    for x, (i, j) in enumerate(edges):
        for y, (k, l) in enumerate(edges):
            if x != y and len({i, j} & {k, l}) > 0:
                a2[x, y] = 1
    f2 = np.stack([feat[i] + feat[j] for i, j in edges]) if E else np.zeros((0, feat.shape[1]))
    return a2, f2
