"""Host-side graph preprocessing of the second-order CCN models -> index tables for the device kernels.

Restates, on numpy arrays, what SMP_beta::complete_computation_graph derives from a DenseGraph before it wires any
operator (GraphFlow/SMP_beta.h:531-552): shortest paths (floyd_warshall, :343-365), the Weisfeiler-Lehman histogram
features (weisfeiler_lehman, :367-389), the vertex ranking (rank_vertices, :403-419), the receptive fields phi_l(v)
(init_receptive_field_permutation_matrix_reduced_adj, :461-489, ordered by rank, :435-444) and the reduced adjacency
matrices (:505-526).  The 0/1 selection matrices X[v][w] / X^T (init_permutation_matrix, :446-459, 491-502) are never
materialised: they become the `pos` index table of ccn_promote_forward (position of phi_l(v)[i] inside phi_{l-1}(w),
or -1), which is what turns MatTensorMul + TensorMatMul into a gather.

The tables depend on the graph only, not on the weights, so they are built once per graph and cached by the caller
(the reference rebuilds them three times per example per BatchLearn, SMP_beta.h:753,758,770).
"""
import numpy as np

INF = 10 ** 9  # SMP_beta::INF (SMP_beta.h:1011)


def shortest_paths(adj):
    """floyd_warshall (SMP_beta.h:343-365) for a symmetric 0/1 adjacency (no self loops needed)."""
    V = adj.shape[0]
    sp = np.full((V, V), INF, np.int64)
    sp[adj > 0] = 1
    sp[(adj > 0).T] = 1
    np.fill_diagonal(sp, 0)
    for k in range(V):
        sp = np.minimum(sp, sp[:, k:k + 1] + sp[k:k + 1, :])
    return sp


def wl_features(sp, feat, n_depth):
    """weisfeiler_lehman (SMP_beta.h:367-389): hist[v, d*F + f] = sum of feat[u, f] over u at distance d from v."""
    V, F = feat.shape
    hist = np.zeros((V, F * (n_depth + 1)), np.float64)
    for d in range(n_depth + 1):
        hist[:, d * F:(d + 1) * F] = (sp.T == d).astype(np.float64) @ feat
    return hist


def _compare(hu, hv):
    """compare_vertices (SMP_beta.h:391-401): lexicographic order of the histogram rows."""
    for x, y in zip(hu, hv):
        if x < y:
            return -1
        if x > y:
            return 1
    return 0


def vertex_rank(hist):
    """rank_vertices (SMP_beta.h:403-419), the reference's exact exchange sort (ties resolve the way it does)."""
    V = hist.shape[0]
    order = list(range(V))
    rows = [tuple(r) for r in hist]
    for i in range(V):
        for j in range(i + 1, V):
            if _compare(rows[order[i]], rows[order[j]]) < 0:
                order[i], order[j] = order[j], order[i]
    rank = np.zeros(V, np.int64)
    for i, v in enumerate(order):
        rank[v] = i
    return rank


def receptive_fields(sp, rank, n_levels):
    """phi[l][v] (SMP_beta.h:461-489): phi_0(v) = {v}; phi_l(v) = union of phi_{l-1}(u) over u within distance 1 of v,
    sorted by rank (ranks are distinct, so the reference's exchange sort is a plain sort)."""
    V = sp.shape[0]
    phi = [[[v] for v in range(V)]]
    for l in range(1, n_levels + 1):
        cur = []
        for v in range(V):
            members = []
            for u in range(V):
                if sp[u, v] <= 1:
                    for w in phi[l - 1][u]:
                        if w not in members:
                            members.append(w)
            members.sort(key=lambda w: rank[w])
            cur.append(members)
        phi.append(cur)
    return phi


def limit_receptive_field(sp, v, members, max_field):
    """limit_receptive_field (SMP_omega_physics.h:367-392): the reference's exchange sort by distance from v (not a
    stable sort: reproduced literally), then whole outermost distance shells are dropped until the field fits."""
    A = list(members)
    for i in range(len(A)):
        for j in range(i + 1, len(A)):
            if sp[v, A[i]] > sp[v, A[j]]:
                A[i], A[j] = A[j], A[i]
    while len(A) > max_field:
        d = sp[v, A[-1]]
        while sp[v, A[-1]] == d:
            A.pop()
    return A


def receptive_fields_omega(sp, n_levels, max_field):
    """SMP_omega_physics (SMP_omega_physics.h:394-418): no rank ordering -- the union keeps insertion order (u ascending,
    then phi_{l-1}(u)'s order) -- and fields larger than max_field are cut by limit_receptive_field."""
    V = sp.shape[0]
    phi = [[[v] for v in range(V)]]
    for l in range(1, n_levels + 1):
        cur = []
        for v in range(V):
            members = []
            for u in range(V):
                if sp[u, v] <= 1:
                    for w in phi[l - 1][u]:
                        if w not in members:
                            members.append(w)
            if len(members) > max_field:
                members = limit_receptive_field(sp, v, members, max_field)
            cur.append(members)
        phi.append(cur)
    return phi


def receptive_fields_omega_wl(sp, rank, n_levels, max_field):
    """SMP_omega (SMP_omega.h:512-530): the union as in SMP_beta; a field larger than max_field is first cut by
    limit_receptive_field (:476-510: members ordered by distance from v, ties by rank, then whole outermost distance shells
    dropped until it fits), and the survivors are ordered by rank (:451-459)."""
    V = sp.shape[0]
    phi = [[[v] for v in range(V)]]
    for l in range(1, n_levels + 1):
        cur = []
        for v in range(V):
            members = []
            for u in range(V):
                if sp[u, v] <= 1:
                    for w in phi[l - 1][u]:
                        if w not in members:
                            members.append(w)
            if len(members) > max_field:
                members.sort(key=lambda a: (sp[v, a], rank[a]))
                while len(members) > max_field:
                    d = sp[v, members[-1]]
                    while members and sp[v, members[-1]] == d:
                        members.pop()
            members.sort(key=lambda a: rank[a])
            cur.append(members)
        phi.append(cur)
    return phi


class GraphTables:
    """Everything the device path needs for one graph: WL input features and, per level l >= 1 and vertex v,
    n = |phi_l(v)|, the reduced adjacency [n, n] and for every slab a (w = phi_l(v)[a]) the source vertex w, the side
    m = |phi_{l-1}(w)| of its level l-1 tensor and the gather positions pos[a][i]."""

    def __init__(self, adj, feat, n_levels, n_depth=None, kind="beta", max_field=None, native=True):
        """kind = "beta": SMP_beta / SMP_2D_ver8 (WL features of depth n_depth, rank-ordered fields);
        kind = "omega": SMP_omega_physics (raw features, insertion-ordered fields limited to max_field members).
        native=True builds the tables with the C++ implementation behind the C-ABI (`ccn_graph_tables_*`,
        csrc/graph_tables.cu); native=False runs the numpy restatement in this file (what the tests compare it with)."""
        adj = np.asarray(adj)
        feat = np.asarray(feat, np.float64)
        self.V = adj.shape[0]
        self.n_levels = n_levels
        if native:
            self._build_native(adj, feat, n_levels, n_depth, kind, max_field)
            return
        sp = shortest_paths(adj)
        if kind == "omega":
            self.features = feat.copy()
            self.rank = None
            self.phi = receptive_fields_omega(sp, n_levels, max_field if max_field is not None else self.V)
        elif kind == "omega_wl":
            self.features = wl_features(sp, feat, n_depth)
            self.rank = vertex_rank(self.features)
            self.phi = receptive_fields_omega_wl(sp, self.rank, n_levels, max_field if max_field is not None else self.V)
        else:
            self.features = wl_features(sp, feat, n_depth)
            self.rank = vertex_rank(self.features)
            self.phi = receptive_fields(sp, self.rank, n_levels)
        self.levels = []
        for l in range(1, n_levels + 1):
            per_vertex = []
            for v in range(self.V):
                field = self.phi[l][v]
                n = len(field)
                red = np.zeros((n, n), np.float32)
                for i, a in enumerate(field):
                    for j, b in enumerate(field):
                        red[i, j] = 1.0 if a == b else float(adj[a, b])  # SMP_beta.h:516-520
                src, m, pos = [], [], np.full((n, n), -1, np.int32)
                for a, w in enumerate(field):
                    prev = self.phi[l - 1][w]
                    index = {u: k for k, u in enumerate(prev)}
                    src.append(w)
                    m.append(len(prev))
                    for i, u in enumerate(field):
                        pos[a, i] = index.get(u, -1)
                per_vertex.append({"n": n, "adj": red, "src": src, "m": m, "pos": pos})
            self.levels.append(per_vertex)

    def _build_native(self, adj, feat, n_levels, n_depth, kind, max_field):
        import ctypes

        from . import _lib

        lib = _lib.load()
        V, F = feat.shape
        a32 = np.ascontiguousarray(adj, np.int32)
        f64 = np.ascontiguousarray(feat, np.float64)
        h = ctypes.c_void_p()
        rc = lib.ccn_graph_tables_create(a32.ctypes.data, f64.ctypes.data, V, F, n_levels, 0 if n_depth is None else n_depth,
                                         {"beta": 0, "omega": 1, "omega_wl": 2}[kind], 0 if max_field is None else max_field,
                                         ctypes.byref(h))
        if rc != 0:
            raise _lib.CCNError("ccn_graph_tables_create failed: %s" % lib.ccn_status_string(rc).decode())
        try:
            width = lib.ccn_graph_tables_feature_width(h)
            self.features = np.ctypeslib.as_array(lib.ccn_graph_tables_features(h), shape=(V, width)).copy()
            rk = lib.ccn_graph_tables_rank(h)
            self.rank = np.ctypeslib.as_array(rk, shape=(V,)).astype(np.int64) if rk else None
            mem = ctypes.POINTER(ctypes.c_int32)()
            self.phi = []
            for l in range(n_levels + 1):
                cur = []
                for v in range(V):
                    n = lib.ccn_graph_tables_field(h, l, v, ctypes.byref(mem))
                    cur.append([int(mem[i]) for i in range(n)])
                self.phi.append(cur)
            pa, ps, pm, pp = (ctypes.POINTER(ctypes.c_float)(), ctypes.POINTER(ctypes.c_int32)(), ctypes.POINTER(ctypes.c_int32)(),
                              ctypes.POINTER(ctypes.c_int32)())
            self.levels = []
            for l in range(1, n_levels + 1):
                per_vertex = []
                for v in range(V):
                    n = lib.ccn_graph_tables_vertex(h, l, v, ctypes.byref(pa), ctypes.byref(ps), ctypes.byref(pm), ctypes.byref(pp))
                    per_vertex.append({"n": n, "adj": np.ctypeslib.as_array(pa, shape=(n, n)).copy(),
                                       "src": [int(ps[i]) for i in range(n)], "m": [int(pm[i]) for i in range(n)],
                                       "pos": np.ctypeslib.as_array(pp, shape=(n, n)).copy()})
                self.levels.append(per_vertex)
        finally:
            lib.ccn_graph_tables_destroy(h)
