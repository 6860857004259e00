// graph_tables.cu -- host-side graph preprocessing of the second-order CCN models -> flat index tables for the device
// kernels (pure host C++; lives in a .cu file only so that the one nvcc command of graphflow_b200/build.py picks it up).
//
// What SMP_beta::complete_computation_graph derives from a DenseGraph before it wires any operator
// (GraphFlow/SMP_beta.h:531-552): shortest paths (floyd_warshall, :343-365), Weisfeiler-Lehman histogram features
// (weisfeiler_lehman, :367-389), the vertex ranking (rank_vertices, :403-419: an exchange sort whose tie behaviour is
// reproduced), the receptive fields phi_l(v) (:461-489, ordered by rank, :435-444) and the reduced adjacency matrices
// (:505-526); for SMP_omega_physics the insertion-ordered fields cut by limit_receptive_field
// (SMP_omega_physics.h:367-418); for SMP_omega the WL features and ranking of SMP_beta with fields cut by distance then rank
// and ordered by rank (SMP_omega.h:476-531).  The 0/1 selection matrices X[v][w] (init_permutation_matrix, SMP_beta.h:446-459) are
// never materialised: they become the `pos` table of ccn_promote_forward.  The reference rebuilds all of this three times
// per example per BatchLearn (SMP_beta.h:753,758,770); here it is built once per graph.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/ccn_b200.h"

namespace {
const int64_t kInf = 1000000000;  // SMP_beta::INF (SMP_beta.h:1011)

struct VertexTable {
    int n;
    std::vector<float> adj;    // [n, n] reduced adjacency, 1 on the diagonal (SMP_beta.h:516-520)
    std::vector<int32_t> src;  // [n]    w = phi_l(v)[a]
    std::vector<int32_t> m;    // [n]    |phi_{l-1}(w)|
    std::vector<int32_t> pos;  // [n, n] position of phi_l(v)[i] inside phi_{l-1}(w), or -1
};
}  // namespace

struct ccn_graph_tables {
    int V, F, L, width;
    std::vector<double> features;                          // [V, width]
    std::vector<int32_t> rank;                             // [V] (beta) or empty (omega)
    std::vector<std::vector<std::vector<int32_t>>> phi;    // [L+1][V][n]
    std::vector<std::vector<VertexTable>> levels;          // [L][V]
};

namespace {

std::vector<int64_t> shortest_paths(const int32_t *adj, int V) {
    std::vector<int64_t> sp((size_t)V * V, kInf);
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j)
            if (adj[i * V + j] > 0) sp[(size_t)i * V + j] = sp[(size_t)j * V + i] = 1;
    for (int i = 0; i < V; ++i) sp[(size_t)i * V + i] = 0;
    for (int k = 0; k < V; ++k)
        for (int i = 0; i < V; ++i) {
            const int64_t ik = sp[(size_t)i * V + k];
            for (int j = 0; j < V; ++j) sp[(size_t)i * V + j] = std::min(sp[(size_t)i * V + j], ik + sp[(size_t)k * V + j]);
        }
    return sp;
}

// compare_vertices (SMP_beta.h:391-401): lexicographic order of two histogram rows
int compare_rows(const double *a, const double *b, int w) {
    for (int i = 0; i < w; ++i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}

void append_unique(std::vector<int32_t> &dst, const std::vector<int32_t> &src) {
    for (size_t i = 0; i < src.size(); ++i)
        if (std::find(dst.begin(), dst.end(), src[i]) == dst.end()) dst.push_back(src[i]);
}

// limit_receptive_field (SMP_omega_physics.h:367-392): the reference's exchange sort by distance from v (not stable:
// reproduced literally), then whole outermost distance shells are dropped until the field fits
void limit_field(const std::vector<int64_t> &sp, int V, int v, std::vector<int32_t> &A, int max_field) {
    for (size_t i = 0; i < A.size(); ++i)
        for (size_t j = i + 1; j < A.size(); ++j)
            if (sp[(size_t)v * V + A[i]] > sp[(size_t)v * V + A[j]]) std::swap(A[i], A[j]);
    while ((int)A.size() > max_field) {
        const int64_t d = sp[(size_t)v * V + A.back()];
        while (!A.empty() && sp[(size_t)v * V + A.back()] == d) A.pop_back();
    }
}

}  // namespace

extern "C" {

int ccn_graph_tables_create(const int32_t *adj, const double *feat, int V, int F, int n_levels, int n_depth, int kind, int max_field,
                            ccn_graph_tables **out) {
    if (!adj || !feat || !out || V <= 0 || F <= 0 || n_levels < 0 ||
        (kind != CCN_GRAPH_BETA && kind != CCN_GRAPH_OMEGA && kind != CCN_GRAPH_OMEGA_WL) || (kind != CCN_GRAPH_OMEGA && n_depth < 0))
        return CCN_ERR_INVALID_ARGUMENT;
    ccn_graph_tables *g = new (std::nothrow) ccn_graph_tables();
    if (!g) return CCN_ERR_OUT_OF_MEMORY;
    g->V = V;
    g->F = F;
    g->L = n_levels;
    const std::vector<int64_t> sp = shortest_paths(adj, V);
    if (kind == CCN_GRAPH_OMEGA) {
        g->width = F;
        g->features.assign(feat, feat + (size_t)V * F);
    } else {
        // weisfeiler_lehman (SMP_beta.h:367-389): hist[v, d*F + f] = sum of feat[u, f] over u at distance d from v
        g->width = F * (n_depth + 1);
        g->features.assign((size_t)V * g->width, 0.0);
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < V; ++u) {
                const int64_t d = sp[(size_t)u * V + v];
                if (d <= n_depth)
                    for (int f = 0; f < F; ++f) g->features[(size_t)v * g->width + d * F + f] += feat[(size_t)u * F + f];
            }
        // rank_vertices (SMP_beta.h:403-419): the reference's exchange sort, ties resolve the way it does
        std::vector<int32_t> order(V);
        for (int i = 0; i < V; ++i) order[i] = i;
        for (int i = 0; i < V; ++i)
            for (int j = i + 1; j < V; ++j)
                if (compare_rows(&g->features[(size_t)order[i] * g->width], &g->features[(size_t)order[j] * g->width], g->width) < 0)
                    std::swap(order[i], order[j]);
        g->rank.assign(V, 0);
        for (int i = 0; i < V; ++i) g->rank[order[i]] = i;
    }
    // receptive fields: phi_0(v) = {v}; phi_l(v) = union of phi_{l-1}(u) over u within distance 1 of v
    g->phi.assign(n_levels + 1, std::vector<std::vector<int32_t>>(V));
    for (int v = 0; v < V; ++v) g->phi[0][v].assign(1, v);
    const int cap = max_field > 0 ? max_field : V;
    for (int l = 1; l <= n_levels; ++l)
        for (int v = 0; v < V; ++v) {
            std::vector<int32_t> &mem = g->phi[l][v];
            for (int u = 0; u < V; ++u)
                if (sp[(size_t)u * V + v] <= 1) append_unique(mem, g->phi[l - 1][u]);
            if (kind == CCN_GRAPH_BETA) {  // ordered by rank (distinct), SMP_beta.h:435-444
                const std::vector<int32_t> &rk = g->rank;
                std::stable_sort(mem.begin(), mem.end(), [&rk](int32_t a, int32_t b) { return rk[a] < rk[b]; });
            } else if (kind == CCN_GRAPH_OMEGA_WL) {
                // SMP_omega.h:512-530: cut the union to max_field members FIRST (limit_receptive_field, :476-510: ordered by
                // distance from v, ties by rank; (distance, rank) is a total order, so the reference's exchange sort is a plain
                // sort; whole outermost shells are dropped), THEN order what is left by rank (sort, :451-459)
                const std::vector<int32_t> &rk = g->rank;
                if ((int)mem.size() > cap) {
                    std::sort(mem.begin(), mem.end(), [&](int32_t a, int32_t b) {
                        const int64_t da = sp[(size_t)v * V + a], db = sp[(size_t)v * V + b];
                        return da != db ? da < db : rk[a] < rk[b];
                    });
                    while ((int)mem.size() > cap) {
                        const int64_t d = sp[(size_t)v * V + mem.back()];
                        while (!mem.empty() && sp[(size_t)v * V + mem.back()] == d) mem.pop_back();
                    }
                }
                std::stable_sort(mem.begin(), mem.end(), [&rk](int32_t a, int32_t b) { return rk[a] < rk[b]; });
            } else if ((int)mem.size() > cap) {
                limit_field(sp, V, v, mem, cap);
            }
        }
    // per level l >= 1 and vertex v: reduced adjacency and the promotion index tables
    g->levels.assign(n_levels, std::vector<VertexTable>(V));
    std::vector<int32_t> index(V);
    for (int l = 1; l <= n_levels; ++l)
        for (int v = 0; v < V; ++v) {
            const std::vector<int32_t> &field = g->phi[l][v];
            VertexTable &t = g->levels[l - 1][v];
            const int n = (int)field.size();
            t.n = n;
            t.adj.resize((size_t)n * n);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j)
                    t.adj[(size_t)i * n + j] = field[i] == field[j] ? 1.0f : (float)adj[(size_t)field[i] * V + field[j]];
            t.src.resize(n);
            t.m.resize(n);
            t.pos.assign((size_t)n * n, -1);
            for (int a = 0; a < n; ++a) {
                const int w = field[a];
                const std::vector<int32_t> &prev = g->phi[l - 1][w];
                std::fill(index.begin(), index.end(), -1);
                for (size_t k = 0; k < prev.size(); ++k) index[prev[k]] = (int32_t)k;
                t.src[a] = w;
                t.m[a] = (int32_t)prev.size();
                for (int i = 0; i < n; ++i) t.pos[(size_t)a * n + i] = index[field[i]];
            }
        }
    *out = g;
    return CCN_OK;
}

void ccn_graph_tables_destroy(ccn_graph_tables *g) { delete g; }

int ccn_graph_tables_feature_width(const ccn_graph_tables *g) { return g ? g->width : 0; }

const double *ccn_graph_tables_features(const ccn_graph_tables *g) { return g ? g->features.data() : nullptr; }

const int32_t *ccn_graph_tables_rank(const ccn_graph_tables *g) { return (g && !g->rank.empty()) ? g->rank.data() : nullptr; }

int ccn_graph_tables_field(const ccn_graph_tables *g, int level, int v, const int32_t **members) {
    if (!g || level < 0 || level > g->L || v < 0 || v >= g->V) return -1;
    if (members) *members = g->phi[level][v].data();
    return (int)g->phi[level][v].size();
}

int ccn_graph_tables_vertex(const ccn_graph_tables *g, int level, int v, const float **adj_red, const int32_t **src, const int32_t **m,
                            const int32_t **pos) {
    if (!g || level < 1 || level > g->L || v < 0 || v >= g->V) return -1;
    const VertexTable &t = g->levels[level - 1][v];
    if (adj_red) *adj_red = t.adj.data();
    if (src) *src = t.src.data();
    if (m) *m = t.m.data();
    if (pos) *pos = t.pos.data();
    return t.n;
}

}  // extern "C"
