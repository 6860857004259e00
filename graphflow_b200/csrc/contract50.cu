// contract50.cu -- StackTensor3D + RisiContraction_50 forward / backward (all 50 order-5 -> order-2 contractions of
// T[a,b,c,f] * adj[d,e]; BASELINE.json config 5), any n, any C; sm_100a.
//
// Replaces GraphFlow/RisiContraction_50.h:73-441 (forward, an N^6 loop nest per channel) and :443-802 (backward).
// Every case factorises into (a reduction / diagonal of T) x (A, its row sums r, column sums cs, diagonal dg, total
// sA or trace tr) -- SURVEY.md Appendix A -- so the N^6 nest becomes three N^3 C sweeps of T (15 N^2 C planes) plus
// 18 plane x A products of N^3 C each.  The per-case plan is the generated table r50_table.inc
// (gen/gen_r50_table.py, validated against numpy einsum and the compiled reference).
//
//   forward   k_r50_adj -> k_r50_fwd_planes -> k_r50_fwd_vectors -> k_r50_fwd_out
//   backward  k_r50_adj -> k_r50_bwd_vectors -> k_r50_bwd_planes -> k_r50_bwd_scatter
//
// One thread per plane / output element, coalesced over the channel index (C = 128 floats = 512 contiguous bytes per
// cell at the config-5 shape).  HBM roofline: 4 (N^3 C + N^2 + 50 N^2 C) bytes per direction per instance.
#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int kThreads = 256;
constexpr int kCases = 50;
constexpr int kPlanes = 15, kVecs = 6, kScals = 5;

struct R50Case {
    unsigned char form, id, aux, flags;
};
__constant__ R50Case c_plan[kCases] = {
#include "r50_table.inc"
};

// adjacency table per instance: A[nm*nm] (row stride n), r[nm], cs[nm], dg[nm], {sA, tr, 0, 0}
struct R50Adj {
    int nm;
    __host__ __device__ int r() const { return nm * nm; }
    __host__ __device__ int cs() const { return r() + nm; }
    __host__ __device__ int dg() const { return cs() + nm; }
    __host__ __device__ int scal() const { return dg() + nm; }
    __host__ __device__ int words() const { return (scal() + 4 + 3) & ~3; }
};

struct R50Scratch {  // planes [15][nm*nm*C], vectors [6][nm*C], scalars [5][C]
    int64_t plane, vec, vecs_off, scal_off, words;
    __host__ __device__ R50Scratch(int nm, int C) {
        plane = (int64_t)nm * nm * C;
        vec = (int64_t)nm * C;
        vecs_off = kPlanes * plane;
        scal_off = vecs_off + kVecs * vec;
        words = (scal_off + kScals * C + 3) & ~(int64_t)3;
    }
};

__device__ __forceinline__ const float *slab_of(const TensorRef &t, int inst, int a, int n, int n_max, int C) {
    return t.slabs ? t.slabs[(int64_t)inst * n_max + a] : t.base + inst * t.stride + (int64_t)a * n * n * C;
}

__global__ void __launch_bounds__(128) k_r50_adj(const float *__restrict__ adj, int64_t stride_adj, Batch b, int positive_part,
                                                 float *__restrict__ tab, int words) {
    const int inst = blockIdx.x, n = b.n_of(inst);
    const R50Adj L{b.n_max};
    const float *A = adj + inst * stride_adj;
    float *t = tab + (int64_t)inst * words;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        float v = A[i];
        if (positive_part && !(v > 0.f)) v = 0.f;
        t[i] = v;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        float rs = 0.f, cs = 0.f;
        for (int e = 0; e < n; ++e) {
            rs += t[d * n + e];
            cs += t[e * n + d];
        }
        t[L.r() + d] = rs;
        t[L.cs() + d] = cs;
        t[L.dg() + d] = t[d * n + d];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sA = 0.f, tr = 0.f;
        for (int d = 0; d < n; ++d) {
            sA += t[L.r() + d];
            tr += t[L.dg() + d];
        }
        t[L.scal()] = sA;
        t[L.scal() + 1] = tr;
    }
}

struct R50Args {
    TensorRef T;        // forward: input; backward: gT destination
    float *out;         // forward: out; backward: gout (read only)
    int64_t stride_out;
    Batch b;
    const float *adjtab;
    int adjtab_words;
    float *scratch;
    int64_t scratch_words;
    float beta;
};

#define R50_PQF()                                                         \
    const int inst = blockIdx.y;                                          \
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;              \
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;     \
    if (idx >= (int64_t)n * n * C) return;                                \
    const int f = (int)(idx % C);                                         \
    const int q = (int)((idx / C) % n);                                   \
    const int p = (int)(idx / ((int64_t)C * n));                          \
    const R50Adj AL{nm};                                                  \
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;         \
    const R50Scratch S(nm, C);                                            \
    float *sc = a.scratch + inst * a.scratch_words;

// ---- forward ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_r50_fwd_planes(R50Args a) {
    R50_PQF();
    const float *r = tab + AL.r(), *cs = tab + AL.cs(), *dg = tab + AL.dg();
    const int64_t cell = C, row = (int64_t)n * C;
    {  // (a, b) = (p, q): reduce over c
        const float *t = slab_of(a.T, inst, p, n, nm, C) + q * row + f;
        float s = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;
        for (int c = 0; c < n; ++c) {
            const float v = t[c * cell];
            s += v;
            w0 = fmaf(v, r[c], w0);
            w1 = fmaf(v, cs[c], w1);
            w2 = fmaf(v, dg[c], w2);
        }
        sc[0 * S.plane + idx] = s;
        sc[3 * S.plane + idx] = w0;
        sc[4 * S.plane + idx] = w1;
        sc[5 * S.plane + idx] = w2;
        sc[13 * S.plane + idx] = t[p * cell];  // T[a,b,a]
        sc[14 * S.plane + idx] = t[q * cell];  // T[a,b,b]
    }
    {  // (a, c) = (p, q): reduce over b
        const float *t = slab_of(a.T, inst, p, n, nm, C) + q * cell + f;
        float s = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;
        for (int bb = 0; bb < n; ++bb) {
            const float v = t[bb * row];
            s += v;
            w0 = fmaf(v, r[bb], w0);
            w1 = fmaf(v, cs[bb], w1);
            w2 = fmaf(v, dg[bb], w2);
        }
        sc[1 * S.plane + idx] = s;
        sc[6 * S.plane + idx] = w0;
        sc[7 * S.plane + idx] = w1;
        sc[8 * S.plane + idx] = w2;
        sc[12 * S.plane + idx] = t[p * row];  // T[a,a,c]
    }
    {  // (b, c) = (p, q): reduce over a
        float s = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;
        for (int aa = 0; aa < n; ++aa) {
            const float v = slab_of(a.T, inst, aa, n, nm, C)[p * row + q * cell + f];
            s += v;
            w0 = fmaf(v, r[aa], w0);
            w1 = fmaf(v, cs[aa], w1);
            w2 = fmaf(v, dg[aa], w2);
        }
        sc[2 * S.plane + idx] = s;
        sc[9 * S.plane + idx] = w0;
        sc[10 * S.plane + idx] = w1;
        sc[11 * S.plane + idx] = w2;
    }
}

__global__ void __launch_bounds__(kThreads) k_r50_fwd_vectors(R50Args a) {
    const int inst = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    float *V = sc + S.vecs_off, *X = sc + S.scal_off;
    const int64_t row = (int64_t)n * C;
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int x = i / C, f = i % C;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f;
        for (int j = 0; j < n; ++j) {
            v0 += sc[0 * S.plane + x * row + j * C + f];   // Sa[a]  = sum_b Pab[a,b]
            v1 += sc[0 * S.plane + j * row + x * C + f];   // Sb[b]  = sum_a Pab[a,b]
            v2 += sc[1 * S.plane + j * row + x * C + f];   // Sc[c]  = sum_a Pac[a,c]
            v3 += sc[14 * S.plane + x * row + j * C + f];  // sum_b T[a,b,b]
            v4 += sc[13 * S.plane + j * row + x * C + f];  // sum_a T[a,b,a]
            v5 += sc[12 * S.plane + j * row + x * C + f];  // sum_a T[a,a,c]
        }
        V[0 * S.vec + i] = v0;
        V[1 * S.vec + i] = v1;
        V[2 * S.vec + i] = v2;
        V[3 * S.vec + i] = v3;
        V[4 * S.vec + i] = v4;
        V[5 * S.vec + i] = v5;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f, x4 = 0.f;
        for (int j = 0; j < n; ++j) {
            x0 += V[0 * S.vec + j * C + f];
            x1 += V[5 * S.vec + j * C + f];
            x2 += V[4 * S.vec + j * C + f];
            x3 += V[3 * S.vec + j * C + f];
            x4 += sc[12 * S.plane + j * row + j * C + f];  // T[a,a,a]
        }
        X[0 * C + f] = x0;
        X[1 * C + f] = x1;
        X[2 * C + f] = x2;
        X[3 * C + f] = x3;
        X[4 * C + f] = x4;
    }
}

__global__ void __launch_bounds__(kThreads) k_r50_fwd_out(R50Args a) {
    R50_PQF();
    const float *A = tab;
    const float *V = sc + S.vecs_off, *X = sc + S.scal_off;
    const int64_t row = (int64_t)n * C;
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    float *o = a.out + inst * a.stride_out + ((int64_t)p * n + q) * ((int64_t)kCases * C) + f;
#pragma unroll 1
    for (int k = 0; k < kCases; ++k) {
        const R50Case cs = c_plan[k];
        float v;
        if (cs.form == 0) {
            v = scal[cs.aux] * sc[cs.id * S.plane + idx];
        } else if (cs.form == 1) {
            v = V[cs.id * S.vec + p * C + f] * tab[(cs.aux ? AL.cs() : AL.r()) + q];
        } else if (cs.form == 2) {
            const float *pl = sc + cs.id * S.plane + f;
            const int64_t ps = (cs.flags & 1) ? row : (int64_t)C;      // stride of j in the plane
            pl += (cs.flags & 1) ? (int64_t)p * C : (int64_t)p * row;  // fixed coordinate x = p
            const float *Ar = (cs.flags & 2) ? A + q : A + q * n;      // A[j,y] : A[y,j]
            const int as = (cs.flags & 2) ? n : 1;
            v = 0.f;
            for (int j = 0; j < n; ++j) v = fmaf(pl[j * ps], Ar[j * as], v);
        } else {
            v = X[cs.id * C + f] * A[p * n + q];
        }
        o[(int64_t)k * C] = v;
    }
}

// ---- backward ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float g50(const float *g, int n, int C, int x, int y, int k, int f) {
    return g[((int64_t)x * n + y) * ((int64_t)kCases * C) + (int64_t)k * C + f];
}

__global__ void __launch_bounds__(kThreads) k_r50_bwd_vectors(R50Args a) {
    const int inst = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    float *gV = sc + S.vecs_off, *gX = sc + S.scal_off;
    const float *g = a.out + inst * a.stride_out;
    for (int i = threadIdx.x; i < n * C; i += blockDim.x) {
        const int x = i / C, f = i % C;
        float acc[kVecs] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int k = 0; k < kCases; ++k) {
            const R50Case cs = c_plan[k];
            if (cs.form != 1) continue;
            const float *w = tab + (cs.aux ? AL.cs() : AL.r());
            float s = 0.f;
            for (int y = 0; y < n; ++y) s = fmaf(w[y], g50(g, n, C, x, y, k, f), s);
#pragma unroll
            for (int v = 0; v < kVecs; ++v)
                if (cs.id == v) acc[v] += s;
        }
#pragma unroll
        for (int v = 0; v < kVecs; ++v) gV[v * S.vec + i] = acc[v];
    }
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        float acc[kScals] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int k = 0; k < kCases; ++k) {
            const R50Case cs = c_plan[k];
            if (cs.form != 3) continue;
            float s = 0.f;
            for (int x = 0; x < n; ++x)
                for (int y = 0; y < n; ++y) s = fmaf(tab[x * n + y], g50(g, n, C, x, y, k, f), s);
#pragma unroll
            for (int v = 0; v < kScals; ++v)
                if (cs.id == v) acc[v] += s;
        }
#pragma unroll
        for (int v = 0; v < kScals; ++v) gX[v * C + f] = acc[v];
    }
}

__global__ void __launch_bounds__(kThreads) k_r50_bwd_planes(R50Args a) {
    R50_PQF();
    const float *A = tab;
    const float *gV = sc + S.vecs_off, *gX = sc + S.scal_off;
    const float *g = a.out + inst * a.stride_out;
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
#pragma unroll 1
    for (int pid = 0; pid < kPlanes; ++pid) {
        float acc = 0.f;
#pragma unroll 1
        for (int k = 0; k < kCases; ++k) {
            const R50Case cs = c_plan[k];
            if (cs.id != pid) continue;
            if (cs.form == 0) {
                acc = fmaf(scal[cs.aux], g50(g, n, C, p, q, k, f), acc);
            } else if (cs.form == 2) {
                // forward: out[x,y] = sum_j PL[x,j] Am[y,j]  (plane read as [j,x] when flag bit 0)
                const int x = (cs.flags & 1) ? q : p, j = (cs.flags & 1) ? p : q;
                const float *Ar = (cs.flags & 2) ? A + j * n : A + j;  // Am[y,j] = A[j,y] : A[y,j]
                const int as = (cs.flags & 2) ? 1 : n;
                float s = 0.f;
                for (int y = 0; y < n; ++y) s = fmaf(g50(g, n, C, x, y, k, f), Ar[y * as], s);
                acc += s;
            }
        }
        // fold the vector and scalar gradients into the planes they were reduced from
        if (pid == 0) acc += gV[0 * S.vec + p * C + f] + gV[1 * S.vec + q * C + f] + gX[0 * C + f];
        if (pid == 1) acc += gV[2 * S.vec + q * C + f];
        if (pid == 12) acc += gV[5 * S.vec + q * C + f] + gX[1 * C + f] + (p == q ? gX[4 * C + f] : 0.f);
        if (pid == 13) acc += gV[4 * S.vec + q * C + f] + gX[2 * C + f];
        if (pid == 14) acc += gV[3 * S.vec + p * C + f] + gX[3 * C + f];
        sc[pid * S.plane + idx] = acc;
    }
}

__global__ void __launch_bounds__(kThreads) k_r50_bwd_scatter(R50Args a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t slab = (int64_t)n * n * C;
    if (idx >= slab * n) return;
    const int s = (int)(idx / slab);  // a
    const int64_t rem = idx - (int64_t)s * slab;
    const int f = (int)(rem % C);
    const int c = (int)((rem / C) % n);
    const int bb = (int)(rem / ((int64_t)C * n));
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const float *r = tab + AL.r(), *cs = tab + AL.cs(), *dg = tab + AL.dg();
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const int64_t ab = ((int64_t)s * n + bb) * C + f, ac = ((int64_t)s * n + c) * C + f, bc = ((int64_t)bb * n + c) * C + f;
    float v = sc[0 * S.plane + ab] + sc[1 * S.plane + ac] + sc[2 * S.plane + bc];
    v = fmaf(sc[3 * S.plane + ab], r[c], v);
    v = fmaf(sc[4 * S.plane + ab], cs[c], v);
    v = fmaf(sc[5 * S.plane + ab], dg[c], v);
    v = fmaf(sc[6 * S.plane + ac], r[bb], v);
    v = fmaf(sc[7 * S.plane + ac], cs[bb], v);
    v = fmaf(sc[8 * S.plane + ac], dg[bb], v);
    v = fmaf(sc[9 * S.plane + bc], r[s], v);
    v = fmaf(sc[10 * S.plane + bc], cs[s], v);
    v = fmaf(sc[11 * S.plane + bc], dg[s], v);
    if (s == bb) v += sc[12 * S.plane + ac];
    if (s == c) v += sc[13 * S.plane + ab];
    if (bb == c) v += sc[14 * S.plane + ab];
    float *dst = (a.T.slabs ? a.T.slabs[(int64_t)inst * nm + s] : a.T.base + inst * a.T.stride + (int64_t)s * slab) + rem;
    *dst = (a.beta != 0.f) ? fmaf(a.beta, *dst, v) : v;
}

inline unsigned blocks_for(int64_t elems) { return (unsigned)((elems + kThreads - 1) / kThreads); }

}  // namespace

int r50_adj_words(int n_max) { return R50Adj{n_max}.words(); }
int64_t r50_scratch_words(int n_max, int C) { return R50Scratch(n_max, C).words; }

cudaError_t launch_r50(bool backward, TensorRef T, float *out, int64_t stride_out, const float *adj, int64_t stride_adj,
                       Batch b, int adj_mode, float *adjtab, float *scratch, float beta, cudaStream_t st, LaunchLog *log) {
    const R50Adj AL{b.n_max};
    CCN_LAUNCH(log, K_R50_ADJ, st,
               k_r50_adj<<<b.count, 128, 0, st>>>(adj, stride_adj, b, adj_mode == 0 ? 1 : 0, adjtab, AL.words()));
    R50Args a;
    a.T = T;
    a.out = out;
    a.stride_out = stride_out;
    a.b = b;
    a.adjtab = adjtab;
    a.adjtab_words = AL.words();
    a.scratch = scratch;
    a.scratch_words = R50Scratch(b.n_max, b.C).words;
    a.beta = beta;
    const int64_t plane = (int64_t)b.n_max * b.n_max * b.C;
    dim3 grid(blocks_for(plane), b.count), grid3(blocks_for(plane * b.n_max), b.count);
    if (!backward) {
        CCN_LAUNCH(log, K_R50_FWD_PLANES, st, k_r50_fwd_planes<<<grid, kThreads, 0, st>>>(a));
        CCN_LAUNCH(log, K_R50_FWD_VECTORS, st, k_r50_fwd_vectors<<<b.count, kThreads, 0, st>>>(a));
        CCN_LAUNCH(log, K_R50_FWD_OUT, st, k_r50_fwd_out<<<grid, kThreads, 0, st>>>(a));
    } else {
        CCN_LAUNCH(log, K_R50_BWD_VECTORS, st, k_r50_bwd_vectors<<<b.count, kThreads, 0, st>>>(a));
        CCN_LAUNCH(log, K_R50_BWD_PLANES, st, k_r50_bwd_planes<<<grid, kThreads, 0, st>>>(a));
        CCN_LAUNCH(log, K_R50_BWD_SCATTER, st, k_r50_bwd_scatter<<<grid3, kThreads, 0, st>>>(a));
    }
    return cudaGetLastError();
}

}  // namespace ccn
