// contract50.cu -- StackTensor3D + RisiContraction_50 forward / backward (all 50 order-5 -> order-2 contractions of
// T[a,b,c,f] * adj[d,e]; BASELINE.json config 5), any n, any C; sm_100a.
//
// Replaces GraphFlow/RisiContraction_50.h:73-441 (forward, an N^6 loop nest per channel) and :443-802 (backward).
// Every case factorises into (a reduction / diagonal of T) x (A, its row sums r, column sums cs, diagonal dg, total
// sA or trace tr) -- SURVEY.md Appendix A -- so the N^6 nest becomes three N^3 C sweeps of T (15 N^2 C planes) plus
// 18 plane x A products of N^3 C each.  The per-case plan is the generated table r50_table.inc
// (gen/gen_r50_table.py, validated against numpy einsum and the compiled reference).
//
//   forward   k_r50_adj -> k_r50_fwd_planes[_plain] -> k_r50_fwd_vectors -> k_r50_fwd_out_v4
//   backward  k_r50_adj -> k_r50_bwd_vectors -> k_r50_bwd_planes_v4<0>, <1> -> k_r50_bwd_scatter_v4 / _plain
//
// Three generations of the heavy stages live here: the "_v4" vector kernels (C % 4 == 0, n <= 48, 16-byte aligned operands: four
// channels per thread, a whole row x per CTA, one cp.async staging round, packed adjacency lists) are what the BASELINE shapes
// run; the "_tiled" kernels are the general path (any C, n up to the shared-memory limit); the one-thread-per-element kernels
// take everything else.  Plans without adjacency-weighted planes (RisiContraction_4 / _10) materialise only the planes they
// reference (R50Args::pm).  HBM roofline: 4 (N^3 C + N^2 + 50 N^2 C) bytes per direction per instance.
#include <cstdlib>

#include "contract18_kernels.cuh"

namespace ccn {

namespace {

constexpr int kThreads = 256;
constexpr int kR50VwFwd = 2, kR50VwBwd = 2;  // channels per thread of the tiled plane x A kernels (measured)
constexpr int kCases = kR50MaxCases;
constexpr int kPlanes = 15, kVecs = 6, kScals = 5;
constexpr int kAdjL = 16;  // list slots per row / column of A in the adjacency table (longer rows: dense kernels)

// forms of a plan entry (R50Case is declared in contract18_kernels.cuh):
//   0  s * PL[x,y]          1  V[x] * w[y]          2  sum_j PLv[x,j] Am[y,j]          3  X * A[x,y]
//   kFormFollower: a form-2 case that the tiled forward writes together with its leader (flags bits 2..7 of the leader
//                  hold the follower's index + 1); everywhere else it is an ordinary form-2 case
//   kFormOff:      a dropped slab (RisiContraction_18_dropout): forward writes zeros, backward ignores it
constexpr int kFormFollower = 12, kFormOff = 9;
__host__ __device__ inline bool r50_is2(int form) { return form == 2 || form == kFormFollower; }

const R50Case h_master[kCases] = {
#include "r50_table.inc"
};

// adjacency table per instance: A[nm*nm] (row stride n), r[nm], cs[nm], dg[nm], {sA, tr, 0, 0}
struct R50Adj {
    int nm;
    __host__ __device__ int r() const { return nm * nm; }
    __host__ __device__ int cs() const { return r() + nm; }
    __host__ __device__ int dg() const { return cs() + nm; }
    __host__ __device__ int scal() const { return dg() + nm; }       // {sA, tr, max list length (int bits), 0}
    // packed non-zeros for the vector kernels: rowl[y][kAdjL] = {j * 8, A[y][j]}, coll[y][kAdjL] = {j * 8, A[j][y]} (int2, the
    // first min(count, kAdjL) entries valid), then the counts cntr[nm], cntc[nm]
    __host__ __device__ int rowl() const { return (scal() + 4 + 3) & ~3; }
    __host__ __device__ int coll() const { return rowl() + 2 * nm * kAdjL; }
    __host__ __device__ int cnt() const { return coll() + 2 * nm * kAdjL; }
    __host__ __device__ int words() const { return (cnt() + 2 * nm + 3) & ~3; }
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
                                                 float scale, float *__restrict__ tab, int words) {
    const int inst = blockIdx.x, n = b.n_of(inst);
    const R50Adj L{b.n_max};
    const float *A = adj + inst * stride_adj;
    float *t = tab + (int64_t)inst * words;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        float v = adj ? A[i] : 0.f;  // no adjacency operand (RisiContraction_4): the weighted planes are unused
        if (positive_part && !(v > 0.f)) v = 0.f;
        t[i] = v * scale;  // every case is linear in the adjacency, so this scales the whole output
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
    __shared__ int maxcnt;
    if (threadIdx.x == 0) {
        float sA = 0.f, tr = 0.f;
        for (int d = 0; d < n; ++d) {
            sA += t[L.r() + d];
            tr += t[L.dg() + d];
        }
        t[L.scal()] = sA;
        t[L.scal() + 1] = tr;
        maxcnt = 0;
    }
    __syncthreads();
    int2 *rowl = reinterpret_cast<int2 *>(t + L.rowl()), *coll = reinterpret_cast<int2 *>(t + L.coll());
    int *cnt = reinterpret_cast<int *>(t + L.cnt());
    for (int d = threadIdx.x; d < 2 * n; d += blockDim.x) {
        const bool isrow = d < n;
        const int y = isrow ? d : d - n;
        int2 *dst = (isrow ? rowl : coll) + y * kAdjL;
        int c = 0;
        for (int j = 0; j < n; ++j) {
            const float v = isrow ? t[y * n + j] : t[j * n + y];
            if (v != 0.f) {
                if (c < kAdjL) dst[c] = make_int2(j * 8, __float_as_int(v));
                ++c;
            }
        }
        cnt[(isrow ? 0 : b.n_max) + y] = c;
        atomicMax(&maxcnt, c);
    }
    __syncthreads();
    if (threadIdx.x == 0) t[L.scal() + 2] = __int_as_float(maxcnt);
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
    int ncases;               // slabs of out / gout
    R50Case plan[kCases];     // by value: lives in the kernel parameter (constant) bank
    // the vector kernels' view of the plan: the (plane, orientation) tiles of the form-2 cases with the case that takes A's row
    // list (A[y,j]) and the one that takes its column list (A[j,y]) (0xff = none), and the form-0 cases
    unsigned char tile_pid[9], tile_or[9], tile_krow[9], tile_kcol[9], z_k[15], z_pid[15], z_aux[15];
    int ntiles, nz;
    // backward: the tiles of each pass (orientation), and per plane its first / second form-0 case (0xff = none)
    unsigned char pt_pid[2][5], pt_krow[2][5], pt_kcol[2][5], pz0[15], pza0[15], pz1[3], pza1[3];
    int npt[2];
    // the planes this launch materialises in scratch: all 15, or -- for a plan that touches no adjacency-weighted plane (3..11;
    // RisiContraction_4, RisiContraction_10) -- only the plain sums / diagonals {0, 1, 2, 12, 13, 14} it references
    unsigned pm;
};

#define R50_PQF()                                                         \
    const int inst = blockIdx.y;                                          \
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;              \
    const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;     \
    if (idx >= (int64_t)n * n * C) return;                                \
    const uint32_t i32 = (uint32_t)idx; /* n*n*C < 2^31, checked by the C-ABI */ \
    const int f = (int)(i32 % (uint32_t)C);                               \
    const int q = (int)((i32 / (uint32_t)C) % (uint32_t)n);               \
    const int p = (int)(i32 / ((uint32_t)C * (uint32_t)n));               \
    const R50Adj AL{nm};                                                  \
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;         \
    const R50Scratch S(nm, C);                                            \
    float *sc = a.scratch + inst * a.scratch_words;

// ---- forward ---------------------------------------------------------------------------------------------------
// V consecutive channels per thread (V = 4: one 16-byte load / store per cell; needs C % 4 == 0 and 16-byte aligned slabs)
template <int V>
struct R50Vec {
    float x[V];
};
template <int V>
__device__ __forceinline__ R50Vec<V> r50_ld(const float *p) {
    R50Vec<V> r;
    if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        r.x[0] = t.x, r.x[1 % V] = t.y, r.x[2 % V] = t.z, r.x[3 % V] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) r.x[i] = __ldg(p + i);
    }
    return r;
}
template <int V>
__device__ __forceinline__ void r50_st(float *p, const R50Vec<V> &v) {
    if (V == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v.x[0], v.x[1 % V], v.x[2 % V], v.x[3 % V]);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) p[i] = v.x[i];
    }
}

// s = sum_i v_i, w0/w1/w2 = sum_i v_i {r,cs,dg}[i] with v_i = *src(i): eight independent loads are issued before the
// first use (the compiler otherwise interleaves load and use and the in-order warp keeps ~1 load in flight); the three
// weights of an element are loaded once for the V channels of the thread.
template <int V, typename Src>
__device__ __forceinline__ void r50_reduce4(Src src, int n, const float *r, const float *cs, const float *dg, R50Vec<V> &s,
                                            R50Vec<V> &w0, R50Vec<V> &w1, R50Vec<V> &w2) {
#pragma unroll
    for (int k = 0; k < V; ++k) s.x[k] = w0.x[k] = w1.x[k] = w2.x[k] = 0.f;
    for (int i0 = 0; i0 < n; i0 += 8) {
        R50Vec<V> v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = r50_ld<V>(src(min(i0 + u, n - 1)));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u;
            if (i < n) {
                const float a0 = __ldg(r + i), a1 = __ldg(cs + i), a2 = __ldg(dg + i);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    s.x[k] += v[u].x[k];
                    w0.x[k] = fmaf(v[u].x[k], a0, w0.x[k]);
                    w1.x[k] = fmaf(v[u].x[k], a1, w1.x[k]);
                    w2.x[k] = fmaf(v[u].x[k], a2, w2.x[k]);
                }
            }
        }
    }
}

// the unweighted sum only (plans without weighted planes)
template <int V, typename Src>
__device__ __forceinline__ void r50_reduce1(Src src, int n, R50Vec<V> &s) {
#pragma unroll
    for (int k = 0; k < V; ++k) s.x[k] = 0.f;
    for (int i0 = 0; i0 < n; i0 += 8) {
        R50Vec<V> v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = r50_ld<V>(src(min(i0 + u, n - 1)));
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + u < n) {
#pragma unroll
                for (int k = 0; k < V; ++k) s.x[k] += v[u].x[k];
            }
    }
}

template <int V>
__global__ void __launch_bounds__(kThreads) k_r50_fwd_planes_plain(R50Args a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const uint32_t CV = (uint32_t)C / V;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (tid >= (int64_t)n * n * CV) return;
    const uint32_t i32 = (uint32_t)tid;
    const int f = (int)(i32 % CV) * V;
    const int q = (int)((i32 / CV) % (uint32_t)n);
    const int p = (int)(i32 / (CV * (uint32_t)n));
    const int64_t idx = ((int64_t)p * n + q) * C + f;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const int64_t cell = C, row = (int64_t)n * C;
    const unsigned pm = a.pm;
    R50Vec<V> s;
    const float *t = slab_of(a.T, inst, p, n, nm, C);
    if (pm & 1u) {  // (a, b) = (p, q): sum over c
        const float *tr = t + q * row + f;
        r50_reduce1<V>([&](int c) { return tr + c * cell; }, n, s);
        r50_st<V>(sc + 0 * S.plane + idx, s);
    }
    if (pm & (1u << 13)) r50_st<V>(sc + 13 * S.plane + idx, r50_ld<V>(t + q * row + f + p * cell));  // T[a,b,a]
    if (pm & (1u << 14)) r50_st<V>(sc + 14 * S.plane + idx, r50_ld<V>(t + q * row + f + q * cell));  // T[a,b,b]
    if (pm & 2u) {  // (a, c) = (p, q): sum over b
        const float *tc = t + q * cell + f;
        r50_reduce1<V>([&](int bb) { return tc + bb * row; }, n, s);
        r50_st<V>(sc + 1 * S.plane + idx, s);
    }
    if (pm & (1u << 12)) r50_st<V>(sc + 12 * S.plane + idx, r50_ld<V>(t + q * cell + f + p * row));  // T[a,a,c]
    if (pm & 4u) {  // (b, c) = (p, q): sum over a
        const int64_t off = p * row + q * cell + f;
        r50_reduce1<V>([&](int aa) { return slab_of(a.T, inst, aa, n, nm, C) + off; }, n, s);
        r50_st<V>(sc + 2 * S.plane + idx, s);
    }
}

template <int V>
__global__ void __launch_bounds__(kThreads) k_r50_fwd_planes(R50Args a) {
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const uint32_t CV = (uint32_t)C / V;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (tid >= (int64_t)n * n * CV) return;
    const uint32_t i32 = (uint32_t)tid;
    const int f = (int)(i32 % CV) * V;
    const int q = (int)((i32 / CV) % (uint32_t)n);
    const int p = (int)(i32 / (CV * (uint32_t)n));
    const int64_t idx = ((int64_t)p * n + q) * C + f;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const float *r = tab + AL.r(), *cs = tab + AL.cs(), *dg = tab + AL.dg();
    const int64_t cell = C, row = (int64_t)n * C;
    R50Vec<V> s, w0, w1, w2;
    {  // (a, b) = (p, q): reduce over c
        const float *t = slab_of(a.T, inst, p, n, nm, C) + q * row + f;
        r50_reduce4<V>([&](int c) { return t + c * cell; }, n, r, cs, dg, s, w0, w1, w2);
        r50_st<V>(sc + 0 * S.plane + idx, s);
        r50_st<V>(sc + 3 * S.plane + idx, w0);
        r50_st<V>(sc + 4 * S.plane + idx, w1);
        r50_st<V>(sc + 5 * S.plane + idx, w2);
        r50_st<V>(sc + 13 * S.plane + idx, r50_ld<V>(t + p * cell));  // T[a,b,a]
        r50_st<V>(sc + 14 * S.plane + idx, r50_ld<V>(t + q * cell));  // T[a,b,b]
    }
    {  // (a, c) = (p, q): reduce over b
        const float *t = slab_of(a.T, inst, p, n, nm, C) + q * cell + f;
        r50_reduce4<V>([&](int bb) { return t + bb * row; }, n, r, cs, dg, s, w0, w1, w2);
        r50_st<V>(sc + 1 * S.plane + idx, s);
        r50_st<V>(sc + 6 * S.plane + idx, w0);
        r50_st<V>(sc + 7 * S.plane + idx, w1);
        r50_st<V>(sc + 8 * S.plane + idx, w2);
        r50_st<V>(sc + 12 * S.plane + idx, r50_ld<V>(t + p * row));  // T[a,a,c]
    }
    {  // (b, c) = (p, q): reduce over a
        const int64_t off = p * row + q * cell + f;
        r50_reduce4<V>([&](int aa) { return slab_of(a.T, inst, aa, n, nm, C) + off; }, n, r, cs, dg, s, w0, w1, w2);
        r50_st<V>(sc + 2 * S.plane + idx, s);
        r50_st<V>(sc + 9 * S.plane + idx, w0);
        r50_st<V>(sc + 10 * S.plane + idx, w1);
        r50_st<V>(sc + 11 * S.plane + idx, w2);
    }
}

// CTA per (row x, instance): the six vectors at x; the five scalars are accumulated with one atomic per (x, f) into the
// zero-initialised scalar block (order-independent up to fp32 rounding of n terms).
__global__ void __launch_bounds__(kThreads) k_r50_fwd_vectors(R50Args a) {
    const int inst = blockIdx.y, x = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    if (x >= n) return;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    float *V = sc + S.vecs_off, *X = sc + S.scal_off;
    const int64_t row = (int64_t)n * C;
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        const int i = x * C + f;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f;
        // only the planes this launch materialised (a.pm): a plan without diagonal cases skips half of the reads
        const bool h0 = a.pm & 1u, h1 = a.pm & 2u, h12 = a.pm & (1u << 12), h13 = a.pm & (1u << 13), h14 = a.pm & (1u << 14);
        for (int j = 0; j < n; ++j) {
            if (h0) v0 += sc[0 * S.plane + x * row + j * C + f];    // Sa[a]  = sum_b Pab[a,b]
            if (h0) v1 += sc[0 * S.plane + j * row + x * C + f];    // Sb[b]  = sum_a Pab[a,b]
            if (h1) v2 += sc[1 * S.plane + j * row + x * C + f];    // Sc[c]  = sum_a Pac[a,c]
            if (h14) v3 += sc[14 * S.plane + x * row + j * C + f];  // sum_b T[a,b,b]
            if (h13) v4 += sc[13 * S.plane + j * row + x * C + f];  // sum_a T[a,b,a]
            if (h12) v5 += sc[12 * S.plane + j * row + x * C + f];  // sum_a T[a,a,c]
        }
        V[0 * S.vec + i] = v0;
        V[1 * S.vec + i] = v1;
        V[2 * S.vec + i] = v2;
        V[3 * S.vec + i] = v3;
        V[4 * S.vec + i] = v4;
        V[5 * S.vec + i] = v5;
        atomicAdd(X + 0 * C + f, v0);                                    // sum T
        atomicAdd(X + 1 * C + f, v5);                                    // sum T[a,a,c]
        atomicAdd(X + 2 * C + f, v4);                                    // sum T[a,b,a]
        atomicAdd(X + 3 * C + f, v3);                                    // sum T[a,b,b]
        if (h12) atomicAdd(X + 4 * C + f, sc[12 * S.plane + x * row + x * C + f]);  // sum T[a,a,a]
    }
}

// zero the scalar block of every instance's scratch (both directions accumulate into it with atomics)
__global__ void __launch_bounds__(kThreads) k_r50_zero_scalars(R50Args a) {
    const R50Scratch S(a.b.n_max, a.b.C);
    float *X = a.scratch + blockIdx.x * a.scratch_words + S.scal_off;
    for (int i = threadIdx.x; i < kScals * a.b.C; i += blockDim.x) X[i] = 0.f;
}

__global__ void __launch_bounds__(kThreads) k_r50_fwd_out(R50Args a) {
    R50_PQF();
    const float *A = tab;
    const float *V = sc + S.vecs_off, *X = sc + S.scal_off;
    const int64_t row = (int64_t)n * C;
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    float *o = a.out + inst * a.stride_out + ((int64_t)p * n + q) * ((int64_t)a.ncases * C) + f;
#pragma unroll 1
    for (int k = 0; k < a.ncases; ++k) {
        const R50Case cs = a.plan[k];
        float v;
        if (cs.form == 0) {
            v = scal[cs.aux] * sc[cs.id * S.plane + idx];
        } else if (cs.form == 1) {
            v = V[cs.id * S.vec + p * C + f] * tab[(cs.aux ? AL.cs() : AL.r()) + q];
        } else if (cs.form == kFormOff) {
            v = 0.f;
        } else if (r50_is2(cs.form)) {
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
__device__ __forceinline__ float g50(const float *g, int n, int C, int ncases, int x, int y, int k, int f) {
    return g[((int64_t)x * n + y) * ((int64_t)ncases * C) + (int64_t)k * C + f];
}

// CTA per (row x, instance): gV[.][x] from the form-1 cases and row x's share of the form-3 scalars (atomics into the
// zero-initialised scalar block; only the non-zeros of A's row x contribute).
__global__ void __launch_bounds__(kThreads) k_r50_bwd_vectors(R50Args a) {
    const int inst = blockIdx.y, x = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    if (x >= n) return;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    float *gV = sc + S.vecs_off, *gX = sc + S.scal_off;
    const float *g = a.out + inst * a.stride_out;
    for (int f = threadIdx.x; f < C; f += blockDim.x) {
        float acc[kVecs] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float xs[kScals] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int k = 0; k < a.ncases; ++k) {
            const R50Case cs = a.plan[k];
            if (cs.form == 1) {
                const float *w = tab + (cs.aux ? AL.cs() : AL.r());
                float s = 0.f;
                for (int y = 0; y < n; ++y) s = fmaf(w[y], g50(g, n, C, a.ncases, x, y, k, f), s);
#pragma unroll
                for (int v = 0; v < kVecs; ++v)
                    if (cs.id == v) acc[v] += s;
            } else if (cs.form == 3) {
                float s = 0.f;
                for (int y = 0; y < n; ++y) {
                    const float w = tab[x * n + y];
                    if (w != 0.f) s = fmaf(w, g50(g, n, C, a.ncases, x, y, k, f), s);
                }
#pragma unroll
                for (int v = 0; v < kScals; ++v)
                    if (cs.id == v) xs[v] += s;
            }
        }
#pragma unroll
        for (int v = 0; v < kVecs; ++v) gV[v * S.vec + x * C + f] = acc[v];
#pragma unroll
        for (int v = 0; v < kScals; ++v) atomicAdd(gX + v * C + f, xs[v]);
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
        for (int k = 0; k < a.ncases; ++k) {
            const R50Case cs = a.plan[k];
            if (cs.id != pid) continue;
            if (cs.form == 0) {
                acc = fmaf(scal[cs.aux], g50(g, n, C, a.ncases, p, q, k, f), acc);
            } else if (r50_is2(cs.form)) {
                // forward: out[x,y] = sum_j PL[x,j] Am[y,j]  (plane read as [j,x] when flag bit 0)
                const int x = (cs.flags & 1) ? q : p, j = (cs.flags & 1) ? p : q;
                const float *Ar = (cs.flags & 2) ? A + j * n : A + j;  // Am[y,j] = A[j,y] : A[y,j]
                const int as = (cs.flags & 2) ? 1 : n;
                float s = 0.f;
                for (int y = 0; y < n; ++y) s = fmaf(g50(g, n, C, a.ncases, x, y, k, f), Ar[y * as], s);
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

// ---- shared-memory tiled versions of the two plane x A stages ------------------------------------------------------
// CTA = (row x, instance, channel chunk of blockDim.x).  The 18 form-2 cases are [n x n] . [n x n] products per channel;
// the CTA stages row x of the operand (n x CB floats) in shared memory once per case and every thread (one channel)
// produces the n results of its row against A (or A^T) kept in shared memory and read as broadcast float4s -- n times
// fewer L2 loads than one thread per output element.
struct R50Tile {
    float *row;  // [n][CB]
    float *A;    // [n][n4]  A[i][j]
    float *At;   // [n][n4]  A[j][i]
    int n4;
};
__device__ __forceinline__ R50Tile r50_tile(float *smem, int n, int CB) {
    R50Tile t;
    t.n4 = (n + 3) & ~3;
    t.row = smem;
    t.A = smem + (size_t)n * CB;
    t.At = t.A + (size_t)n * t.n4;
    return t;
}
__host__ __device__ inline size_t r50_tile_bytes(int nm, int CB) { return ((size_t)nm * CB + 2 * (size_t)nm * ((nm + 3) & ~3)) * 4; }
__host__ __device__ inline size_t r50_tile2_bytes(int nm, int CB) { return r50_tile_bytes(nm, CB) + (size_t)nm * CB * 4; }

__device__ __forceinline__ void r50_load_adj(const R50Tile &t, const float *A, int n) {
    for (int i = threadIdx.x; i < n * t.n4; i += blockDim.x) {
        const int r = i / t.n4, c = i % t.n4;
        t.A[i] = c < n ? A[r * n + c] : 0.f;
        t.At[i] = c < n ? A[c * n + r] : 0.f;
    }
}

// Sparse form of the adjacency tiles.  When every row and column of A has at most L - 4 = n4/2 - 4 non-zeros (molecular
// graphs: a handful), the dense tiles are replaced IN PLACE by packed lists: for q, col[q*L] = {count} followed by the
// {y * CB, A[y][q]} entries, row[q*L] likewise with {y * CB, A[q][y]}.  `stage` is scratch of 2*n*n4 words (a row tile that is not
// in use yet).  Returns false (block-uniform) and leaves the dense tiles untouched otherwise.
// Requires r50_load_adj + __syncthreads() before the call.
struct R50Lists {
    const int2 *col, *row;
    int L;
};
__device__ __forceinline__ bool r50_build_lists(const R50Tile &t, float *stage, int n, int CB, R50Lists &ls) {
    const int L = t.n4 / 2;
    if (L < 5) return false;  // no room for a padded list; block-uniform
    int2 *sc = reinterpret_cast<int2 *>(stage), *sr = sc + n * L;
    bool ok = true;
    for (int q = threadIdx.x; q < 2 * n; q += blockDim.x) {
        const bool isrow = q >= n;
        const int j = isrow ? q - n : q;
        int2 *dst = (isrow ? sr : sc) + j * L;
        int cnt = 0;
        for (int y = 0; y < n; ++y) {
            const float v = isrow ? t.A[j * t.n4 + y] : t.A[y * t.n4 + j];
            if (v != 0.f) {
                ++cnt;
                if (cnt <= L - 4) dst[cnt] = make_int2(y * CB, __float_as_int(v));  // row offset, value
            }
        }
        dst[0] = make_int2(cnt, 0);
        for (int e = min(cnt, L - 4) + 1; e < L; ++e) dst[e] = make_int2(0, 0);  // padding: r50_sdot reads in fours
        ok = ok && cnt <= L - 4;
    }
    ok = __syncthreads_and(ok);
    int2 *fin = reinterpret_cast<int2 *>(t.A);
    if (ok) {
        for (int i = threadIdx.x; i < 2 * n * L; i += blockDim.x) fin[i] = sc[i];
    }
    __syncthreads();
    ls.col = fin;
    ls.row = fin + n * L;
    ls.L = L;
    return ok;
}
__device__ __forceinline__ void r50_cp4(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void r50_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// VW consecutive channels (VW = 1, 2, 4): asynchronous global -> shared copy, shared-memory load, streaming store
template <int VW>
__device__ __forceinline__ void r50_cpv(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(VW * 4) : "memory");
}
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_lds(const float *p) {
    R50Vec<VW> r;
    if (VW == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        r.x[0] = t.x, r.x[1 % VW] = t.y, r.x[2 % VW] = t.z, r.x[3 % VW] = t.w;
    } else if (VW == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        r.x[0] = t.x, r.x[1 % VW] = t.y;
    } else {
        r.x[0] = *p;
    }
    return r;
}
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_ldv(const float *p) {  // read-only global
    R50Vec<VW> r;
    if (VW == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        r.x[0] = t.x, r.x[1 % VW] = t.y, r.x[2 % VW] = t.z, r.x[3 % VW] = t.w;
    } else if (VW == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        r.x[0] = t.x, r.x[1 % VW] = t.y;
    } else {
        r.x[0] = __ldg(p);
    }
    return r;
}
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_ldp(const float *p) {  // plain (coherent) global load
    R50Vec<VW> r;
    if (VW == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        r.x[0] = t.x, r.x[1 % VW] = t.y, r.x[2 % VW] = t.z, r.x[3 % VW] = t.w;
    } else if (VW == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        r.x[0] = t.x, r.x[1 % VW] = t.y;
    } else {
        r.x[0] = *p;
    }
    return r;
}
template <int VW>
__device__ __forceinline__ void r50_stv(float *p, const R50Vec<VW> &v, bool streaming) {
    if (VW == 4) {
        const float4 t = make_float4(v.x[0], v.x[1 % VW], v.x[2 % VW], v.x[3 % VW]);
        if (streaming) __stcs(reinterpret_cast<float4 *>(p), t); else *reinterpret_cast<float4 *>(p) = t;
    } else if (VW == 2) {
        const float2 t = make_float2(v.x[0], v.x[1 % VW]);
        if (streaming) __stcs(reinterpret_cast<float2 *>(p), t); else *reinterpret_cast<float2 *>(p) = t;
    } else {
        if (streaming) __stcs(p, v.x[0]); else *p = v.x[0];
    }
}
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_zero() {
    R50Vec<VW> r;
#pragma unroll
    for (int k = 0; k < VW; ++k) r.x[k] = 0.f;
    return r;
}
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_scale(const R50Vec<VW> &v, float s) {
    R50Vec<VW> r;
#pragma unroll
    for (int k = 0; k < VW; ++k) r.x[k] = v.x[k] * s;
    return r;
}
// sum over the list entries of row[y] * value for the thread's VW channels; four entries per step so that the list
// and row reads of a step are independent (entries past the count are {0, 0.0f} padding up to the list's L slots,
// and count <= L - 4)
template <int VW>
__device__ __forceinline__ R50Vec<VW> r50_sdotv(const float *rowf, const int2 *list) {
    const int cnt = list[0].x;
    R50Vec<VW> acc = r50_zero<VW>();
    for (int e = 1; e <= cnt; e += 4) {
        int2 en[4];
        R50Vec<VW> g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = list[e + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = r50_lds<VW>(rowf + en[u].x);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < VW; ++k) acc.x[k] = fmaf(g[u].x[k], __int_as_float(en[u].y), acc.x[k]);
    }
    return acc;
}
// dense: r0[i] = sum_y row[y] M0[y][q0+i], r1[i] = sum_y row[y] M1[y][q0+i]   (i = 0..3)
template <int VW>
__device__ __forceinline__ void r50_dot4x2v(const float *rowf, int CB, const float *M0, const float *M1, int n4, int n, int q0,
                                            R50Vec<VW> (&r0)[4], R50Vec<VW> (&r1)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r0[i] = r50_zero<VW>(), r1[i] = r50_zero<VW>();
    for (int y = 0; y < n; ++y) {
        const R50Vec<VW> g = r50_lds<VW>(rowf + y * CB);
        const float4 m0 = *reinterpret_cast<const float4 *>(M0 + y * n4 + q0);
        const float4 m1 = *reinterpret_cast<const float4 *>(M1 + y * n4 + q0);
        const float a0[4] = {m0.x, m0.y, m0.z, m0.w}, a1[4] = {m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < VW; ++k) {
                r0[i].x[k] = fmaf(g.x[k], a0[i], r0[i].x[k]);
                r1[i].x[k] = fmaf(g.x[k], a1[i], r1[i].x[k]);
            }
    }
}
// dense: r[i] = sum_y rowA[y] M0[y][q0+i] + rowB[y] M1[y][q0+i]
template <int VW>
__device__ __forceinline__ void r50_dot4_pairv(const float *rowA, const float *rowB, int CB, const float *M0, const float *M1,
                                               int n4, int n, int q0, R50Vec<VW> (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = r50_zero<VW>();
    for (int y = 0; y < n; ++y) {
        const R50Vec<VW> g0 = r50_lds<VW>(rowA + y * CB), g1 = r50_lds<VW>(rowB + y * CB);
        const float4 m0 = *reinterpret_cast<const float4 *>(M0 + y * n4 + q0);
        const float4 m1 = *reinterpret_cast<const float4 *>(M1 + y * n4 + q0);
        const float a0[4] = {m0.x, m0.y, m0.z, m0.w}, a1[4] = {m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < VW; ++k) r[i].x[k] = fmaf(g0.x[k], a0[i], fmaf(g1.x[k], a1[i], r[i].x[k]));
    }
}

// CTA = (x, instance, CB-channel chunk); blockDim.x = CB / VW threads, each owning VW consecutive channels.
template <int VW>
__global__ void __launch_bounds__(128) k_r50_fwd_out_tiled(R50Args a, int CB) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y, x = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    if (x >= n) return;
    const int fo = threadIdx.x * VW;
    const int f = blockIdx.z * CB + fo;
    const bool live = f < C;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const float *V = sc + S.vecs_off, *X = sc + S.scal_off;
    const int64_t row = (int64_t)n * C;
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    const R50Tile t = r50_tile(smem50, n, CB);
    r50_load_adj(t, tab, n);
    __syncthreads();
    R50Lists ls;
    const bool sparse = (2 * t.n4 <= CB) && r50_build_lists(t, t.row, n, CB, ls);
    float *o = a.out + inst * a.stride_out + ((int64_t)x * n) * ((int64_t)a.ncases * C) + f;  // + y*50C + k*C
    const int64_t ostride = (int64_t)a.ncases * C;
#pragma unroll 1
    for (int k = 0; k < a.ncases; ++k) {
        const R50Case cs = a.plan[k];
        if (cs.form == kFormFollower) continue;  // written together with its leader
        if (cs.form == 2) {
            // leader of a pair that reads the same plane the same way with the two orientations of A (or a single case)
            const int k2 = (int)(cs.flags >> 2) - 1;
            const bool lead_t = (cs.flags & 2) != 0;
            __syncthreads();  // previous users of t.row are done
            if (live) {
                const float *pl = sc + cs.id * S.plane + f + ((cs.flags & 1) ? (int64_t)x * C : (int64_t)x * row);
                const int64_t ps = (cs.flags & 1) ? row : (int64_t)C;
                for (int j = 0; j < n; ++j) r50_cpv<VW>(t.row + j * CB + fo, pl + j * ps);  // all n loads in flight
            }
            r50_cp_wait();
            __syncthreads();
            if (live && sparse) {
                for (int y = 0; y < n; ++y) {
                    const int2 *l1 = (lead_t ? ls.col : ls.row) + y * ls.L, *l2 = (lead_t ? ls.row : ls.col) + y * ls.L;
                    r50_stv<VW>(o + y * ostride + (int64_t)k * C, r50_sdotv<VW>(t.row + fo, l1), true);
                    if (k2 >= 0) r50_stv<VW>(o + y * ostride + (int64_t)k2 * C, r50_sdotv<VW>(t.row + fo, l2), true);
                }
            } else if (live) {
                // out[x,y] = sum_j PLv[x,j] Am[y,j];  Am[y,j] = A[y,j] (flag bit 1 clear) -> M = At ;  A[j,y] (set) -> M = A
                for (int y0 = 0; y0 < n; y0 += 4) {
                    R50Vec<VW> r0[4], r1[4];
                    r50_dot4x2v<VW>(t.row + fo, CB, lead_t ? t.A : t.At, lead_t ? t.At : t.A, t.n4, n, y0, r0, r1);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (y0 + i < n) {
                            r50_stv<VW>(o + (y0 + i) * ostride + (int64_t)k * C, r0[i], true);
                            if (k2 >= 0) r50_stv<VW>(o + (y0 + i) * ostride + (int64_t)k2 * C, r1[i], true);
                        }
                }
            }
        } else if (live) {
            // eight loads in flight, then eight stores (a load -> store loop would expose the full latency per element)
            if (cs.form == kFormOff) {
                for (int y = 0; y < n; ++y) r50_stv<VW>(o + y * ostride + (int64_t)k * C, r50_zero<VW>(), true);
            } else if (cs.form == 0) {
                const float *pl = sc + cs.id * S.plane + (int64_t)x * row + f;
                const float sv = scal[cs.aux];
                constexpr int LB = 16;  // loads in flight per thread
                for (int y0 = 0; y0 < n; y0 += LB) {
                    R50Vec<VW> v[LB];
#pragma unroll
                    for (int u = 0; u < LB; ++u) v[u] = r50_ldv<VW>(pl + (int64_t)min(y0 + u, n - 1) * C);
#pragma unroll
                    for (int u = 0; u < LB; ++u)
                        if (y0 + u < n) r50_stv<VW>(o + (y0 + u) * ostride + (int64_t)k * C, r50_scale<VW>(v[u], sv), true);
                }
            } else {
                R50Vec<VW> v;
                const float *w;
                if (cs.form == 1) {
                    v = r50_ldv<VW>(V + cs.id * S.vec + x * C + f);
                    w = tab + (cs.aux ? AL.cs() : AL.r());
                } else {
                    v = r50_ldv<VW>(X + cs.id * C + f);
                    w = tab + x * n;
                }
                for (int y0 = 0; y0 < n; y0 += 8) {
                    float wv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + min(y0 + u, n - 1));
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (y0 + u < n) r50_stv<VW>(o + (y0 + u) * ostride + (int64_t)k * C, r50_scale<VW>(v, wv[u]), true);
                }
            }
        }
    }
}

// ---- vector kernels (C % 4 == 0, 16-byte aligned operands, n <= 48, every row and column of A with <= kAdjL non-zeros) -------
// Shared helpers of k_r50_fwd_out_v4 / k_r50_bwd_planes_v4: a thread owns FOUR channels (one 16-byte access per cell); eight lanes
// cover a 32-channel chunk; threadIdx / 8 is the y (forward) or j (backward) coordinate, so the whole row x is in flight at once.
constexpr int V4_CB = 32, V4_Q = V4_CB / 4, V4_MAXN = 48;
__device__ __forceinline__ float4 f4_ldg(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void f4_fma(float4 &v, const float4 &x, float s) {
    v.x = fmaf(x.x, s, v.x), v.y = fmaf(x.y, s, v.y), v.z = fmaf(x.z, s, v.z), v.w = fmaf(x.w, s, v.w);
}
__device__ __forceinline__ void f4_add(float4 &v, const float4 &x) { v.x += x.x, v.y += x.y, v.z += x.z, v.w += x.w; }
__device__ __forceinline__ float4 f4_scale(const float4 &x, float s) { return make_float4(x.x * s, x.y * s, x.z * s, x.w * s); }
__device__ __forceinline__ void cp16_cg(void *dst_smem, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Forward output, CTA = (x, instance, 32-channel chunk), thread = (y, channel quad).  The <= 9 plane rows the form-2 cases read
// (PL[x, j] or PL[j, x], j = 0..n-1) and the row / column lists of A are staged in shared memory with ONE round of 16-byte
// cp.async (the form-0 plane values of the thread's own cell are loaded into registers meanwhile); a thread then walks the
// row list and the column list of its y once each, accumulating all tiles per entry, and writes its 50 cells.  The round-1
// kernel staged one tile at a time in a 64-thread CTA (two barriers and a global round trip per case, 14 % warps active).
__host__ __device__ inline size_t r50_out4_smem(int nm) {
    return (size_t)9 * nm * V4_Q * 16 + (size_t)4 * nm * kAdjL * 4 + (size_t)(kVecs + kScals) * V4_Q * 16;
}
__global__ void __launch_bounds__(V4_MAXN * V4_Q, 2) k_r50_fwd_out_v4(R50Args a) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int nchunk = (C + V4_CB - 1) / V4_CB;
    const int chunk = blockIdx.x % nchunk, x = blockIdx.x / nchunk;  // the chunks of a row run together (they interleave in out)
    if (x >= n) return;
    const int q = threadIdx.x & (V4_Q - 1), y = threadIdx.x / V4_Q, nthr = blockDim.x;
    const int f = chunk * V4_CB + q * 4;
    const bool live = f < C && y < n;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const int64_t row = (int64_t)n * C;
    float4 *tiles = reinterpret_cast<float4 *>(smem50);                        // [ntiles][n][V4_Q]
    int2 *lists = reinterpret_cast<int2 *>(tiles + (size_t)9 * nm * V4_Q);      // rowl[n][kAdjL] | coll[n][kAdjL]
    float4 *vx = reinterpret_cast<float4 *>(lists + (size_t)2 * nm * kAdjL);    // V[6][V4_Q] | X[5][V4_Q]
    for (int i = threadIdx.x; i < a.ntiles * n * V4_Q; i += nthr) {
        const int qq = i & (V4_Q - 1), j = (i / V4_Q) % n, t = i / (V4_Q * n);
        const int ff = min(chunk * V4_CB + qq * 4, C - 4);
        const float *pl = sc + a.tile_pid[t] * S.plane + ff;
        cp16_cg(tiles + i, a.tile_or[t] ? pl + (int64_t)j * row + (int64_t)x * C : pl + (int64_t)x * row + (int64_t)j * C);
    }
    if (a.ntiles > 0) {  // the two list blocks are contiguous in the table (rowl | coll), nm rows each; unused without form-2 cases
        const float *src = tab + AL.rowl();
        for (int i = threadIdx.x; i < nm * kAdjL; i += nthr) cp16_cg(lists + 2 * i, src + 4 * i);
    }
    for (int i = threadIdx.x; i < (kVecs + kScals) * V4_Q; i += nthr) {
        const int qq = i & (V4_Q - 1), v = i / V4_Q;
        const int ff = min(chunk * V4_CB + qq * 4, C - 4);
        cp16_cg(vx + i, v < kVecs ? sc + S.vecs_off + v * S.vec + (int64_t)x * C + ff : sc + S.scal_off + (v - kVecs) * C + ff);
    }
    cp_commit();
    const int yy = min(y, n - 1), fc = min(f, C - 4);
    const int64_t cell = ((int64_t)x * n + yy) * C + fc;
    // form 0 of the thread's own cell: issued before the wait so that they fly with the staging
    float4 z[15];
#pragma unroll
    for (int i = 0; i < 15; ++i)
        if (i < a.nz) z[i] = f4_ldg(sc + a.z_pid[i] * S.plane + cell);
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    const float wr = tab[AL.r() + yy], wc = tab[AL.cs() + yy], axy = tab[x * n + yy];
    const int cr = min(reinterpret_cast<const int *>(tab + AL.cnt())[yy], kAdjL), cc = min(reinterpret_cast<const int *>(tab + AL.cnt())[nm + yy], kAdjL);
    cp_wait_group<0>();
    __syncthreads();
    if (!live) return;
    float *o = a.out + inst * a.stride_out + ((int64_t)x * n + y) * ((int64_t)a.ncases * C) + f;  // + k*C
#pragma unroll
    for (int i = 0; i < 15; ++i)
        if (i < a.nz) __stcs(reinterpret_cast<float4 *>(o + (int64_t)a.z_k[i] * C), f4_scale(z[i], scal[a.z_aux[i]]));
#pragma unroll 1
    for (int k = 0; k < a.ncases; ++k) {
        const R50Case cs = a.plan[k];
        if (cs.form == 1)
            __stcs(reinterpret_cast<float4 *>(o + (int64_t)k * C), f4_scale(vx[cs.id * V4_Q + q], cs.aux ? wc : wr));
        else if (cs.form == 3)
            __stcs(reinterpret_cast<float4 *>(o + (int64_t)k * C), f4_scale(vx[(kVecs + cs.id) * V4_Q + q], axy));
        else if (cs.form == kFormOff)
            __stcs(reinterpret_cast<float4 *>(o + (int64_t)k * C), make_float4(0.f, 0.f, 0.f, 0.f));
    }
    // form 2: out_k[x,y] = sum_j PL[x,j] A[y,j] (row list of y) and sum_j PL[x,j] A[j,y] (column list of y), all tiles per entry
    const float4 *tq = tiles + q;
    const size_t tstride = (size_t)n * V4_Q;
    const bool sparse = __float_as_int(tab[AL.scal() + 2]) <= kAdjL;
    if (a.ntiles == 0) return;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const int2 *l = lists + (size_t)(side * nm + y) * kAdjL;
        const int cnt = side ? cc : cr;
        float4 acc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sparse) {
            for (int e = 0; e < cnt; ++e) {
                const int2 en = l[e];
                const float w = __int_as_float(en.y);
#pragma unroll
                for (int t = 0; t < 9; ++t)
                    if (t < a.ntiles) f4_fma(acc[t], tq[t * tstride + en.x], w);
            }
        } else {  // some row or column of A has more than kAdjL non-zeros: walk the dense row / column
            for (int j = 0; j < n; ++j) {
                const float w = side ? tab[j * n + y] : tab[y * n + j];
                if (w != 0.f) {
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        if (t < a.ntiles) f4_fma(acc[t], tq[t * tstride + j * V4_Q], w);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            if (t < a.ntiles) {
                const int k = side ? a.tile_kcol[t] : a.tile_krow[t];
                if (k != 0xff) __stcs(reinterpret_cast<float4 *>(o + (int64_t)k * C), acc[t]);
            }
        }
    }
}

// Gradient planes, CTA = (x, instance, 32-channel chunk), thread = (j, channel quad); the transpose of k_r50_fwd_out_v4.
// PASS 0 writes row x of every plane exactly once: planes 3..11 are a scaled copy of one gout slab row (loaded before the staging
// wait), planes 0, 1, 2, 12, 13, 14 add the folded vector / scalar gradients and the form-2 pairs that read the plane as [x, j]:
// d PL[x,j] = sum_y g_krow[x,y] A[y,j] (column list of j) + g_kcol[x,y] A[j,y] (row list of j), the slab rows g_k[x, :] staged
// in shared memory with one round of cp.async.  PASS 1 adds the pairs that read the plane as [j, x] into column x.
__host__ __device__ inline size_t r50_planes4_smem(int nm) { return (size_t)10 * nm * V4_Q * 16 + (size_t)4 * nm * kAdjL * 4; }
template <int PASS>
__global__ void __launch_bounds__(V4_MAXN * V4_Q, 2) k_r50_bwd_planes_v4(R50Args a) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int nchunk = (C + V4_CB - 1) / V4_CB;
    const int chunk = blockIdx.x % nchunk, x = blockIdx.x / nchunk;
    if (x >= n) return;
    const int q = threadIdx.x & (V4_Q - 1), j = threadIdx.x / V4_Q, nthr = blockDim.x;
    const int f = chunk * V4_CB + q * 4;
    const bool live = f < C && j < n;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const int64_t row = (int64_t)n * C, gstride = (int64_t)a.ncases * C;
    const float *gx = a.out + inst * a.stride_out + ((int64_t)x * n) * gstride;  // gout[x, y, k, f] = gx[y * gstride + k * C + f]
    float4 *slabs = reinterpret_cast<float4 *>(smem50);                       // [2 * npt][n][V4_Q]: krow rows, then kcol rows
    int2 *lists = reinterpret_cast<int2 *>(slabs + (size_t)10 * nm * V4_Q);    // rowl[n][kAdjL] | coll[n][kAdjL]
    const int npt = a.npt[PASS];
    for (int i = threadIdx.x; i < 2 * npt * n * V4_Q; i += nthr) {
        const int qq = i & (V4_Q - 1), y = (i / V4_Q) % n, s = i / (V4_Q * n);
        const int k = s < npt ? a.pt_krow[PASS][s] : a.pt_kcol[PASS][s - npt];
        const int ff = min(chunk * V4_CB + qq * 4, C - 4);
        if (k != 0xff)
            cp16_cg(slabs + i, gx + (int64_t)y * gstride + (int64_t)k * C + ff);
        else
            slabs[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // the dropped member of a pair
    }
    if (npt > 0) {
        const float *src = tab + AL.rowl();
        for (int i = threadIdx.x; i < nm * kAdjL; i += nthr) cp16_cg(lists + 2 * i, src + 4 * i);
    }
    cp_commit();
    const int jj = min(j, n - 1), fc = min(f, C - 4);
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    const float *gcell = gx + (int64_t)jj * gstride + fc;  // gout[x, j, k, f] = gcell[k * C]
    float *prow = sc + (int64_t)x * row + (int64_t)jj * C + fc;  // PL[x, j] of plane 0
    float *pcol = sc + (int64_t)jj * row + (int64_t)x * C + fc;  // PL[j, x] of plane 0
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 old1[5];
    const unsigned pm = a.pm;
    if (PASS == 0 && (pm & 0x0ff8u)) {
        // planes 3..11: one form-0 case each, nothing else
        float4 zs[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) zs[i] = a.pz0[3 + i] != 0xff ? f4_ldg(gcell + (int64_t)a.pz0[3 + i] * C) : zero4;
        if (live) {
#pragma unroll
            for (int i = 0; i < 9; ++i)
                *reinterpret_cast<float4 *>(prow + (3 + i) * S.plane) = f4_scale(zs[i], a.pz0[3 + i] != 0xff ? scal[a.pza0[3 + i]] : 0.f);
        }
    } else if (PASS == 1) {
#pragma unroll
        for (int t = 0; t < 5; ++t)
            if (t < npt) old1[t] = *reinterpret_cast<const float4 *>(pcol + a.pt_pid[1][t] * S.plane);
    }
    const int cr = min(reinterpret_cast<const int *>(tab + AL.cnt())[jj], kAdjL), cc = min(reinterpret_cast<const int *>(tab + AL.cnt())[nm + jj], kAdjL);
    const bool sparse = __float_as_int(tab[AL.scal() + 2]) <= kAdjL;
    cp_wait_group<0>();
    __syncthreads();
    if (!live) return;
    float4 acc[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) acc[t] = zero4;
    const float4 *sq = slabs + q;
    const size_t sstride = (size_t)n * V4_Q;
#pragma unroll 1
    for (int side = 0; side < (npt > 0 ? 2 : 0); ++side) {
        // side 0: the krow cases, A[y,j] over the column list of j; side 1: the kcol cases, A[j,y] over the row list of j
        const int2 *l = lists + (size_t)((side ? 0 : nm) + j) * kAdjL;
        const int cnt = side ? cr : cc;
        const float4 *sp = sq + (side ? npt * sstride : 0);
        if (sparse) {
            for (int e = 0; e < cnt; ++e) {
                const int2 en = l[e];
                const float w = __int_as_float(en.y);
#pragma unroll
                for (int t = 0; t < 5; ++t)
                    if (t < npt) f4_fma(acc[t], sp[t * sstride + en.x], w);
            }
        } else {
            for (int y = 0; y < n; ++y) {
                const float w = side ? tab[j * n + y] : tab[y * n + j];
                if (w != 0.f) {
#pragma unroll
                    for (int t = 0; t < 5; ++t)
                        if (t < npt) f4_fma(acc[t], sp[t * sstride + y * V4_Q], w);
                }
            }
        }
    }
    if (PASS == 1) {
#pragma unroll
        for (int t = 0; t < 5; ++t)
            if (t < npt) {
                f4_add(old1[t], acc[t]);
                *reinterpret_cast<float4 *>(pcol + a.pt_pid[1][t] * S.plane) = old1[t];
            }
        return;
    }
    // PASS 0, planes 0, 1, 2, 12, 13, 14: folded vector / scalar gradients + form-0 cases + the pair
    const float *gV = sc + S.vecs_off + fc, *gX = sc + S.scal_off + fc;
    auto tile_of = [&](int pid) {
        float4 v = zero4;
#pragma unroll
        for (int t = 0; t < 5; ++t)
            if (t < npt && a.pt_pid[0][t] == pid) f4_add(v, acc[t]);
        return v;
    };
    auto form0 = [&](int pid, float4 &v) {
        if (a.pz0[pid] != 0xff) f4_fma(v, f4_ldg(gcell + (int64_t)a.pz0[pid] * C), scal[a.pza0[pid]]);
        if (pid < 3 && a.pz1[pid < 3 ? pid : 0] != 0xff) f4_fma(v, f4_ldg(gcell + (int64_t)a.pz1[pid < 3 ? pid : 0] * C), scal[a.pza1[pid < 3 ? pid : 0]]);
    };
    const int64_t vx = (int64_t)x * C, vj = (int64_t)j * C;
    if (pm & (1u << 0)) {
        float4 v = tile_of(0);
        f4_add(v, f4_ldg(gV + 0 * S.vec + vx)), f4_add(v, f4_ldg(gX + 0 * C)), f4_add(v, f4_ldg(gV + 1 * S.vec + vj));
        form0(0, v);
        *reinterpret_cast<float4 *>(prow + 0 * S.plane) = v;
    }
    if (pm & (1u << 1)) {
        float4 v = tile_of(1);
        f4_add(v, f4_ldg(gV + 2 * S.vec + vj));
        form0(1, v);
        *reinterpret_cast<float4 *>(prow + 1 * S.plane) = v;
    }
    if (pm & (1u << 2)) {
        float4 v = tile_of(2);
        form0(2, v);
        *reinterpret_cast<float4 *>(prow + 2 * S.plane) = v;
    }
    if (pm & (1u << 12)) {
        float4 v = tile_of(12);
        f4_add(v, f4_ldg(gV + 5 * S.vec + vj)), f4_add(v, f4_ldg(gX + 1 * C));
        if (x == j) f4_add(v, f4_ldg(gX + 4 * C));
        form0(12, v);
        *reinterpret_cast<float4 *>(prow + 12 * S.plane) = v;
    }
    if (pm & (1u << 13)) {
        float4 v = tile_of(13);
        f4_add(v, f4_ldg(gV + 4 * S.vec + vj)), f4_add(v, f4_ldg(gX + 2 * C));
        form0(13, v);
        *reinterpret_cast<float4 *>(prow + 13 * S.plane) = v;
    }
    if (pm & (1u << 14)) {
        float4 v = tile_of(14);
        f4_add(v, f4_ldg(gV + 3 * S.vec + vx)), f4_add(v, f4_ldg(gX + 3 * C));
        form0(14, v);
        *reinterpret_cast<float4 *>(prow + 14 * S.plane) = v;
    }
}

// Gradient planes, one CTA per (x, instance, channel chunk); blockDim.x = CB / VW threads of VW channels each.
// pass 0 writes row x of every plane exactly once: the folded vector / scalar gradients + the plane's form-0 cases
//        (scaled copies of a slab row) + the pair of form-2 cases that read the plane as [x, j].
// pass 1 adds the pair of form-2 cases that read the plane as [j, x] into column x (read-modify-write, one writer per
//        element within the pass).
// The two slab rows of a pair are staged in shared memory with cp.async, all 2n loads in flight at once.
template <int PASS, int VW>
__global__ void __launch_bounds__(128) k_r50_bwd_planes_tiled(R50Args a, int CB) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y, x = blockIdx.x;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    if (x >= n) return;
    const int fo = threadIdx.x * VW;
    const int f = blockIdx.z * CB + fo;
    const bool live = f < C;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    float *sc = a.scratch + inst * a.scratch_words;
    const float *gV = sc + S.vecs_off, *gX = sc + S.scal_off;
    const float *g = a.out + inst * a.stride_out + ((int64_t)x * n) * ((int64_t)a.ncases * C) + f;  // + y*50C + k*C
    const int64_t gstride = (int64_t)a.ncases * C, row = (int64_t)n * C;
    const float scal[3] = {1.f, tab[AL.scal()], tab[AL.scal() + 1]};
    float *rowA = smem50, *rowB = smem50 + (size_t)n * CB;
    const R50Tile t = r50_tile(smem50 + (size_t)n * CB, n, CB);  // t.row aliases rowB
    r50_load_adj(t, tab, n);
    __syncthreads();
    R50Lists ls;
    const bool sparse = (t.n4 <= CB) && r50_build_lists(t, rowA, n, CB, ls);

#pragma unroll 1
    for (int pid = 0; pid < kPlanes; ++pid) {
        int k1 = -1, k2 = -1, z0 = -1, z1 = -1;  // the form-2 pair of this pass, the form-0 cases
        for (int k = 0; k < a.ncases; ++k) {
            const R50Case cs = a.plan[k];
            if (cs.id != pid) continue;
            if (r50_is2(cs.form) && (int)(cs.flags & 1) == PASS) {
                if (cs.flags & 2) k2 = k; else k1 = k;
            } else if (cs.form == 0 && PASS == 0) {
                if (z0 < 0) z0 = k; else z1 = k;
            }
        }
        const bool pair = k1 >= 0 || k2 >= 0;  // a dropped slab leaves a pair with one member: its row is zero-filled
        const float sz0 = z0 >= 0 ? scal[a.plan[z0].aux] : 0.f, sz1 = z1 >= 0 ? scal[a.plan[z1].aux] : 0.f;
        if (PASS == 1 && !pair) continue;
        __syncthreads();
        if (pair && live) {
            if (k1 >= 0)
                for (int y = 0; y < n; ++y) r50_cpv<VW>(rowA + y * CB + fo, g + y * gstride + (int64_t)k1 * C);
            else
                for (int y = 0; y < n; ++y) r50_stv<VW>(rowA + y * CB + fo, r50_zero<VW>(), false);
            if (k2 >= 0)
                for (int y = 0; y < n; ++y) r50_cpv<VW>(rowB + y * CB + fo, g + y * gstride + (int64_t)k2 * C);
            else
                for (int y = 0; y < n; ++y) r50_stv<VW>(rowB + y * CB + fo, r50_zero<VW>(), false);
        }
        r50_cp_wait();
        __syncthreads();
        if (!live) continue;
        // d PLv[x, j] = sum_y g_k1[x,y] A[y,j] + g_k2[x,y] A[j,y]
        float *dst = sc + pid * S.plane + f + (PASS ? (int64_t)x * C : (int64_t)x * row);
        const int64_t ds = PASS ? row : (int64_t)C;
        // The raw loads of the NEXT group are issued before the dot products of the current one and only combined
        // afterwards (no arithmetic on them in between: an in-order warp would stall at the first use).
        constexpr int G = 8 / VW < 2 ? 2 : 8 / VW;
        const float *vq = nullptr;  // the vector gradient indexed by q that folds into this plane
        R50Vec<VW> cst = r50_zero<VW>(), gx4 = r50_zero<VW>();
        auto add = [](R50Vec<VW> u, const R50Vec<VW> &v) {
#pragma unroll
            for (int k = 0; k < VW; ++k) u.x[k] += v.x[k];
            return u;
        };
        if (PASS == 0) {
            if (pid == 0) vq = gV + 1 * S.vec + f, cst = add(r50_ldv<VW>(gV + 0 * S.vec + x * C + f), r50_ldv<VW>(gX + 0 * C + f));
            if (pid == 1) vq = gV + 2 * S.vec + f;
            if (pid == 12) vq = gV + 5 * S.vec + f, cst = r50_ldv<VW>(gX + 1 * C + f), gx4 = r50_ldv<VW>(gX + 4 * C + f);
            if (pid == 13) vq = gV + 4 * S.vec + f, cst = r50_ldv<VW>(gX + 2 * C + f);
            if (pid == 14) cst = add(r50_ldv<VW>(gV + 3 * S.vec + x * C + f), r50_ldv<VW>(gX + 3 * C + f));
        }
        auto load_raw = [&](int j0, R50Vec<VW>(&raw)[G][3]) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int q = min(j0 + i, n - 1);
                if (PASS == 1) {
                    raw[i][0] = r50_ldp<VW>(dst + q * ds);
                } else {
                    raw[i][0] = vq ? r50_ldv<VW>(vq + q * C) : r50_zero<VW>();
                    raw[i][1] = z0 >= 0 ? r50_ldv<VW>(g + q * gstride + (int64_t)z0 * C) : r50_zero<VW>();
                    raw[i][2] = z1 >= 0 ? r50_ldv<VW>(g + q * gstride + (int64_t)z1 * C) : r50_zero<VW>();
                }
            }
        };
        R50Vec<VW> nxt[G][3];
        load_raw(0, nxt);
        for (int j0 = 0; j0 < n; j0 += G) {
            R50Vec<VW> cur[G][3], r[G];
#pragma unroll
            for (int i = 0; i < G; ++i) cur[i][0] = nxt[i][0], cur[i][1] = nxt[i][1], cur[i][2] = nxt[i][2], r[i] = r50_zero<VW>();
            if (j0 + G < n) load_raw(j0 + G, nxt);
            if (pair && sparse) {
#pragma unroll
                for (int i = 0; i < G; ++i)
                    if (j0 + i < n)
                        r[i] = add(r50_sdotv<VW>(rowA + fo, ls.col + (j0 + i) * ls.L), r50_sdotv<VW>(rowB + fo, ls.row + (j0 + i) * ls.L));
            } else if (pair) {
#pragma unroll
                for (int h = 0; h < G; h += 4) {
                    if (j0 + h >= n) break;
                    R50Vec<VW> r4[4];
                    r50_dot4_pairv<VW>(rowA + fo, rowB + fo, CB, t.A, t.At, t.n4, n, j0 + h, r4);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (h + i < G) r[h + i] = r4[i];
                }
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                R50Vec<VW> base;
#pragma unroll
                for (int k = 0; k < VW; ++k) {
                    if (PASS == 1)
                        base.x[k] = cur[i][0].x[k];
                    else
                        base.x[k] = fmaf(sz1, cur[i][2].x[k], fmaf(sz0, cur[i][1].x[k], cst.x[k] + cur[i][0].x[k])) +
                                    ((pid == 12 && x == j0 + i) ? gx4.x[k] : 0.f);
                }
                if (j0 + i < n) r50_stv<VW>(dst + (j0 + i) * ds, add(base, r[i]), false);
            }
        }
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

// Tiled scatter: CTA = (4 consecutive a, instance, 32-channel chunk), warp <-> b (strided), lane <-> channel.
// The four [a,c]-type gradient planes of the CTA's four a are staged in shared memory once; the four [b,c]-type values
// are loaded once per (b, c) and reused for the four a; the [a,b]-type values live in registers across the c loop.
// Per output element that is one global load and four shared-memory loads instead of twelve global loads.
constexpr int SC_TA = 4, SC_CB = 32, SC_THREADS = 256, SC_UC = 4;
__host__ __device__ inline size_t r50_scatter_smem(int nm) { return ((size_t)4 * SC_TA * nm * SC_CB + 3 * nm) * 4; }

__global__ void __launch_bounds__(SC_THREADS, 2) k_r50_bwd_scatter_tiled(R50Args a) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int a0 = blockIdx.z * SC_TA;
    if (a0 >= n) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * SC_CB + lane;  // channel chunk fastest: the CTAs that interleave 128-byte pieces of a cell run together
    const bool live = f < C;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    float *AC = smem50;                                  // [4 planes][SC_TA][n][SC_CB]
    float *wr = smem50 + (size_t)4 * SC_TA * nm * SC_CB;  // r | cs | dg
    float *wc = wr + nm, *wd = wc + nm;
    for (int i = threadIdx.x; i < n; i += SC_THREADS) {
        wr[i] = tab[AL.r() + i];
        wc[i] = tab[AL.cs() + i];
        wd[i] = tab[AL.dg() + i];
    }
    const int plane_id[4] = {1, 6, 7, 8};
    for (int i = warp; i < 4 * SC_TA * n; i += SC_THREADS / 32) {  // i = (p * SC_TA + ai) * n + c
        const int c = i % n, ai = (i / n) % SC_TA, p = i / (n * SC_TA);
        const int aa = a0 + ai;
        float *d = AC + (size_t)i * SC_CB + lane;
        if (live && aa < n)
            r50_cp4(d, sc + plane_id[p] * S.plane + ((int64_t)aa * n + c) * C + f);  // asynchronous: no load -> store chain
        else
            *d = 0.f;
    }
    r50_cp_wait();
    __syncthreads();
    if (!live) return;
    const int64_t slab = (int64_t)n * n * C;
    float ra[SC_TA], ca[SC_TA], da[SC_TA];
#pragma unroll
    for (int ai = 0; ai < SC_TA; ++ai) {
        const int aa = min(a0 + ai, n - 1);
        ra[ai] = wr[aa];
        ca[ai] = wc[aa];
        da[ai] = wd[aa];
    }
    for (int b = warp; b < n; b += SC_THREADS / 32) {
        float ab0[SC_TA], ab3[SC_TA], ab4[SC_TA], ab5[SC_TA], ab13[SC_TA], ab14[SC_TA];
#pragma unroll
        for (int ai = 0; ai < SC_TA; ++ai) {
            const int aa = min(a0 + ai, n - 1);
            const int64_t ab = ((int64_t)aa * n + b) * C + f;
            ab0[ai] = sc[0 * S.plane + ab];
            ab3[ai] = sc[3 * S.plane + ab];
            ab4[ai] = sc[4 * S.plane + ab];
            ab5[ai] = sc[5 * S.plane + ab];
            ab13[ai] = sc[13 * S.plane + ab];
            ab14[ai] = sc[14 * S.plane + ab];
        }
        const float rb = wr[b], cb = wc[b], db = wd[b];
        const float *bcp = sc + ((int64_t)b * n) * C + f;
        float nx[4][SC_UC];  // the next group's [b,c]-type values, loaded one group ahead of their use
        auto load_bc = [&](int c0) {
#pragma unroll
            for (int u = 0; u < SC_UC; ++u) {
                const int64_t o = (int64_t)min(c0 + u, n - 1) * C;
                nx[0][u] = bcp[2 * S.plane + o];
                nx[1][u] = bcp[9 * S.plane + o];
                nx[2][u] = bcp[10 * S.plane + o];
                nx[3][u] = bcp[11 * S.plane + o];
            }
        };
        load_bc(0);
        for (int c0 = 0; c0 < n; c0 += SC_UC) {
            float bc2[SC_UC], bc9[SC_UC], bc10[SC_UC], bc11[SC_UC];
#pragma unroll
            for (int u = 0; u < SC_UC; ++u) bc2[u] = nx[0][u], bc9[u] = nx[1][u], bc10[u] = nx[2][u], bc11[u] = nx[3][u];
            if (c0 + SC_UC < n) load_bc(c0 + SC_UC);
#pragma unroll
            for (int u = 0; u < SC_UC; ++u) {
                const int c = c0 + u;
                if (c >= n) break;
                const float rc = wr[c], cc = wc[c], dc = wd[c];
#pragma unroll
                for (int ai = 0; ai < SC_TA; ++ai) {
                    const int aa = a0 + ai;
                    if (aa < n) {
                        const float *acp = AC + ((size_t)ai * n + c) * SC_CB + lane;
                        const size_t pstride = (size_t)SC_TA * n * SC_CB;
                        float v = ab0[ai] + acp[0] + bc2[u];
                        v = fmaf(ab3[ai], rc, v);
                        v = fmaf(ab4[ai], cc, v);
                        v = fmaf(ab5[ai], dc, v);
                        v = fmaf(acp[pstride], rb, v);
                        v = fmaf(acp[2 * pstride], cb, v);
                        v = fmaf(acp[3 * pstride], db, v);
                        v = fmaf(bc9[u], ra[ai], v);
                        v = fmaf(bc10[u], ca[ai], v);
                        v = fmaf(bc11[u], da[ai], v);
                        if (aa == b) v += sc[12 * S.plane + ((int64_t)aa * n + c) * C + f];
                        if (aa == c) v += ab13[ai];
                        if (b == c) v += ab14[ai];
                        float *dst = (a.T.slabs ? a.T.slabs[(int64_t)inst * nm + aa] : a.T.base + inst * a.T.stride + (int64_t)aa * slab) +
                                     ((int64_t)b * n + c) * C + f;
                        __stcs(dst, (a.beta != 0.f) ? fmaf(a.beta, *dst, v) : v);
                    }
                }
            }
        }
    }
}

// Vector scatter (C % 4 == 0, dense 16-byte aligned gT, n <= 48): CTA = (S4_TA consecutive a, instance, 32-channel chunk); a
// thread owns FOUR channels (one 16-byte access per cell) of ONE b row; eight lanes cover the chunk, so a warp writes four
// 128-byte segments per instruction.  The [a,c]-type planes of the CTA's a are staged in shared memory (16-byte cp.async), the
// [a,b]-type values live in registers across the c loop, and the [b,c]-type values of step c + 2 are fetched with cp.async into
// a three-deep ring of thread-private shared-memory slots: they are shared by the S4_TA a, and -- unlike register loads, which
// ptxas puts on the same scoreboard as the step's shared-memory loads, so that the first use of an LDS result waits for the
// global loads in flight as well (ncu: 46 % of all stall samples on that one FADD) -- they complete out of the steps' way.
// The diagonal terms are kept out of the element loop: T[a,a,c] rides in the ring for the rows that own it, T[a,b,a] and
// T[a,b,b] are added into their two cells per (a, b) row after the loop.  Twelve FMA/FADD per element.
constexpr int S4_TA = 4, S4_CB = 32, S4_Q = S4_CB / 4, S4_D = 3, S4_MAXN = 48;
__host__ __device__ inline size_t r50_scatter4_smem(int nm) {
    return ((size_t)4 * S4_TA * nm * S4_CB + 4 * nm) * 4 + (size_t)S4_D * 5 * nm * S4_Q * 16;
}

__global__ void __launch_bounds__(S4_MAXN * S4_Q, 1) k_r50_bwd_scatter_v4(R50Args a) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int nchunk = (C + S4_CB - 1) / S4_CB;
    const int chunk = blockIdx.x % nchunk, a0 = (blockIdx.x / nchunk) * S4_TA;  // the a groups of an instance run together (L2)
    if (a0 >= n) return;
    const int q = threadIdx.x & (S4_Q - 1), b = threadIdx.x / S4_Q, nthr = blockDim.x;
    const int f = chunk * S4_CB + q * 4;
    const bool live = f < C && b < n;
    const R50Adj AL{nm};
    const float *tab = a.adjtab + (int64_t)inst * a.adjtab_words;
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    float4 *AC = reinterpret_cast<float4 *>(smem50);                                   // [4 planes][S4_TA][n][S4_Q]
    float4 *wt = reinterpret_cast<float4 *>(smem50 + (size_t)4 * S4_TA * nm * S4_CB);  // {r, cs, dg, 0}[n]
    float4 *ring = wt + nm;                                                            // [S4_D][5][nthr]
    for (int i = threadIdx.x; i < n; i += nthr) wt[i] = make_float4(tab[AL.r() + i], tab[AL.cs() + i], tab[AL.dg() + i], 0.f);
    {
        const int plane_id[4] = {1, 6, 7, 8};
        const int per = S4_TA * n * S4_Q;
        for (int i = threadIdx.x; i < 4 * per; i += nthr) {  // i = ((p * S4_TA + ai) * n + c) * S4_Q + q
            const int qq = i & (S4_Q - 1), c = (i / S4_Q) % n, ai = (i / (S4_Q * n)) % S4_TA, p = i / per;
            const int aa = a0 + ai, ff = chunk * S4_CB + qq * 4;
            if (aa < n && ff < C)
                cp16_cg(AC + i, sc + plane_id[p] * S.plane + ((int64_t)aa * n + c) * C + ff);
            else
                AC[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const bool own = b >= a0 && b < a0 + S4_TA;  // this row carries the T[a,a,c] term of a = b
    float4 *slot = ring + threadIdx.x;           // slot[(stage * 5 + k) * nthr]
    if (!own)
        for (int st = 0; st < S4_D; ++st) slot[(st * 5 + 4) * nthr] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *bcp = sc + ((int64_t)min(b, n - 1) * n) * C + min(f, C - 4);
    auto fetch = [&](int c, int st) {  // the [b,c]-type values of step c
        const float *src = bcp + (int64_t)c * C;
        float4 *d = slot + (size_t)st * 5 * nthr;
        cp16_cg(d, src + 2 * S.plane), cp16_cg(d + nthr, src + 9 * S.plane), cp16_cg(d + 2 * nthr, src + 10 * S.plane);
        cp16_cg(d + 3 * nthr, src + 11 * S.plane);
        if (own) cp16_cg(d + 4 * nthr, src + 12 * S.plane);
    };
    fetch(0, 0);
    cp_commit();  // group 0: the staged planes + step 0
    if (n > 1) fetch(1, 1);
    cp_commit();
    cp_wait_group<1>();
    __syncthreads();
    if (!live) return;
    const int64_t slab = (int64_t)n * n * C;
    const size_t pstride = (size_t)S4_TA * n * S4_Q;
    float ra[S4_TA], ca[S4_TA], da[S4_TA];
    float4 g0[S4_TA], g3[S4_TA], g4[S4_TA], g5[S4_TA];
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai) {
        const int aa = min(a0 + ai, n - 1);
        const float4 w = wt[aa];
        ra[ai] = w.x, ca[ai] = w.y, da[ai] = w.z;
        const int64_t ab = ((int64_t)aa * n + b) * C + f;
        g0[ai] = f4_ldg(sc + 0 * S.plane + ab), g3[ai] = f4_ldg(sc + 3 * S.plane + ab);
        g4[ai] = f4_ldg(sc + 4 * S.plane + ab), g5[ai] = f4_ldg(sc + 5 * S.plane + ab);
    }
    const float4 wb = wt[b];
    float *const dstb = a.T.base + inst * a.T.stride + (int64_t)a0 * slab + (int64_t)b * n * C + f;
    const bool accumulate = a.beta != 0.f;
    int st = 0, stn = 2;
#pragma unroll 1
    for (int c = 0; c < n; ++c) {
        if (c + 2 < n) fetch(c + 2, stn);
        cp_commit();
        cp_wait_group<2>();  // everything up to step c has landed (own slots only: no barrier needed)
        const float4 *d = slot + (size_t)st * 5 * nthr;
        const float4 c2 = d[0], c9 = d[nthr], c10 = d[2 * nthr], c11 = d[3 * nthr], c12 = d[4 * nthr];
        st = st + 1 == S4_D ? 0 : st + 1;
        stn = stn + 1 == S4_D ? 0 : stn + 1;
        const float4 wc = wt[c];
        const float4 *acp = AC + (size_t)c * S4_Q + q;
#pragma unroll
        for (int ai = 0; ai < S4_TA; ++ai) {
            const int aa = a0 + ai;
            if (aa < n) {
                const float4 *ap = acp + (size_t)ai * n * S4_Q;
                float4 v = ap[0];
                f4_add(v, g0[ai]);
                f4_add(v, c2);
                f4_fma(v, g3[ai], wc.x), f4_fma(v, g4[ai], wc.y), f4_fma(v, g5[ai], wc.z);
                f4_fma(v, ap[pstride], wb.x), f4_fma(v, ap[2 * pstride], wb.y), f4_fma(v, ap[3 * pstride], wb.z);
                f4_fma(v, c9, ra[ai]), f4_fma(v, c10, ca[ai]), f4_fma(v, c11, da[ai]);
                if (aa == b) f4_add(v, c12);  // T[a,a,c]
                float4 *dst = reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)c * C);
                if (accumulate) {
                    const float4 o = *dst;
                    v.x = fmaf(a.beta, o.x, v.x), v.y = fmaf(a.beta, o.y, v.y), v.z = fmaf(a.beta, o.z, v.z), v.w = fmaf(a.beta, o.w, v.w);
                }
                __stcs(dst, v);
            }
        }
    }
    // the other two diagonals, T[a,b,a] -> cell c = a and T[a,b,b] -> cell c = b of the rows this thread has just written:
    // the loads are issued together, then added into the (L2-resident) cells -- once per (a, b) row instead of a test per element
    float4 d13[S4_TA], d14[S4_TA], o13[S4_TA], o14[S4_TA];
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai) {
        const int aa = min(a0 + ai, n - 1);
        const int64_t ab = ((int64_t)aa * n + b) * C + f;
        d13[ai] = f4_ldg(sc + 13 * S.plane + ab), d14[ai] = f4_ldg(sc + 14 * S.plane + ab);
        o13[ai] = __ldcg(reinterpret_cast<const float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C));
        o14[ai] = __ldcg(reinterpret_cast<const float4 *>(dstb + (int64_t)ai * slab + (int64_t)b * C));
    }
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai) {
        const int aa = a0 + ai;
        if (aa < n) {
            if (aa == b) {  // the same cell takes both
                f4_add(o13[ai], d13[ai]), f4_add(o13[ai], d14[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C), o13[ai]);
            } else {
                f4_add(o13[ai], d13[ai]), f4_add(o14[ai], d14[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C), o13[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)b * C), o14[ai]);
            }
        }
    }
}

// The scatter of a plan without weighted planes (RisiContraction_4 / _10): gT[a,b,c] = g0[a,b] + g1[a,c] + g2[b,c] + diagonals, each
// term only if the plan materialised its plane (a.pm).  Same tiling as k_r50_bwd_scatter_v4 with one staged plane and a two-slot ring.
__host__ __device__ inline size_t r50_scatter4p_smem(int nm) { return (size_t)S4_TA * nm * S4_CB * 4 + (size_t)S4_D * 2 * nm * S4_Q * 16; }
__global__ void __launch_bounds__(S4_MAXN * S4_Q, 2) k_r50_bwd_scatter_plain(R50Args a) {
    extern __shared__ __align__(16) float smem50[];
    const int inst = blockIdx.y;
    const int n = a.b.n_of(inst), C = a.b.C, nm = a.b.n_max;
    const int nchunk = (C + S4_CB - 1) / S4_CB;
    const int chunk = blockIdx.x % nchunk, a0 = (blockIdx.x / nchunk) * S4_TA;
    if (a0 >= n) return;
    const int q = threadIdx.x & (S4_Q - 1), b = threadIdx.x / S4_Q, nthr = blockDim.x;
    const int f = chunk * S4_CB + q * 4;
    const bool live = f < C && b < n;
    const R50Scratch S(nm, C);
    const float *sc = a.scratch + inst * a.scratch_words;
    const unsigned pm = a.pm;
    const bool h0 = pm & 1u, h1 = pm & 2u, h2 = pm & 4u, h12 = pm & (1u << 12), h13 = pm & (1u << 13), h14 = pm & (1u << 14);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 *AC = reinterpret_cast<float4 *>(smem50);      // [S4_TA][n][S4_Q]: plane 1
    float4 *ring = AC + (size_t)S4_TA * nm * S4_Q;         // [S4_D][2][nthr]: plane 2, plane 12 (own rows)
    for (int i = threadIdx.x; i < S4_TA * n * S4_Q; i += nthr) {
        const int qq = i & (S4_Q - 1), c = (i / S4_Q) % n, ai = i / (S4_Q * n);
        const int aa = a0 + ai, ff = chunk * S4_CB + qq * 4;
        if (h1 && aa < n && ff < C)
            cp16_cg(AC + i, sc + 1 * S.plane + ((int64_t)aa * n + c) * C + ff);
        else
            AC[i] = zero4;
    }
    const bool own = h12 && b >= a0 && b < a0 + S4_TA;
    float4 *slot = ring + threadIdx.x;
    for (int st = 0; st < S4_D; ++st) {
        if (!h2) slot[(st * 2) * nthr] = zero4;
        if (!own) slot[(st * 2 + 1) * nthr] = zero4;
    }
    const float *bcp = sc + ((int64_t)min(b, n - 1) * n) * C + min(f, C - 4);
    auto fetch = [&](int c, int st) {
        const float *src = bcp + (int64_t)c * C;
        float4 *d = slot + (size_t)st * 2 * nthr;
        if (h2) cp16_cg(d, src + 2 * S.plane);
        if (own) cp16_cg(d + nthr, src + 12 * S.plane);
    };
    fetch(0, 0);
    cp_commit();
    if (n > 1) fetch(1, 1);
    cp_commit();
    cp_wait_group<1>();
    __syncthreads();
    if (!live) return;
    const int64_t slab = (int64_t)n * n * C;
    float4 g0[S4_TA];
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai)
        g0[ai] = h0 ? f4_ldg(sc + ((int64_t)min(a0 + ai, n - 1) * n + b) * C + f) : zero4;
    float *const dstb = a.T.base + inst * a.T.stride + (int64_t)a0 * slab + (int64_t)b * n * C + f;
    const bool accumulate = a.beta != 0.f;
    int st = 0, stn = 2;
#pragma unroll 1
    for (int c = 0; c < n; ++c) {
        if (c + 2 < n) fetch(c + 2, stn);
        cp_commit();
        cp_wait_group<2>();
        const float4 *d = slot + (size_t)st * 2 * nthr;
        const float4 c2 = d[0], c12 = d[nthr];
        st = st + 1 == S4_D ? 0 : st + 1;
        stn = stn + 1 == S4_D ? 0 : stn + 1;
        const float4 *acp = AC + (size_t)c * S4_Q + q;
#pragma unroll
        for (int ai = 0; ai < S4_TA; ++ai) {
            const int aa = a0 + ai;
            if (aa < n) {
                float4 v = acp[(size_t)ai * n * S4_Q];
                f4_add(v, g0[ai]);
                f4_add(v, c2);
                if (aa == b) f4_add(v, c12);
                float4 *dst = reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)c * C);
                if (accumulate) {
                    const float4 o = *dst;
                    v.x = fmaf(a.beta, o.x, v.x), v.y = fmaf(a.beta, o.y, v.y), v.z = fmaf(a.beta, o.z, v.z), v.w = fmaf(a.beta, o.w, v.w);
                }
                __stcs(dst, v);
            }
        }
    }
    if (!h13 && !h14) return;
    float4 d13[S4_TA], d14[S4_TA], o13[S4_TA], o14[S4_TA];
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai) {
        const int aa = min(a0 + ai, n - 1);
        const int64_t ab = ((int64_t)aa * n + b) * C + f;
        d13[ai] = h13 ? f4_ldg(sc + 13 * S.plane + ab) : zero4, d14[ai] = h14 ? f4_ldg(sc + 14 * S.plane + ab) : zero4;
        o13[ai] = __ldcg(reinterpret_cast<const float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C));
        o14[ai] = __ldcg(reinterpret_cast<const float4 *>(dstb + (int64_t)ai * slab + (int64_t)b * C));
    }
#pragma unroll
    for (int ai = 0; ai < S4_TA; ++ai) {
        const int aa = a0 + ai;
        if (aa < n) {
            if (aa == b) {
                f4_add(o13[ai], d13[ai]), f4_add(o13[ai], d14[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C), o13[ai]);
            } else {
                f4_add(o13[ai], d13[ai]), f4_add(o14[ai], d14[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)aa * C), o13[ai]);
                __stcs(reinterpret_cast<float4 *>(dstb + (int64_t)ai * slab + (int64_t)b * C), o14[ai]);
            }
        }
    }
}

inline unsigned blocks_for(int64_t elems) { return (unsigned)((elems + kThreads - 1) / kThreads); }

}  // namespace

cudaError_t r50_configure() {
    cudaError_t e = cudaSuccess;
#define R50_SMEM(fn)                                                                              \
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);       \
    if (e != cudaSuccess) return e;
    R50_SMEM(k_r50_fwd_out_tiled<1>) R50_SMEM(k_r50_fwd_out_tiled<2>) R50_SMEM(k_r50_fwd_out_tiled<4>)
    R50_SMEM((k_r50_bwd_planes_tiled<0, 1>)) R50_SMEM((k_r50_bwd_planes_tiled<0, 2>)) R50_SMEM((k_r50_bwd_planes_tiled<0, 4>))
    R50_SMEM((k_r50_bwd_planes_tiled<1, 1>)) R50_SMEM((k_r50_bwd_planes_tiled<1, 2>)) R50_SMEM((k_r50_bwd_planes_tiled<1, 4>))
#undef R50_SMEM
    e = cudaFuncSetAttribute(k_r50_fwd_out_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r50_out4_smem(V4_MAXN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_r50_bwd_planes_v4<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r50_planes4_smem(V4_MAXN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_r50_bwd_planes_v4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r50_planes4_smem(V4_MAXN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_r50_bwd_scatter_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r50_scatter4p_smem(S4_MAXN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_r50_bwd_scatter_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r50_scatter4_smem(S4_MAXN));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_r50_bwd_scatter_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

int r50_adj_words(int n_max) { return R50Adj{n_max}.words(); }
int64_t r50_scratch_words(int n_max, int C) { return R50Scratch(n_max, C).words; }

int r50_make_plan(int variant, uint64_t keep_mask, R50Plan *plan) {
    static const int sub18[18] = {1, 3, 5, 6, 10, 11, 13, 17, 18, 23, 26, 27, 28, 38, 40, 43, 46, 50};  // RisiContraction_18.h:102-318
    R50Plan &P = *plan;
    if (variant == 50 || variant == 10) {
        P.ncases = variant;  // RisiContraction_10 = cases 1..10 of the 50 (RisiContraction_10.h:91-140)
        for (int k = 0; k < P.ncases; ++k) P.c[k] = h_master[k];
    } else if (variant == 18) {
        P.ncases = 18;
        for (int k = 0; k < 18; ++k) P.c[k] = h_master[sub18[k] - 1];
    } else if (variant == 4) {
        // RisiContraction_4.h:79-118: sum_c T[a,b,c] -> (a,b); sum_a -> (b,c); T[a,a,c] -> (a,c); T[a,b,b] -> (a,b); no adjacency
        P.ncases = 4;
        const R50Case c4[4] = {{0, 0, 0, 0}, {0, 2, 0, 0}, {0, 12, 0, 0}, {0, 14, 0, 0}};
        for (int k = 0; k < 4; ++k) P.c[k] = c4[k];
    } else {
        return -1;
    }
    for (int k = 0; k < P.ncases; ++k)
        if (!((keep_mask >> k) & 1)) P.c[k].form = kFormOff;
    // pair the form-2 cases that read the same plane the same way with opposite orientations of A
    for (int k = 0; k < P.ncases; ++k) {
        if (P.c[k].form != 2 || (P.c[k].flags >> 2)) continue;
        for (int k2 = k + 1; k2 < P.ncases; ++k2) {
            const R50Case &c2 = P.c[k2];
            if (c2.form == 2 && c2.id == P.c[k].id && ((c2.flags ^ P.c[k].flags) & 3) == 2) {
                P.c[k].flags |= (unsigned char)((k2 + 1) << 2);
                P.c[k2].form = kFormFollower;
                break;
            }
        }
    }
    return 0;
}

cudaError_t launch_r50(bool backward, const R50Plan &plan, TensorRef T, float *out, int64_t stride_out, const float *adj,
                       int64_t stride_adj, Batch b, int adj_mode, float adj_scale, float *adjtab, float *scratch, float beta,
                       cudaStream_t st, LaunchLog *log) {
    const R50Adj AL{b.n_max};
    CCN_LAUNCH(log, K_R50_ADJ, st,
               k_r50_adj<<<b.count, 128, 0, st>>>(adj, stride_adj, b, adj_mode == 0 ? 1 : 0, adj_scale, adjtab, AL.words()));
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
    a.ncases = plan.ncases;
    for (int k = 0; k < kCases; ++k) a.plan[k] = plan.c[k];
    a.ntiles = a.nz = 0;
    for (int k = 0; k < plan.ncases; ++k) {
        const R50Case &cs = plan.c[k];
        if (r50_is2(cs.form)) {
            int t = 0;
            while (t < a.ntiles && !(a.tile_pid[t] == cs.id && a.tile_or[t] == (cs.flags & 1))) ++t;
            if (t == a.ntiles) {
                a.tile_pid[t] = cs.id, a.tile_or[t] = cs.flags & 1, a.tile_krow[t] = a.tile_kcol[t] = 0xff;
                ++a.ntiles;
            }
            ((cs.flags & 2) ? a.tile_kcol[t] : a.tile_krow[t]) = (unsigned char)k;  // A[j,y] : A[y,j]
        } else if (cs.form == 0) {
            a.z_k[a.nz] = (unsigned char)k, a.z_pid[a.nz] = cs.id, a.z_aux[a.nz] = cs.aux;
            ++a.nz;
        }
    }
    // planes a case needs, directly or through the vectors (V0, V1 <- plane 0; V2 <- 1; V3 <- 14; V4 <- 13; V5 <- 12) and the
    // scalars (X0 <- V0; X1 <- V5; X2 <- V4; X3 <- V3; X4 <- plane 12)
    unsigned need = 0;
    {
        const int vsrc[6] = {0, 0, 1, 14, 13, 12}, xsrc[5] = {0, 12, 13, 14, 12};
        for (int k = 0; k < plan.ncases; ++k) {
            const R50Case &cs = plan.c[k];
            if (cs.form == 0 || r50_is2(cs.form)) need |= 1u << cs.id;
            else if (cs.form == 1) need |= 1u << vsrc[cs.id];
            else if (cs.form == 3) need |= 1u << xsrc[cs.id];
        }
    }
    bool uses_vectors = false;
    for (int k = 0; k < plan.ncases; ++k) uses_vectors = uses_vectors || plan.c[k].form == 1 || plan.c[k].form == 3;
    static const bool no_plain = getenv("CCN_R50_NO_PLAIN") != nullptr;  // A/B knob
    const bool plain = (need & 0x0ff8u) == 0 && !no_plain;
    a.pm = 0x7fffu;
    bool v4_plan = true;  // the vector backward assumes the master table's shape: at most two form-0 cases on planes 0..2, one elsewhere, no form-2 tile on planes 3..11
    a.npt[0] = a.npt[1] = 0;
    for (int t = 0; t < a.ntiles; ++t) {
        const int o = a.tile_or[t];
        if (a.npt[o] == 5) { v4_plan = false; break; }
        a.pt_pid[o][a.npt[o]] = a.tile_pid[t], a.pt_krow[o][a.npt[o]] = a.tile_krow[t], a.pt_kcol[o][a.npt[o]] = a.tile_kcol[t];
        if (a.tile_pid[t] >= 3 && a.tile_pid[t] <= 11) v4_plan = false;
        ++a.npt[o];
    }
    for (int i = 0; i < 15; ++i) a.pz0[i] = a.pza0[i] = 0xff;
    for (int i = 0; i < 3; ++i) a.pz1[i] = a.pza1[i] = 0xff;
    for (int i = 0; i < a.nz; ++i) {
        const int pid = a.z_pid[i];
        if (a.pz0[pid] == 0xff) a.pz0[pid] = a.z_k[i], a.pza0[pid] = a.z_aux[i];
        else if (pid < 3 && a.pz1[pid] == 0xff) a.pz1[pid] = a.z_k[i], a.pza1[pid] = a.z_aux[i];
        else v4_plan = false;
    }
    const int64_t plane = (int64_t)b.n_max * b.n_max * b.C;
    dim3 grid(blocks_for(plane), b.count), grid3(blocks_for(plane * b.n_max), b.count);
    const int vthreads = b.C >= 256 ? 256 : ((b.C + 31) / 32) * 32;
    const int CB = b.C >= 128 ? 128 : ((b.C + 31) / 32) * 32;
    const size_t tile_bytes = r50_tile_bytes(b.n_max, CB);
    const size_t tile2_bytes = r50_tile2_bytes(b.n_max, CB);
    const bool tiled = tile2_bytes <= 200 * 1024;  // else the one-thread-per-element kernels
    dim3 gridt(b.n_max, b.count, (b.C + CB - 1) / CB);
    const bool vec4 = b.C % 4 == 0 && !T.slabs && ((uintptr_t)T.base & 15) == 0 && T.stride % 4 == 0 && ((uintptr_t)scratch & 15) == 0;
    // channels per thread in the plane x A kernels: 16-byte alignment of out / scratch rows needs C % 4 == 0
    const bool al16 = b.C % 4 == 0 && CB % 128 == 0 && ((uintptr_t)out & 15) == 0 && stride_out % 4 == 0 && ((uintptr_t)scratch & 15) == 0;
    int vw_f = al16 ? kR50VwFwd : 1, vw_b = al16 ? kR50VwBwd : 1;
    if (const char *ev = getenv("CCN_R50_VW")) {  // tuning knob: "<fwd><bwd>", e.g. 42
        if (al16 && (ev[0] == '1' || ev[0] == '2' || ev[0] == '4')) vw_f = ev[0] - '0';
        if (al16 && (ev[1] == '1' || ev[1] == '2' || ev[1] == '4')) vw_b = ev[1] - '0';
    }
    CCN_LAUNCH(log, K_R50_ADJ, st, k_r50_zero_scalars<<<b.count, kThreads, 0, st>>>(a));
    if (!backward) {
        if (plain) a.pm = need & 0x7007u;
        dim3 grid4(blocks_for(plane / 4), b.count);
        if (vec4 && plain)
            CCN_LAUNCH(log, K_R50_FWD_PLANES, st, k_r50_fwd_planes_plain<4><<<grid4, kThreads, 0, st>>>(a));
        else if (plain)
            CCN_LAUNCH(log, K_R50_FWD_PLANES, st, k_r50_fwd_planes_plain<1><<<grid, kThreads, 0, st>>>(a));
        else if (vec4)
            CCN_LAUNCH(log, K_R50_FWD_PLANES, st, k_r50_fwd_planes<4><<<grid4, kThreads, 0, st>>>(a));
        else
            CCN_LAUNCH(log, K_R50_FWD_PLANES, st, k_r50_fwd_planes<1><<<grid, kThreads, 0, st>>>(a));
        if (uses_vectors)  // no form-1 / form-3 case (RisiContraction_4): nothing reads the vectors or the scalars
            CCN_LAUNCH(log, K_R50_FWD_VECTORS, st, (k_r50_fwd_vectors<<<dim3(b.n_max, b.count), vthreads, 0, st>>>(a)));
        static const bool old_out = getenv("CCN_R50_OLD_OUT") != nullptr;  // A/B knob
        const bool v4 = b.C % 4 == 0 && b.n_max <= V4_MAXN && ((uintptr_t)out & 15) == 0 && stride_out % 4 == 0 && ((uintptr_t)scratch & 15) == 0 &&
                        ((uintptr_t)adjtab & 15) == 0;
        if (v4 && !old_out) {
            dim3 gridv(((b.C + V4_CB - 1) / V4_CB) * b.n_max, b.count);
            CCN_LAUNCH(log, K_R50_FWD_OUT, st, (k_r50_fwd_out_v4<<<gridv, b.n_max * V4_Q, r50_out4_smem(b.n_max), st>>>(a)));
        } else if (tiled && vw_f == 4)
            CCN_LAUNCH(log, K_R50_FWD_OUT, st, (k_r50_fwd_out_tiled<4><<<gridt, CB / 4, tile_bytes, st>>>(a, CB)));
        else if (tiled && vw_f == 2)
            CCN_LAUNCH(log, K_R50_FWD_OUT, st, (k_r50_fwd_out_tiled<2><<<gridt, CB / 2, tile_bytes, st>>>(a, CB)));
        else if (tiled)
            CCN_LAUNCH(log, K_R50_FWD_OUT, st, (k_r50_fwd_out_tiled<1><<<gridt, CB, tile_bytes, st>>>(a, CB)));
        else
            CCN_LAUNCH(log, K_R50_FWD_OUT, st, k_r50_fwd_out<<<grid, kThreads, 0, st>>>(a));
    } else {
        CCN_LAUNCH(log, K_R50_BWD_VECTORS, st, (k_r50_bwd_vectors<<<dim3(b.n_max, b.count), vthreads, 0, st>>>(a)));
        static const bool old_planes = getenv("CCN_R50_OLD_PLANES") != nullptr;  // A/B knob
        const bool v4b = v4_plan && b.C % 4 == 0 && b.n_max <= V4_MAXN && ((uintptr_t)out & 15) == 0 && stride_out % 4 == 0 &&
                         ((uintptr_t)scratch & 15) == 0 && ((uintptr_t)adjtab & 15) == 0;
        static const bool old_scatter = getenv("CCN_R50_OLD_SCATTER") != nullptr;  // A/B knob
        const bool v4s = vec4 && !old_scatter && b.n_max <= S4_MAXN;
        const bool plain_b = plain && v4b && !old_planes && v4s;  // both vector kernels, or every plane is materialised
        if (plain_b) a.pm = need & 0x7007u;
        if (v4b && !old_planes) {
            dim3 gridv(((b.C + V4_CB - 1) / V4_CB) * b.n_max, b.count);
            CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_v4<0><<<gridv, b.n_max * V4_Q, r50_planes4_smem(b.n_max), st>>>(a)));
            if (a.npt[1] > 0)
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_v4<1><<<gridv, b.n_max * V4_Q, r50_planes4_smem(b.n_max), st>>>(a)));
        } else if (tiled) {
            if (vw_b == 4) {
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<0, 4><<<gridt, CB / 4, tile2_bytes, st>>>(a, CB)));
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<1, 4><<<gridt, CB / 4, tile2_bytes, st>>>(a, CB)));
            } else if (vw_b == 2) {
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<0, 2><<<gridt, CB / 2, tile2_bytes, st>>>(a, CB)));
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<1, 2><<<gridt, CB / 2, tile2_bytes, st>>>(a, CB)));
            } else {
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<0, 1><<<gridt, CB, tile2_bytes, st>>>(a, CB)));
                CCN_LAUNCH(log, K_R50_BWD_PLANES, st, (k_r50_bwd_planes_tiled<1, 1><<<gridt, CB, tile2_bytes, st>>>(a, CB)));
            }
        } else {
            CCN_LAUNCH(log, K_R50_BWD_PLANES, st, k_r50_bwd_planes<<<grid, kThreads, 0, st>>>(a));
        }
        if (plain_b) {
            dim3 grids(((b.C + S4_CB - 1) / S4_CB) * ((b.n_max + S4_TA - 1) / S4_TA), b.count);
            CCN_LAUNCH(log, K_R50_BWD_SCATTER, st,
                       (k_r50_bwd_scatter_plain<<<grids, b.n_max * S4_Q, r50_scatter4p_smem(b.n_max), st>>>(a)));
        } else if (v4s) {
            dim3 grids(((b.C + S4_CB - 1) / S4_CB) * ((b.n_max + S4_TA - 1) / S4_TA), b.count);
            CCN_LAUNCH(log, K_R50_BWD_SCATTER, st,
                       (k_r50_bwd_scatter_v4<<<grids, b.n_max * S4_Q, r50_scatter4_smem(b.n_max), st>>>(a)));
        } else if (r50_scatter_smem(b.n_max) <= 100 * 1024) {
            dim3 grids((b.C + SC_CB - 1) / SC_CB, b.count, (b.n_max + SC_TA - 1) / SC_TA);
            CCN_LAUNCH(log, K_R50_BWD_SCATTER, st,
                       (k_r50_bwd_scatter_tiled<<<grids, SC_THREADS, r50_scatter_smem(b.n_max), st>>>(a)));
        } else {
            CCN_LAUNCH(log, K_R50_BWD_SCATTER, st, k_r50_bwd_scatter<<<grid3, kThreads, 0, st>>>(a));
        }
    }
    return cudaGetLastError();
}

}  // namespace ccn
