/*
 * oracle/ccn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the second-order CCN hot path of HyTruongSon/GraphFlow:
 *   StackTensor3D -> RisiContraction_18 (fwd+bwd) -> Reshape2D/MatMul -> VectorAddTensor -> LeakyReLU3D
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product path (graphflow_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this restatement against
 *   (1) golden vectors produced by the *unmodified* reference headers compiled here
 *       (oracle/_ref, recipe in oracle/Makefile, generator tests/golden/make_golden.py), and
 *   (2) oracle/_ref itself whenever that library is present.
 *
 * The contraction is restated as a table of index patterns instead of the reference's 18 hand-unrolled
 * loop nests; each row cites the reference line whose update it reproduces.  Loop order keeps (d,e)
 * outermost and skips adj <= 0 exactly where the reference does, so the operation count is the
 * reference's nnz(adj) * N^3 * C for cases 1-5 (GraphFlow/RisiContraction_18.h:86-125).
 *
 * Build twice: -DCCN_REAL=double -DCCN_SUF=f64 and -DCCN_REAL=float -DCCN_SUF=f32
 * (GraphFlow/ is the double tree, GraphFlow_32bit/ the float tree; same code otherwise).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef CCN_REAL
#define CCN_REAL double
#define CCN_SUF f64
#endif
#define CCN_CAT2(a, b) a##_##b
#define CCN_CAT(a, b) CCN_CAT2(a, b)
#define FN(name) CCN_CAT(name, CCN_SUF)

typedef CCN_REAL real;

/* Loop variables: 0=a 1=b 2=c (indices of T) 3=d 4=e (indices of adj). */
typedef struct {
    signed char t[3]; /* variable used at T's (row, column, depth-of-stack) positions */
    signed char m[2]; /* variable used at adj's (row, column) positions */
    signed char o[2]; /* variables that index the output (row, column) */
    short ref_line;   /* GraphFlow/RisiContraction_18.h line of the forward update */
} ccn_pattern;

/* The 18 contractions in the reference's slab order (slab k-1 lives at depth (k-1)*C + f). */
static const ccn_pattern PATTERNS18[18] = {
    {{0, 1, 2}, {3, 4}, {0, 1}, 102}, /*  1: keep a,b ; sum c,d,e                    */
    {{0, 1, 2}, {3, 4}, {0, 3}, 106}, /*  2: keep a,d ; sum b,c,e                    */
    {{0, 1, 2}, {3, 4}, {1, 2}, 110}, /*  3: keep b,c ; sum a,d,e                    */
    {{0, 1, 2}, {3, 4}, {1, 3}, 114}, /*  4: keep b,d ; sum a,c,e                    */
    {{0, 1, 2}, {3, 4}, {3, 4}, 118}, /*  5: keep d,e ; sum a,b,c                    */
    {{0, 1, 3}, {3, 4}, {0, 1}, 133}, /*  6: keep a,b ; c tied to d ; sum e          */
    {{0, 1, 2}, {3, 3}, {0, 1}, 149}, /*  7: keep a,b ; e tied to d ; sum c          */
    {{0, 1, 1}, {3, 4}, {0, 3}, 165}, /*  8: keep a,d ; c tied to b ; sum e          */
    {{0, 4, 2}, {3, 4}, {0, 3}, 180}, /*  9: keep a,d ; b tied to e ; sum c          */
    {{3, 1, 2}, {3, 4}, {1, 2}, 195}, /* 10: keep b,c ; a tied to d ; sum e          */
    {{0, 1, 0}, {3, 4}, {1, 3}, 211}, /* 11: keep b,d ; c tied to a ; sum e          */
    {{4, 1, 2}, {3, 4}, {1, 3}, 226}, /* 12: keep b,d ; a tied to e ; sum c          */
    {{0, 1, 4}, {3, 4}, {1, 3}, 241}, /* 13: keep b,d ; c tied to e ; sum a          */
    {{0, 0, 2}, {3, 4}, {3, 4}, 256}, /* 14: keep d,e ; b tied to a ; sum c          */
    {{0, 1, 1}, {3, 4}, {3, 4}, 271}, /* 15: keep d,e ; c tied to b ; sum a          */
    {{0, 4, 4}, {3, 4}, {0, 3}, 290}, /* 16: keep a,d ; b,c tied to e                */
    {{4, 1, 4}, {3, 4}, {1, 3}, 304}, /* 17: keep b,d ; a,c tied to e                */
    {{0, 0, 0}, {3, 4}, {3, 4}, 318}, /* 18: keep d,e ; a,b,c tied together          */
};

int FN(ccn_oracle_num_patterns)(void) { return 18; }
int FN(ccn_oracle_pattern_ref_line)(int k) { return (k >= 0 && k < 18) ? PATTERNS18[k].ref_line : -1; }

/*
 * One pattern, forward (dir = 0: out += T * adj) or backward (dir = 1: gT += gout * adj).
 *   T / gT  : [N, N, N, C]   index ((a*N + b)*N + c)*C + f      (StackTensor3D layout, StackTensor3D.h:54-66)
 *   adj     : [N, N]         index d*N + e                        (Matrix.h:34-36)
 *   out/gout: [N, N, 18*C]   index (x*N + y)*18C + k*C + f        (Tensor3D.h:37-39)
 * positive_part != 0 : skip adj <= 0 (RisiContraction_18.h:90, :345);
 * positive_part == 0 : multiply by the raw entry (RisiContraction_18_thread.h:70-72).
 */
static void run_pattern(const ccn_pattern *p, int k, int nslabs, int dir, real *T, const real *adj, real *out, int N,
                        int C, int positive_part) {
    int used[5] = {0, 0, 0, 0, 0};
    int lim[5], v[5];
    const size_t depth = (size_t)nslabs * C;
    for (int i = 0; i < 3; ++i) used[p->t[i]] = 1;
    for (int i = 0; i < 2; ++i) used[p->m[i]] = 1;
    for (int i = 0; i < 5; ++i) lim[i] = used[i] ? N : 1;

    for (v[3] = 0; v[3] < lim[3]; ++v[3]) {
        for (v[4] = 0; v[4] < lim[4]; ++v[4]) {
            const real w = adj[(size_t)v[p->m[0]] * N + v[p->m[1]]];
            if (positive_part && !(w > 0)) continue;
            for (v[0] = 0; v[0] < lim[0]; ++v[0]) {
                for (v[1] = 0; v[1] < lim[1]; ++v[1]) {
                    for (v[2] = 0; v[2] < lim[2]; ++v[2]) {
                        real *t = T + (((size_t)v[p->t[0]] * N + v[p->t[1]]) * N + v[p->t[2]]) * C;
                        real *o = out + ((size_t)v[p->o[0]] * N + v[p->o[1]]) * depth + (size_t)k * C;
                        if (dir == 0) {
                            for (int f = 0; f < C; ++f) o[f] += t[f] * w;
                        } else {
                            for (int f = 0; f < C; ++f) t[f] += o[f] * w;
                        }
                    }
                }
            }
        }
    }
}

/* RisiContraction_18::forward (RisiContraction_18.h:73-331): zero `out`, then accumulate all 18 slabs. */
void FN(ccn_oracle_contract18_forward)(const real *T, const real *adj, real *out, int N, int C, int positive_part) {
    memset(out, 0, sizeof(real) * (size_t)N * N * 18 * C);
    for (int k = 0; k < 18; ++k) run_pattern(&PATTERNS18[k], k, 18, 0, (real *)T, adj, out, N, C, positive_part);
}

/* RisiContraction_18::backward (RisiContraction_18.h:333-560): gT += transpose(gout).  Accumulates (never zeroes). */
void FN(ccn_oracle_contract18_backward)(const real *gout, const real *adj, real *gT, int N, int C, int positive_part) {
    for (int k = 0; k < 18; ++k) run_pattern(&PATTERNS18[k], k, 18, 1, gT, adj, (real *)gout, N, C, positive_part);
}

/* One slab only (used by tests to localise a mismatch). */
void FN(ccn_oracle_contract18_forward_case)(const real *T, const real *adj, real *out, int N, int C, int positive_part,
                                            int k) {
    run_pattern(&PATTERNS18[k], k, 18, 0, (real *)T, adj, out, N, C, positive_part);
}

/*
 * RisiContraction_50 (GraphFlow/RisiContraction_50.h:73-802): all 50 ways of keeping two of the five indices of
 * T[a,b,c] * adj[d,e] (config 5 of BASELINE.json).  Same pattern engine; the reference multiplies by the RAW
 * adjacency entry here (value_at, RisiContraction_50.h:63-65), so positive_part is 0.  ref_line is the `Case = k;`
 * line of the forward loop nest; the backward nest (:443-802) walks the same cases through set_gradient_for (:67-69).
 * The einsum in each comment is SURVEY.md Appendix A's machine-checked statement of the case.
 */
static const ccn_pattern PATTERNS50[50] = {
    {{0, 1, 2}, {3, 4}, {0, 1}, 95}, /*  1: abcf,de->abf */
    {{0, 1, 2}, {3, 4}, {0, 2}, 100}, /*  2: abcf,de->acf */
    {{0, 1, 2}, {3, 4}, {0, 3}, 105}, /*  3: abcf,de->adf */
    {{0, 1, 2}, {3, 4}, {0, 4}, 110}, /*  4: abcf,de->aef */
    {{0, 1, 2}, {3, 4}, {1, 2}, 115}, /*  5: abcf,de->bcf */
    {{0, 1, 2}, {3, 4}, {1, 3}, 120}, /*  6: abcf,de->bdf */
    {{0, 1, 2}, {3, 4}, {1, 4}, 125}, /*  7: abcf,de->bef */
    {{0, 1, 2}, {3, 4}, {2, 3}, 130}, /*  8: abcf,de->cdf */
    {{0, 1, 2}, {3, 4}, {2, 4}, 135}, /*  9: abcf,de->cef */
    {{0, 1, 2}, {3, 4}, {3, 4}, 140}, /* 10: abcf,de->def */
    {{0, 1, 3}, {3, 4}, {0, 1}, 149}, /* 11: abcf,ce->abf */
    {{0, 1, 4}, {3, 4}, {0, 1}, 156}, /* 12: abcf,dc->abf */
    {{0, 1, 2}, {3, 3}, {0, 1}, 163}, /* 13: abcf,dd->abf */
    {{0, 3, 2}, {3, 4}, {0, 2}, 170}, /* 14: abcf,be->acf */
    {{0, 4, 2}, {3, 4}, {0, 2}, 177}, /* 15: abcf,db->acf */
    {{0, 1, 2}, {3, 3}, {0, 2}, 184}, /* 16: abcf,dd->acf */
    {{0, 1, 1}, {3, 4}, {0, 3}, 191}, /* 17: abbf,de->adf */
    {{0, 4, 2}, {3, 4}, {0, 3}, 198}, /* 18: abcf,db->adf */
    {{0, 1, 4}, {3, 4}, {0, 3}, 205}, /* 19: abcf,dc->adf */
    {{0, 1, 1}, {3, 4}, {0, 4}, 212}, /* 20: abbf,de->aef */
    {{0, 3, 2}, {3, 4}, {0, 4}, 219}, /* 21: abcf,be->aef */
    {{0, 1, 3}, {3, 4}, {0, 4}, 226}, /* 22: abcf,ce->aef */
    {{3, 1, 2}, {3, 4}, {1, 2}, 233}, /* 23: abcf,ae->bcf */
    {{4, 1, 2}, {3, 4}, {1, 2}, 240}, /* 24: abcf,da->bcf */
    {{0, 1, 2}, {3, 3}, {1, 2}, 247}, /* 25: abcf,dd->bcf */
    {{0, 1, 0}, {3, 4}, {1, 3}, 254}, /* 26: abaf,de->bdf */
    {{4, 1, 2}, {3, 4}, {1, 3}, 261}, /* 27: abcf,da->bdf */
    {{0, 1, 4}, {3, 4}, {1, 3}, 268}, /* 28: abcf,dc->bdf */
    {{0, 1, 0}, {3, 4}, {1, 4}, 275}, /* 29: abaf,de->bef */
    {{3, 1, 2}, {3, 4}, {1, 4}, 282}, /* 30: abcf,ae->bef */
    {{0, 1, 3}, {3, 4}, {1, 4}, 289}, /* 31: abcf,ce->bef */
    {{0, 0, 2}, {3, 4}, {2, 3}, 296}, /* 32: aacf,de->cdf */
    {{4, 1, 2}, {3, 4}, {2, 3}, 303}, /* 33: abcf,da->cdf */
    {{0, 4, 2}, {3, 4}, {2, 3}, 310}, /* 34: abcf,db->cdf */
    {{0, 0, 2}, {3, 4}, {2, 4}, 317}, /* 35: aacf,de->cef */
    {{3, 1, 2}, {3, 4}, {2, 4}, 324}, /* 36: abcf,ae->cef */
    {{0, 3, 2}, {3, 4}, {2, 4}, 331}, /* 37: abcf,be->cef */
    {{0, 0, 2}, {3, 4}, {3, 4}, 338}, /* 38: aacf,de->def */
    {{0, 1, 0}, {3, 4}, {3, 4}, 345}, /* 39: abaf,de->def */
    {{0, 1, 1}, {3, 4}, {3, 4}, 352}, /* 40: abbf,de->def */
    {{0, 1, 3}, {3, 3}, {0, 1}, 363}, /* 41: abcf,cc->abf */
    {{0, 3, 2}, {3, 3}, {0, 2}, 370}, /* 42: abcf,bb->acf */
    {{0, 4, 4}, {3, 4}, {0, 3}, 377}, /* 43: abbf,db->adf */
    {{0, 3, 3}, {3, 4}, {0, 4}, 384}, /* 44: abbf,be->aef */
    {{3, 1, 2}, {3, 3}, {1, 2}, 391}, /* 45: abcf,aa->bcf */
    {{4, 1, 4}, {3, 4}, {1, 3}, 398}, /* 46: abaf,da->bdf */
    {{3, 1, 3}, {3, 4}, {1, 4}, 405}, /* 47: abaf,ae->bef */
    {{4, 4, 2}, {3, 4}, {2, 3}, 412}, /* 48: aacf,da->cdf */
    {{3, 3, 2}, {3, 4}, {2, 4}, 419}, /* 49: aacf,ae->cef */
    {{0, 0, 0}, {3, 4}, {3, 4}, 426}, /* 50: aaaf,de->def */
};

int FN(ccn_oracle_pattern50_ref_line)(int k) { return (k >= 0 && k < 50) ? PATTERNS50[k].ref_line : -1; }

void FN(ccn_oracle_contract50_forward)(const real *T, const real *adj, real *out, int N, int C) {
    memset(out, 0, sizeof(real) * (size_t)N * N * 50 * C);
    for (int k = 0; k < 50; ++k) run_pattern(&PATTERNS50[k], k, 50, 0, (real *)T, adj, out, N, C, 0);
}

/* gT += transpose(gout); accumulates like the reference's set_gradient_for. */
void FN(ccn_oracle_contract50_backward)(const real *gout, const real *adj, real *gT, int N, int C) {
    for (int k = 0; k < 50; ++k) run_pattern(&PATTERNS50[k], k, 50, 1, gT, adj, (real *)gout, N, C, 0);
}

/* StackTensor3D::forward (StackTensor3D.h:54-72): S[a,b,c,f] = tensors[a][b,c,f]. */
void FN(ccn_oracle_stack_forward)(const real *const *tensors, real *S, int N, int C) {
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) memcpy(S + a * slab, tensors[a], sizeof(real) * slab);
}

/* StackTensor3D::backward (StackTensor3D.h:74-90): tensors[a].gradient += S.gradient[a]. */
void FN(ccn_oracle_stack_backward)(const real *gS, real *const *gtensors, int N, int C) {
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a)
        for (size_t i = 0; i < slab; ++i) gtensors[a][i] += gS[a * slab + i];
}

/* MatMul::forward (MatMul.h:48-66): Y[M,P] = X[M,K] * W[K,P], ijk order, fresh output. */
void FN(ccn_oracle_matmul_forward)(const real *X, const real *W, real *Y, int M, int K, int P) {
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < P; ++j) {
            real acc = 0;
            for (int k = 0; k < K; ++k) acc += X[(size_t)i * K + k] * W[(size_t)k * P + j];
            Y[(size_t)i * P + j] = acc;
        }
}

/* MatMul::backward (MatMul.h:68-82): gX += gY * W^T ; gW += X^T * gY. */
void FN(ccn_oracle_matmul_backward)(const real *X, const real *W, const real *gY, real *gX, real *gW, int M, int K,
                                    int P) {
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < P; ++j) {
            const real g = gY[(size_t)i * P + j];
            for (int k = 0; k < K; ++k) {
                gX[(size_t)i * K + k] += g * W[(size_t)k * P + j];
                gW[(size_t)k * P + j] += g * X[(size_t)i * K + k];
            }
        }
}

/* VectorAddTensor::forward (VectorAddTensor.h:46-59) then LeakyReLU3D::forward (LeakyReLU3D.h:60-72). */
void FN(ccn_oracle_bias_lrelu_forward)(const real *Y, const real *bias, real *Z, int64_t rows, int P, real alpha) {
    for (int64_t i = 0; i < rows; ++i)
        for (int j = 0; j < P; ++j) {
            const real s = Y[i * P + j] + bias[j];
            Z[i * P + j] = (s > 0) ? s : alpha * s;
        }
}

/* LeakyReLU3D::backward (LeakyReLU3D.h:74-82) then VectorAddTensor::backward (VectorAddTensor.h:61-71):
 * gY += gZ * (pre > 0 ? 1 : alpha) ; gbias += column sums of that. `Y` is the pre-bias value. */
void FN(ccn_oracle_bias_lrelu_backward)(const real *Y, const real *bias, const real *gZ, real *gY, real *gbias,
                                        int64_t rows, int P, real alpha) {
    for (int64_t i = 0; i < rows; ++i)
        for (int j = 0; j < P; ++j) {
            const real s = Y[i * P + j] + bias[j];
            const real g = (s > 0) ? gZ[i * P + j] : gZ[i * P + j] * alpha;
            gY[i * P + j] += g;
            gbias[j] += g;
        }
}

/* TensorMul::forward (TensorMul.h:48-66): out[i,j,d] = sum_k A[i,k,d] * B[k,j,d]; fresh output. */
void FN(ccn_oracle_tensor_mul_forward)(const real *A, const real *B, real *out, int R, int K, int Cc, int D) {
    for (int d = 0; d < D; ++d)
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < Cc; ++j) {
                real acc = 0;
                for (int k = 0; k < K; ++k) acc += A[((size_t)i * K + k) * D + d] * B[((size_t)k * Cc + j) * D + d];
                out[((size_t)i * Cc + j) * D + d] = acc;
            }
}

/* TensorMul::backward (TensorMul.h:68-82): gA += g . B^T, gB += A^T . g, per channel. */
void FN(ccn_oracle_tensor_mul_backward)(const real *A, const real *B, const real *g, real *gA, real *gB, int R, int K, int Cc,
                                        int D) {
    for (int d = 0; d < D; ++d)
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < Cc; ++j) {
                const real gv = g[((size_t)i * Cc + j) * D + d];
                for (int k = 0; k < K; ++k) {
                    gA[((size_t)i * K + k) * D + d] += gv * B[((size_t)k * Cc + j) * D + d];
                    gB[((size_t)k * Cc + j) * D + d] += gv * A[((size_t)i * K + k) * D + d];
                }
            }
}

/* CustomMatMulTensor::forward (CustomMatMulTensor.h:47-63): Y[r,k] = sum_v Kt[k,v] * X[r,v], rows r = (i,j) flattened. */
void FN(ccn_oracle_custom_matmul_tensor_forward)(const real *Kt, const real *X, real *Y, int64_t rows, int V, int P) {
    for (int64_t r = 0; r < rows; ++r)
        for (int k = 0; k < P; ++k) {
            real acc = 0;
            for (int v = 0; v < V; ++v) acc += Kt[(size_t)k * V + v] * X[r * V + v];
            Y[r * P + k] = acc;
        }
}

/* CustomMatMulTensor::backward (CustomMatMulTensor.h:65-85): gKt[k,v] += gY[r,k] X[r,v]; gX[r,v] += gY[r,k] Kt[k,v]. */
void FN(ccn_oracle_custom_matmul_tensor_backward)(const real *Kt, const real *X, const real *gY, real *gKt, real *gX,
                                                  int64_t rows, int V, int P) {
    for (int64_t r = 0; r < rows; ++r)
        for (int k = 0; k < P; ++k) {
            const real g = gY[r * P + k];
            for (int v = 0; v < V; ++v) {
                gKt[(size_t)k * V + v] += g * X[r * V + v];
                gX[r * V + v] += g * Kt[(size_t)k * V + v];
            }
        }
}

/* Promotion X . f . X^T with a 0/1 selection X (MatTensorMul.h:47-65 then TensorMatMul.h:46-64 as wired at
 * SMP_beta.h:588-594; X from init_permutation_matrix, SMP_beta.h:446-459): Q[i,j,:] = f[pos[i], pos[j], :] when both
 * positions exist, else 0.  dir = 0 forward (fresh Q); dir = 1 backward (gf += selected gQ). */
void FN(ccn_oracle_promote)(real *f, const int *pos, real *Q, int n, int m, int C, int dir) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            real *q = Q + ((size_t)i * n + j) * C;
            if (pos[i] < 0 || pos[j] < 0) {
                if (dir == 0) memset(q, 0, sizeof(real) * C);
                continue;
            }
            real *src = f + ((size_t)pos[i] * m + pos[j]) * C;
            for (int c = 0; c < C; ++c) {
                if (dir == 0) q[c] = src[c];
                else src[c] += q[c];
            }
        }
}
