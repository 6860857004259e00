// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" shim around the UNMODIFIED reference headers of HyTruongSon/GraphFlow.  The reference sources are
// never copied into this repository: this file is compiled with -I/root/reference/GraphFlow (double tree) or
// -I/root/reference/GraphFlow_32bit (float tree) by oracle/Makefile, and the resulting shared objects go to
// oracle/_ref/ (git-ignored; they travel to the GPU box with the gpurun snapshot).
//
// What it exposes: the reference operators of the CCN hot path driven exactly the way the reference's own tests
// drive them (tests/test_RisiContraction_18_gpu.cu:90-135, 201-216), on caller-provided flat arrays.
//
// Compile-time switches:  -DGFREF_SUF=f64|f32  symbol suffix.
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <type_traits>
#include <vector>

#include "Matrix.h"
#include "Tensor3D.h"
#include "Tensor4D.h"
#include "RisiContraction_18.h"
#include "RisiContraction_18_thread.h"
#include "RisiContraction_50.h"
#include "RisiContraction_4.h"
#include "RisiContraction_10.h"
#include "RisiContraction_18_dropout.h"
#include "StackTensor3D.h"
#include "Reshape2D.h"
#include "MatMul.h"
#include "VectorAddTensor.h"
#include "LeakyReLU3D.h"
#include "TensorMul.h"
#include "CustomMatMulTensor.h"
#include "MatTensorMul.h"
#include "TensorMatMul.h"

#ifndef GFREF_SUF
#define GFREF_SUF f64
#endif
#define GF_CAT2(a, b) a##_##b
#define GF_CAT(a, b) GF_CAT2(a, b)
#define FN(name) GF_CAT(name, GFREF_SUF)

typedef std::remove_pointer<decltype(Vector::value)>::type real;

namespace {

struct Instance {
    int N, C;
    std::vector<Tensor3D *> tensors;
    Matrix *adj;
    Instance(int N_, int C_, const real *T, const real *A) : N(N_), C(C_) {
        const size_t slab = (size_t)N * N * C;
        for (int a = 0; a < N; ++a) {
            Tensor3D *t = new Tensor3D(N, N, C);
            if (T) std::memcpy(t->value, T + a * slab, sizeof(real) * slab);
            std::memset(t->gradient, 0, sizeof(real) * slab);  // the reference leaves new[] memory uninitialised
            tensors.push_back(t);
        }
        adj = new Matrix(N, N);
        std::memcpy(adj->value, A, sizeof(real) * N * N);
    }
    ~Instance() {
        for (size_t i = 0; i < tensors.size(); ++i) delete tensors[i];
        delete adj;
    }
};

template <class Op>
void wire(Op *op, Instance &in) {
    op->clear();
    for (int a = 0; a < in.N; ++a) op->add_tensor(in.tensors[a]);
    op->set_adjacency(in.adj);
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

}  // namespace

extern "C" {

int FN(gfref_sizeof_real)() { return (int)sizeof(real); }

// RisiContraction_18::forward (serial class).  T: [N,N,N,C] stacked, adj: [N,N], out: [N,N,18C].
void FN(gfref_contract18_forward)(const real *T, const real *adj, real *out, int N, int C) {
    Instance in(N, C, T, adj);
    RisiContraction_18 *op = new RisiContraction_18(N, C);
    wire(op, in);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    delete op;
}

// RisiContraction_18::backward.  gT is accumulated into (+=), as the reference does into tensors[a]->gradient.
void FN(gfref_contract18_backward)(const real *gout, const real *adj, real *gT, int N, int C) {
    Instance in(N, C, NULL, adj);
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) std::memcpy(in.tensors[a]->gradient, gT + a * slab, sizeof(real) * slab);
    RisiContraction_18 *op = new RisiContraction_18(N, C);
    wire(op, in);
    std::memcpy(op->gradient, gout, sizeof(real) * op->size);
    op->backward();
    for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
    delete op;
}

// The N^6 definition kept in the reference as DEPRECATED_forward (raw adj, explicit index guards).
void FN(gfref_contract18_forward_definition)(const real *T, const real *adj, real *out, int N, int C) {
    Instance in(N, C, T, adj);
    RisiContraction_18 *op = new RisiContraction_18(N, C);
    wire(op, in);
    op->DEPRECATED_forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    delete op;
}

// RisiContraction_18_thread::forward (6 std::threads; needs -DNDEBUG, its ctor asserts DEPRECATED == false).
void FN(gfref_contract18_thread_forward)(const real *T, const real *adj, real *out, int N, int C) {
#ifdef NDEBUG
    Instance in(N, C, T, adj);
    RisiContraction_18_thread *op = new RisiContraction_18_thread(N, C);
    wire(op, in);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    delete op;
#else
    (void)T; (void)adj; (void)out; (void)N; (void)C;
    std::abort();
#endif
}

// RisiContraction_4 (no adjacency, RisiContraction_4.h:68-180) and RisiContraction_10 (RisiContraction_10.h:72-230):
// forward, then backward of `gout` into a zero gradient.  out: [N,N,4C] / [N,N,10C]; gT: [N,N,N,C].
void FN(gfref_contract4)(const real *T, const real *gout, real *out, real *gT, int N, int C) {
    std::vector<real> eye((size_t)N * N, 0);
    Instance in(N, C, T, &eye[0]);
    RisiContraction_4 *op = new RisiContraction_4(N, C);
    op->clear();
    for (int a = 0; a < N; ++a) op->add_tensor(in.tensors[a]);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    std::memcpy(op->gradient, gout, sizeof(real) * op->size);
    op->backward();
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
    delete op;
}

void FN(gfref_contract10)(const real *T, const real *adj, const real *gout, real *out, real *gT, int N, int C) {
    Instance in(N, C, T, adj);
    RisiContraction_10 *op = new RisiContraction_10(N, C);
    wire(op, in);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    std::memcpy(op->gradient, gout, sizeof(real) * op->size);
    op->backward();
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
    delete op;
}

// RisiContraction_18_dropout (RisiContraction_18_dropout.h:104-797).  train != 0: srand(seed), then forward() draws the
// nKept kept slabs with rand() (:113-126); use_out[18] receives the mask it drew; backward of gout into a zero gradient.
// train == 0: test mode (all slabs, :128-131); the reference's backward asserts train mode (:485), so gT is left alone.
void FN(gfref_contract18_dropout)(const real *T, const real *adj, const real *gout, real *out, real *gT, int N, int C, int nKept,
                                  unsigned seed, int train, int *use_out) {
    Instance in(N, C, T, adj);
    RisiContraction_18_dropout *op = new RisiContraction_18_dropout(N, C);
    wire(op, in);
    op->setContractions(nKept);
    op->setMode(train != 0);
    srand(seed);
    op->forward();
    for (int k = 0; k < 18; ++k) use_out[k] = op->use[k] ? 1 : 0;
    std::memcpy(out, op->value, sizeof(real) * op->size);
    if (train) {
        std::memcpy(op->gradient, gout, sizeof(real) * op->size);
        op->backward();
        const size_t slab = (size_t)N * N * C;
        for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
    }
    delete op;
}

// RisiContraction_50::forward / backward (RisiContraction_50.h:73-441, 443-802): the N^6 loops, raw adjacency.
void FN(gfref_contract50_forward)(const real *T, const real *adj, real *out, int N, int C) {
    Instance in(N, C, T, adj);
    RisiContraction_50 *op = new RisiContraction_50(N, C);
    wire(op, in);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    delete op;
}

void FN(gfref_contract50_backward)(const real *gout, const real *adj, real *gT, int N, int C) {
    Instance in(N, C, NULL, adj);
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) std::memcpy(in.tensors[a]->gradient, gT + a * slab, sizeof(real) * slab);
    RisiContraction_50 *op = new RisiContraction_50(N, C);
    wire(op, in);
    std::memcpy(op->gradient, gout, sizeof(real) * op->size);
    op->backward();
    for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
    delete op;
}

// TensorMul::forward/backward (TensorMul.h:48-86): per-channel [R x K] . [K x Cc] products.  gA, gB accumulate.
void FN(gfref_tensor_mul)(const real *A, const real *B, real *out, const real *gout, real *gA, real *gB, int R, int K,
                          int Cc, int D) {
    Tensor3D *a = new Tensor3D(R, K, D), *b = new Tensor3D(K, Cc, D);
    std::memcpy(a->value, A, sizeof(real) * a->size);
    std::memcpy(b->value, B, sizeof(real) * b->size);
    TensorMul *op = new TensorMul(a, b);
    op->forward();
    std::memcpy(out, op->value, sizeof(real) * op->size);
    if (gout) {
        std::memcpy(a->gradient, gA, sizeof(real) * a->size);
        std::memcpy(b->gradient, gB, sizeof(real) * b->size);
        std::memcpy(op->gradient, gout, sizeof(real) * op->size);
        op->backward();
        std::memcpy(gA, a->gradient, sizeof(real) * a->size);
        std::memcpy(gB, b->gradient, sizeof(real) * b->size);
    }
    delete op; delete a; delete b;
}

// CustomMatMulTensor::forward/backward (CustomMatMulTensor.h:47-85): Y[i,j,k] = sum_v Kt[k,v] X[i,j,v].
void FN(gfref_custom_matmul_tensor)(const real *Kt, const real *X, real *Y, const real *gY, real *gKt, real *gX, int R, int Cc,
                                    int V, int P) {
    Matrix *k = new Matrix(P, V);
    Tensor3D *x = new Tensor3D(R, Cc, V);
    std::memcpy(k->value, Kt, sizeof(real) * k->size);
    std::memcpy(x->value, X, sizeof(real) * x->size);
    CustomMatMulTensor *op = new CustomMatMulTensor(k, x);
    op->forward();
    std::memcpy(Y, op->value, sizeof(real) * op->size);
    if (gY) {
        std::memcpy(k->gradient, gKt, sizeof(real) * k->size);
        std::memcpy(x->gradient, gX, sizeof(real) * x->size);
        std::memcpy(op->gradient, gY, sizeof(real) * op->size);
        op->backward();
        std::memcpy(gKt, k->gradient, sizeof(real) * k->size);
        std::memcpy(gX, x->gradient, sizeof(real) * x->size);
    }
    delete op; delete k; delete x;
}

// Promotion Q = X . f . X^T through MatTensorMul + TensorMatMul exactly as wired at SMP_beta.h:588-594, with the 0/1
// selection matrices of init_permutation_matrix (SMP_beta.h:446-459) built from pos[i] = column of the 1 in row i of X
// (or -1 for an all-zero row).  f: [m, m, C]; Q: [n, n, C]; backward adds into gf.
void FN(gfref_promote)(const real *f, const int *pos, real *Q, const real *gQ, real *gf, int n, int m, int C) {
    Matrix *X = new Matrix(n, m), *Xt = new Matrix(m, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            X->value[X->index(i, j)] = (pos[i] == j) ? 1 : 0;
            Xt->value[Xt->index(j, i)] = (pos[i] == j) ? 1 : 0;
        }
    std::memset(X->gradient, 0, sizeof(real) * X->size);
    std::memset(Xt->gradient, 0, sizeof(real) * Xt->size);
    Tensor3D *fw = new Tensor3D(m, m, C);
    std::memcpy(fw->value, f, sizeof(real) * fw->size);
    MatTensorMul *xf = new MatTensorMul(X, fw);
    TensorMatMul *q = new TensorMatMul(xf, Xt);
    xf->forward();
    q->forward();
    std::memcpy(Q, q->value, sizeof(real) * q->size);
    if (gQ) {
        std::memcpy(fw->gradient, gf, sizeof(real) * fw->size);
        std::memcpy(q->gradient, gQ, sizeof(real) * q->size);
        q->backward();
        xf->backward();
        std::memcpy(gf, fw->gradient, sizeof(real) * fw->size);
    }
    delete q; delete xf; delete fw; delete X; delete Xt;
}

// Stack -> contraction -> Reshape2D -> MatMul(K) -> Reshape3D is folded (flat copy) -> VectorAddTensor -> LeakyReLU3D,
// wired like SMP_beta.h:596-616.  Outputs the activation Z [N,N,Cout]; with gZ given also runs the whole backward
// chain and returns the tensors' gradients (stacked), gK and gb.  All gradient outputs are fresh (zeroed first).
void FN(gfref_level_forward_backward)(const real *T, const real *adj, const real *K, const real *bias, int N, int C,
                                      int Cout, real *contracted, real *Z, const real *gZ, real *gT, real *gK,
                                      real *gbias) {
    Instance in(N, C, T, adj);
    RisiContraction_18 *contract = new RisiContraction_18(N, C);
    wire(contract, in);
    Reshape2D *flat = new Reshape2D(contract, N * N, 18 * C);
    Matrix *Kmat = new Matrix(18 * C, Cout);
    std::memcpy(Kmat->value, K, sizeof(real) * 18 * C * Cout);
    Vector *b = new Vector(Cout);
    std::memcpy(b->value, bias, sizeof(real) * Cout);
    MatMul *mix = new MatMul(flat, Kmat);
    // Reshape3D is a flat copy [N*N, Cout] -> [N, N, Cout]; a Tensor3D view of the same numbers is equivalent.
    Tensor3D *mix3d = new Tensor3D(N, N, Cout);
    VectorAddTensor *biased = new VectorAddTensor(b, mix3d);
    LeakyReLU3D *act = new LeakyReLU3D(biased);

    contract->forward();
    flat->forward();
    Kmat->forward();
    b->forward();
    mix->forward();
    std::memcpy(mix3d->value, mix->value, sizeof(real) * mix->size);
    std::memset(mix3d->gradient, 0, sizeof(real) * mix3d->size);
    biased->forward();
    act->forward();
    if (contracted) std::memcpy(contracted, contract->value, sizeof(real) * contract->size);
    std::memcpy(Z, act->value, sizeof(real) * act->size);

    if (gZ) {
        std::memcpy(act->gradient, gZ, sizeof(real) * act->size);
        act->backward();
        biased->backward();
        for (int i = 0; i < mix->size; ++i) mix->gradient[i] += mix3d->gradient[i];
        mix->backward();
        flat->backward();
        contract->backward();
        const size_t slab = (size_t)N * N * C;
        for (int a = 0; a < N; ++a) std::memcpy(gT + a * slab, in.tensors[a]->gradient, sizeof(real) * slab);
        std::memcpy(gK, Kmat->gradient, sizeof(real) * Kmat->size);
        std::memcpy(gbias, b->gradient, sizeof(real) * Cout);
    }
    delete act; delete biased; delete mix3d; delete mix; delete b; delete Kmat; delete flat; delete contract;
}

// MatMul::forward/backward on flat arrays (MatMul.h:48-82).  gX, gW accumulate.
void FN(gfref_matmul_forward)(const real *X, const real *W, real *Y, int M, int K, int P) {
    Matrix *x = new Matrix(M, K), *w = new Matrix(K, P);
    std::memcpy(x->value, X, sizeof(real) * M * K);
    std::memcpy(w->value, W, sizeof(real) * K * P);
    MatMul *op = new MatMul(x, w);
    op->forward();
    std::memcpy(Y, op->value, sizeof(real) * M * P);
    delete op; delete x; delete w;
}

void FN(gfref_matmul_backward)(const real *X, const real *W, const real *gY, real *gX, real *gW, int M, int K, int P) {
    Matrix *x = new Matrix(M, K), *w = new Matrix(K, P);
    std::memcpy(x->value, X, sizeof(real) * M * K);
    std::memcpy(w->value, W, sizeof(real) * K * P);
    std::memcpy(x->gradient, gX, sizeof(real) * M * K);
    std::memcpy(w->gradient, gW, sizeof(real) * K * P);
    MatMul *op = new MatMul(x, w);
    std::memcpy(op->gradient, gY, sizeof(real) * M * P);
    op->backward();
    std::memcpy(gX, x->gradient, sizeof(real) * M * K);
    std::memcpy(gW, w->gradient, sizeof(real) * K * P);
    delete op; delete x; delete w;
}

// Replica-parallel timing of the reference's real data-parallel scheme (SMP_beta.h:697-739: one private replica per
// host thread): `threads` host threads each run `reps` x (RisiContraction_18::forward + backward) on a private
// instance built from the same T/adj.  Returns wall seconds for the whole fan-out; contractions done = threads*reps.
double FN(gfref_contract18_time_replicas)(const real *T, const real *adj, const real *gout, int N, int C, int threads,
                                          int reps) {
    std::vector<Instance *> inst(threads);
    std::vector<RisiContraction_18 *> ops(threads);
    for (int t = 0; t < threads; ++t) {
        inst[t] = new Instance(N, C, T, adj);
        ops[t] = new RisiContraction_18(N, C);
        wire(ops[t], *inst[t]);
    }
    const double t0 = now_s();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        RisiContraction_18 *op = ops[t];
        pool.push_back(std::thread([op, gout, reps]() {
            for (int r = 0; r < reps; ++r) {
                op->forward();
                std::memcpy(op->gradient, gout, sizeof(real) * op->size);
                op->backward();
            }
        }));
    }
    for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
    const double t1 = now_s();
    for (int t = 0; t < threads; ++t) {
        delete ops[t];
        delete inst[t];
    }
    return t1 - t0;
}

// Wall seconds of `reps` x (RisiContraction_18_thread::forward + backward) -- the 6-threads-per-op variant that
// BASELINE.json's north_star names (RisiContraction_18_thread.h:79-781: N^6 loops, 3 cases per thread).  Its backward is
// racy in the reference (several threads += into the same gradients), so only the time is meaningful here.
double FN(gfref_contract18_thread_time)(const real *T, const real *adj, const real *gout, int N, int C, int reps) {
#ifdef NDEBUG
    Instance in(N, C, T, adj);
    RisiContraction_18_thread *op = new RisiContraction_18_thread(N, C);
    wire(op, in);
    const double t0 = now_s();
    for (int r = 0; r < reps; ++r) {
        op->forward();
        std::memcpy(op->gradient, gout, sizeof(real) * op->size);
        op->backward();
    }
    const double t1 = now_s();
    delete op;
    return t1 - t0;
#else
    (void)T; (void)adj; (void)gout; (void)N; (void)C; (void)reps;
    return -1.0;
#endif
}

}  // extern "C"
