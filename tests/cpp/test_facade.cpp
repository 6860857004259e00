// tests/cpp/test_facade.cpp -- parity of the header-compatible C++ operators (include/graphflow_b200/ccn_ops_b200.h)
// against the UNMODIFIED reference CPU operators, in the reference's own host language and through its own
// Entity / setParameter / forward() / backward() API.  Built by tests/cpp/Makefile once per reference tree
// (-I/root/reference/GraphFlow_32bit -> _build/test_facade_f32, -I/root/reference/GraphFlow -> _build/test_facade_f64);
// the binaries travel to the GPU box and are run by tests/test_facade_gpu.py.
//
// Scenarios (each prints one line `name key=value ...` and the program exits non-zero if any check fails):
//   contract N C   the procedure of the reference's tests/test_RisiContraction_18_gpu.cu:80-231 (seed 123456789,
//                  symmetric rand()%10 tensors, I + random symmetric 0/1 adjacency, rand()%100 output gradients):
//                  ccn_b200::StackTensor3D + ccn_b200::RisiContraction_18_gpu  vs  ::RisiContraction_18
//   hostapi N C    ccn_b200::RisiContraction_18 (add_tensor API) and RisiContraction_18_gpu on a host Tensor4D built
//                  by the reference's own ::StackTensor3D, real-valued inputs, non-zero initial input gradients (+=)
//   matmul         the procedure of tests/test_MatMul_gpu.cu:22-26,54-60,103-116 (1600x720 . 720x40, rand()%100,
//                  non-zero initial gradients): ccn_b200::MatMul_gpu vs ::MatMul
//   r50 N C        ccn_b200::RisiContraction_50 vs ::RisiContraction_50 (N^6 reference loops; keep N small)
//   family N C     ccn_b200::RisiContraction_4 / _10 / _18_dropout vs the reference classes of the same name; the dropout
//                  operator is run in train mode from the same srand() seed on both sides (identical use[] draws,
//                  RisiContraction_18_dropout.h:113-126), forward + backward, and in test mode (nKept/18 scaling)
//   aux N C        ccn_b200::TensorMul and ccn_b200::CustomMatMulTensor vs the reference classes of the same name
//   batch C P      ccn_b200::LevelBatch: six vertices with different receptive-field sizes in one launch set vs six
//                  independent reference chains sharing K and b
//   threads N C    the reference's multi-stream replica scheme (GraphFlow_gpu/SMP_beta_gpu_multistreams.h:701-718): four
//                  host threads, each with its own stream and its own (thread-local) context, run
//                  RisiContraction_18_gpu forward + backward on different inputs concurrently and repeatedly; every
//                  thread's results must equal the ones the same operator produced single-threaded
//   level N C P    ccn_b200::CCNLevel vs the reference chain StackTensor3D -> RisiContraction_18 -> Reshape2D ->
//                  MatMul -> Reshape3D -> VectorAddTensor -> LeakyReLU3D (SMP_beta.h:600-616), both driven through
//                  ccn_b200::Executor
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "Matrix.h"  // RisiContraction_18.h uses Matrix without including it
#include "RisiContraction_18.h"
#include "StackTensor3D.h"
#include "MatMul.h"
#include "Reshape2D.h"
#include "Reshape3D.h"
#include "VectorAddTensor.h"
#include "LeakyReLU3D.h"
#include "TensorMul.h"
#include "RisiContraction_50.h"
#include "RisiContraction_4.h"
#include "RisiContraction_10.h"
#include "RisiContraction_18_dropout.h"
#include "CustomMatMulTensor.h"

#include "graphflow_b200/ccn_ops_b200.h"

typedef ccn_b200::real real;

static int failures = 0;

static double max_abs(const real *x, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; ++i) m = std::max(m, (double)std::fabs(x[i]));
    return m;
}
static double max_diff(const real *x, const real *y, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; ++i) m = std::max(m, (double)std::fabs(x[i] - y[i]));
    return m;
}
// slab-normalised error of an [cells, slabs*C] array: max_k max|x-ref| / max|ref_k|
static double slab_err(const real *x, const real *ref, size_t cells, int slabs, int C) {
    double worst = 0;
    for (int k = 0; k < slabs; ++k) {
        double num = 0, den = 0;
        for (size_t i = 0; i < cells; ++i)
            for (int f = 0; f < C; ++f) {
                const size_t j = (i * slabs + k) * C + f;
                num = std::max(num, (double)std::fabs(x[j] - ref[j]));
                den = std::max(den, (double)std::fabs(ref[j]));
            }
        worst = std::max(worst, den > 0 ? num / den : num);
    }
    return worst;
}
static void check(const char *name, const char *what, double err, double tol) {
    const bool ok = err <= tol;
    std::printf("%s %s err=%.3e tol=%.1e %s\n", name, what, err, tol, ok ? "ok" : "FAIL");
    if (!ok) ++failures;
}
static double uniform() { return 2.0 * (rand() / (RAND_MAX + 1.0)) - 1.0; }

static void scenario_contract(int N, int C) {
    srand(123456789);
    ccn_b200::RisiContraction_18_gpu *contract = new ccn_b200::RisiContraction_18_gpu(N, C);
    RisiContraction_18 *truth = new RisiContraction_18(N, C);
    ccn_b200::StackTensor3D *stack = new ccn_b200::StackTensor3D(N, N, N, C);
    std::vector<Tensor3D *> tensors(N);
    for (int i = 0; i < N; ++i) {
        tensors[i] = new Tensor3D(N, N, C);
        for (int f = 0; f < C; ++f)
            for (int r = 0; r < N; ++r)
                for (int c = r; c < N; ++c) {
                    const int v = rand() % 10;
                    tensors[i]->value[tensors[i]->index(r, c, f)] = v;
                    tensors[i]->value[tensors[i]->index(c, r, f)] = v;
                }
        std::memset(tensors[i]->gradient, 0, sizeof(real) * tensors[i]->size);  // new[] leaves it uninitialised
    }
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < N; ++i) {
        adj->value[adj->index(i, i)] = 1;
        for (int j = i + 1; j < N; ++j) {
            const int v = rand() % 2;
            adj->value[adj->index(i, j)] = v;
            adj->value[adj->index(j, i)] = v;
        }
    }
    stack->clear();
    truth->clear();
    for (int i = 0; i < N; ++i) {
        stack->add_tensor(tensors[i]);
        truth->add_tensor(tensors[i]);
    }
    contract->setParameter(stack, adj);
    truth->set_adjacency(adj);
    if (contract->size != truth->size || contract->nRows != truth->nRows || contract->nDepth != truth->nDepth) ++failures;

    stack->forward();
    contract->forward();
    truth->forward();
    const bool exact = max_abs(truth->value, truth->size) < 16777216.0;  // integers below 2^24 are exact in fp32
    check("contract", "forward", slab_err(contract->value, truth->value, (size_t)N * N, 18, C), exact ? 0.0 : 1e-4);
    check("contract", "forward_zeroes_gradient", max_abs(contract->gradient, contract->size), 0.0);

    for (int i = 0; i < truth->size; ++i) {
        truth->gradient[i] = rand() % 100;
        contract->gradient[i] = truth->gradient[i];
    }
    truth->backward();  // += into tensors[a]->gradient (zero before)
    std::vector<real> want((size_t)N * N * N * C);
    const size_t slab = (size_t)N * N * C;
    for (int a = 0; a < N; ++a) {
        std::memcpy(&want[a * slab], tensors[a]->gradient, sizeof(real) * slab);
        std::memset(tensors[a]->gradient, 0, sizeof(real) * slab);
    }
    contract->backward();  // accumulates on the device
    stack->backward();     // += into tensors[a]->gradient
    std::vector<real> got((size_t)N * N * N * C);
    for (int a = 0; a < N; ++a) std::memcpy(&got[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    const double scale = max_abs(&want[0], want.size());
    check("contract", "backward", max_diff(&got[0], &want[0], want.size()) / scale, scale < 16777216.0 ? 0.0 : 1e-4);
    contract->release();
    stack->release();
}

static void scenario_hostapi(int N, int C) {
    srand(2024);
    std::vector<Tensor3D *> tensors(N);
    const size_t slab = (size_t)N * N * C;
    std::vector<real> g0(slab * N);
    for (int i = 0; i < N; ++i) {
        tensors[i] = new Tensor3D(N, N, C);
        for (size_t j = 0; j < slab; ++j) {
            tensors[i]->value[j] = uniform();
            g0[i * slab + j] = uniform();
        }
    }
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) {
            const real v = (i == j) ? 1 : ((rand() % 4 == 0) ? 1 : 0);
            adj->value[adj->index(i, j)] = v;
            adj->value[adj->index(j, i)] = v;
        }
    RisiContraction_18 *truth = new RisiContraction_18(N, C);
    ccn_b200::RisiContraction_18 *ours = new ccn_b200::RisiContraction_18(N, C);
    StackTensor3D *ref_stack = new StackTensor3D(N, N, N, C);  // the reference's own host stack
    ccn_b200::RisiContraction_18_gpu *gpu = new ccn_b200::RisiContraction_18_gpu(N, C);
    for (int i = 0; i < N; ++i) {
        truth->add_tensor(tensors[i]);
        ours->add_tensor(tensors[i]);
        ref_stack->add_tensor(tensors[i]);
    }
    truth->set_adjacency(adj);
    ours->set_adjacency(adj);
    ref_stack->forward();
    gpu->setParameter(static_cast<Tensor4D *>(ref_stack), adj);
    truth->forward();
    ours->forward();
    gpu->forward();
    check("hostapi", "forward_add_tensor_api", slab_err(ours->value, truth->value, (size_t)N * N, 18, C), 1e-4);
    check("hostapi", "forward_tensor4d_api", slab_err(gpu->value, truth->value, (size_t)N * N, 18, C), 1e-4);

    for (int i = 0; i < truth->size; ++i) truth->gradient[i] = ours->gradient[i] = gpu->gradient[i] = uniform();
    std::vector<real> want(slab * N), got(slab * N);
    // reference
    for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
    truth->backward();
    for (int a = 0; a < N; ++a) std::memcpy(&want[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    const double scale = max_abs(&want[0], want.size());
    // add_tensor API, += on top of the same initial gradients
    for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
    ours->backward();
    for (int a = 0; a < N; ++a) std::memcpy(&got[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    check("hostapi", "backward_add_tensor_api", max_diff(&got[0], &want[0], want.size()) / scale, 1e-4);
    // Tensor4D API: += into ref_stack->gradient, then the reference's StackTensor3D::backward
    for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
    std::memset(ref_stack->gradient, 0, sizeof(real) * ref_stack->size);
    gpu->backward();
    ref_stack->backward();
    for (int a = 0; a < N; ++a) std::memcpy(&got[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    check("hostapi", "backward_tensor4d_api", max_diff(&got[0], &want[0], want.size()) / scale, 1e-4);
    ours->release();
    gpu->release();
}

static void scenario_matmul() {
    const int Ar = 1600, Ac = 720, Bc = 40;  // the N=40, C=40 feature-mix shape of tests/test_MatMul_gpu.cu:22-26
    srand(123456789);
    Matrix *A = new Matrix(Ar, Ac), *B = new Matrix(Ac, Bc);
    for (int i = 0; i < A->size; ++i) A->value[i] = rand() % 100;
    for (int i = 0; i < B->size; ++i) B->value[i] = rand() % 100;
    ccn_b200::MatMul_gpu *obj = new ccn_b200::MatMul_gpu(A, B);
    MatMul *truth = new MatMul(A, B);
    obj->forward();
    truth->forward();
    // integer inputs, every partial sum below 2^24: exact in fp32 whatever the summation order
    check("matmul", "forward", max_diff(obj->value, truth->value, truth->size), 0.0);
    std::vector<real> a0(A->size), b0(B->size), ga(A->size), gb(B->size);
    for (int i = 0; i < obj->size; ++i) obj->gradient[i] = truth->gradient[i] = rand() % 10;
    for (int i = 0; i < A->size; ++i) a0[i] = rand() % 100;
    for (int i = 0; i < B->size; ++i) b0[i] = rand() % 100;
    std::memcpy(A->gradient, &a0[0], sizeof(real) * A->size);
    std::memcpy(B->gradient, &b0[0], sizeof(real) * B->size);
    obj->backward();
    std::memcpy(&ga[0], A->gradient, sizeof(real) * A->size);
    std::memcpy(&gb[0], B->gradient, sizeof(real) * B->size);
    std::memcpy(A->gradient, &a0[0], sizeof(real) * A->size);
    std::memcpy(B->gradient, &b0[0], sizeof(real) * B->size);
    truth->backward();
    check("matmul", "backward_first", max_diff(&ga[0], A->gradient, A->size), 0.0);
    check("matmul", "backward_second", max_diff(&gb[0], B->gradient, B->size), 0.0);
    obj->release();
}

static void scenario_level(int N, int C, int P) {
    srand(777);
    const size_t slab = (size_t)N * N * C;
    std::vector<Tensor3D *> tensors(N);
    for (int i = 0; i < N; ++i) {
        tensors[i] = new Tensor3D(N, N, C);
        for (size_t j = 0; j < slab; ++j) tensors[i]->value[j] = uniform();
    }
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) {
            const real v = (i == j) ? 1 : ((rand() % 4 == 0) ? 1 : 0);
            adj->value[adj->index(i, j)] = v;
            adj->value[adj->index(j, i)] = v;
        }
    Matrix *K = new Matrix(18 * C, P);
    Vector *b = new Vector(P);
    for (int i = 0; i < K->size; ++i) K->value[i] = 0.05 * uniform();
    for (int i = 0; i < b->size; ++i) b->value[i] = 0.5 * uniform();

    // reference chain, SMP_beta.h:600-616, driven through the executor's RefOp adapters
    StackTensor3D *r_stack = new StackTensor3D(N, N, N, C);  // not part of SMP_beta's chain; harmless here
    RisiContraction_18 *r_con = new RisiContraction_18(N, C);
    for (int i = 0; i < N; ++i) r_con->add_tensor(tensors[i]);
    r_con->set_adjacency(adj);
    Reshape2D *r_2d = new Reshape2D(r_con, N * N, 18 * C);
    MatMul *r_mm = new MatMul(r_2d, K);
    Reshape3D *r_3d = new Reshape3D(r_mm, N, N, P);
    VectorAddTensor *r_add = new VectorAddTensor(b, r_3d);
    LeakyReLU3D *r_act = new LeakyReLU3D(r_add);
    (void)r_stack;
    ccn_b200::Executor ref;
    ref.add(r_con, 40);
    ref.add(r_2d);
    ref.add(r_mm);
    ref.add(r_3d);
    ref.add(r_add);
    ref.add(r_act);

    ccn_b200::CCNLevel *lvl = new ccn_b200::CCNLevel(N, C, P);
    lvl->setParameter(N, C, K, b);
    for (int i = 0; i < N; ++i) lvl->add_tensor(tensors[i]);
    lvl->set_adjacency(adj);
    ccn_b200::Executor mine;
    mine.add(lvl, ccn_b200::CCNLEVEL_B200);

    ref.forward();
    mine.forward();
    check("level", "forward", max_diff(lvl->value, r_act->value, r_act->size) / max_abs(r_act->value, r_act->size), 1e-4);

    std::vector<real> gz(r_act->size);
    for (size_t i = 0; i < gz.size(); ++i) gz[i] = uniform();
    std::vector<real> want_T(slab * N), want_K(K->size), want_b(b->size);
    // reference backward (parameters and inputs start from zero gradients)
    for (int a = 0; a < N; ++a) std::memset(tensors[a]->gradient, 0, sizeof(real) * slab);
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(b->gradient, 0, sizeof(real) * b->size);
    std::memcpy(r_act->gradient, &gz[0], sizeof(real) * gz.size());
    ref.backward();
    for (int a = 0; a < N; ++a) std::memcpy(&want_T[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    std::memcpy(&want_K[0], K->gradient, sizeof(real) * K->size);
    std::memcpy(&want_b[0], b->gradient, sizeof(real) * b->size);
    // ours
    for (int a = 0; a < N; ++a) std::memset(tensors[a]->gradient, 0, sizeof(real) * slab);
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(b->gradient, 0, sizeof(real) * b->size);
    std::memcpy(lvl->gradient, &gz[0], sizeof(real) * gz.size());
    mine.backward();
    std::vector<real> got_T(slab * N);
    for (int a = 0; a < N; ++a) std::memcpy(&got_T[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    check("level", "backward_tensors", max_diff(&got_T[0], &want_T[0], want_T.size()) / max_abs(&want_T[0], want_T.size()), 1e-4);
    check("level", "backward_K", max_diff(K->gradient, &want_K[0], want_K.size()) / max_abs(&want_K[0], want_K.size()), 1e-4);
    check("level", "backward_b", max_diff(b->gradient, &want_b[0], want_b.size()) / max_abs(&want_b[0], want_b.size()), 1e-4);
    lvl->release();
}

// Four replicas on four host threads and four streams, as in SMP_beta_gpu_multistreams.h:701-718.
struct Replica {
    Tensor4D *T;
    Matrix *adj;
    ccn_b200::RisiContraction_18_gpu *op;
    std::vector<real> gout, want_out, want_gT;
    double err_out, err_gT;
    int status;
};
static void replica_run(Replica *r, int rounds) {
    ccn_ctx *ctx = ccn_b200::context();  // this thread's own context
    void *st = NULL;
    if (ccn_stream_create(ctx, &st) != CCN_OK) {
        r->status = 1;
        return;
    }
    r->op->set_gpu_stream((cudaStream_t)st);
    r->err_out = r->err_gT = 0;
    for (int it = 0; it < rounds; ++it) {
        std::memset(r->T->gradient, 0, sizeof(real) * r->T->size);
        r->op->forward();
        r->err_out = std::max(r->err_out, max_diff(r->op->value, &r->want_out[0], r->want_out.size()));
        std::memcpy(r->op->gradient, &r->gout[0], sizeof(real) * r->gout.size());
        r->op->backward();
        r->err_gT = std::max(r->err_gT, max_diff(r->T->gradient, &r->want_gT[0], r->want_gT.size()));
    }
    r->op->release();
    ccn_stream_destroy(ctx, st);
    r->status = 0;
}
static void scenario_threads(int N, int C) {
    const int R = 4;
    std::vector<Replica> reps(R);
    for (int i = 0; i < R; ++i) {
        srand(1000 + i);
        Replica &r = reps[i];
        r.T = new Tensor4D(N, N, N, C);
        for (int j = 0; j < r.T->size; ++j) r.T->value[j] = uniform();
        r.adj = new Matrix(N, N);
        for (int a = 0; a < N; ++a)
            for (int b = a; b < N; ++b) {
                const real v = (a == b) ? 1 : ((rand() % 3 == 0) ? 1 : 0);
                r.adj->value[r.adj->index(a, b)] = v;
                r.adj->value[r.adj->index(b, a)] = v;
            }
        r.op = new ccn_b200::RisiContraction_18_gpu(r.T, r.adj);
        r.gout.resize(r.op->size);
        for (size_t j = 0; j < r.gout.size(); ++j) r.gout[j] = uniform();
        // single-threaded pass on the main thread (default stream): the expected results
        std::memset(r.T->gradient, 0, sizeof(real) * r.T->size);
        r.op->forward();
        r.want_out.assign(r.op->value, r.op->value + r.op->size);
        std::memcpy(r.op->gradient, &r.gout[0], sizeof(real) * r.gout.size());
        r.op->backward();
        r.want_gT.assign(r.T->gradient, r.T->gradient + r.T->size);
        r.op->release();  // the device buffers belong to the main thread's context; the replica thread makes its own
    }
    std::vector<std::thread> pool;
    for (int i = 0; i < R; ++i) pool.push_back(std::thread(replica_run, &reps[i], 8));
    for (int i = 0; i < R; ++i) pool[i].join();
    double eo = 0, eg = 0;
    int bad = 0;
    for (int i = 0; i < R; ++i) {
        eo = std::max(eo, reps[i].err_out / max_abs(&reps[i].want_out[0], reps[i].want_out.size()));
        eg = std::max(eg, reps[i].err_gT / max_abs(&reps[i].want_gT[0], reps[i].want_gT.size()));
        bad += reps[i].status;
    }
    check("threads", "stream_create", bad, 0);
    check("threads", "forward_vs_single_thread", eo, 1e-6);
    check("threads", "backward_vs_single_thread", eg, 1e-6);
}

// One member of the family against the reference class of the same name: forward, then backward `+=` into non-zero
// input gradients.  `prepare` runs on both operators before each forward (seeding, modes).
template <class Ref, class Ours, class Prep>
static void family_case(const char *name, int N, int C, bool with_adj, bool do_backward, Prep prepare) {
    std::vector<Tensor3D *> tensors(N);
    const size_t slab = (size_t)N * N * C;
    std::vector<real> g0(slab * N);
    for (int i = 0; i < N; ++i) {
        tensors[i] = new Tensor3D(N, N, C);
        for (size_t j = 0; j < slab; ++j) {
            tensors[i]->value[j] = uniform();
            g0[i * slab + j] = uniform();
        }
    }
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < adj->size; ++i) adj->value[i] = uniform();
    Ref *truth = new Ref(N, C);
    Ours *ours = new Ours(N, C);
    for (int i = 0; i < N; ++i) {
        truth->add_tensor(tensors[i]);
        ours->add_tensor(tensors[i]);
    }
    prepare(truth, ours, adj);
    srand(4242);
    truth->forward();
    srand(4242);
    ours->forward();
    const int S = Ours::nContractions;
    std::string tag = std::string(name) + "_forward";
    check("family", tag.c_str(), max_diff(ours->value, truth->value, truth->size) / max_abs(truth->value, truth->size), 1e-4);
    int zeros_ok = 1;  // slabs the reference left at exactly zero must be exactly zero here too (dropped slabs)
    for (int k = 0; k < S; ++k) {
        bool ref_zero = true, our_zero = true;
        for (size_t cell = 0; cell < (size_t)N * N; ++cell)
            for (int f = 0; f < C; ++f) {
                ref_zero = ref_zero && truth->value[(cell * S + k) * C + f] == 0;
                our_zero = our_zero && ours->value[(cell * S + k) * C + f] == 0;
            }
        if (ref_zero != our_zero) zeros_ok = 0;
    }
    tag = std::string(name) + "_dropped_slabs_exact";
    check("family", tag.c_str(), zeros_ok ? 0 : 1, 0);
    if (do_backward) {
        for (int i = 0; i < truth->size; ++i) truth->gradient[i] = ours->gradient[i] = uniform();
        std::vector<real> want(slab * N), got(slab * N);
        for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
        truth->backward();
        for (int a = 0; a < N; ++a) std::memcpy(&want[a * slab], tensors[a]->gradient, sizeof(real) * slab);
        for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
        ours->backward();
        for (int a = 0; a < N; ++a) std::memcpy(&got[a * slab], tensors[a]->gradient, sizeof(real) * slab);
        tag = std::string(name) + "_backward";
        check("family", tag.c_str(), max_diff(&got[0], &want[0], want.size()) / max_abs(&want[0], want.size()), 1e-4);
    }
    ours->release();
}
struct PrepNone {
    template <class R, class O>
    void operator()(R *, O *, Matrix *) const {}
};
struct PrepAdj {
    template <class R, class O>
    void operator()(R *r, O *o, Matrix *adj) const {
        r->set_adjacency(adj);
        o->set_adjacency(adj);
    }
};
struct PrepDropout {
    int kept;
    bool train;
    template <class R, class O>
    void operator()(R *r, O *o, Matrix *adj) const {
        r->set_adjacency(adj);
        o->set_adjacency(adj);
        r->setContractions(kept);
        o->setContractions(kept);
        r->setMode(train);
        o->setMode(train);
    }
};
static void scenario_family(int N, int C) {
    srand(909);
    family_case<RisiContraction_4, ccn_b200::RisiContraction_4>("r4", N, C, false, true, PrepNone());
    family_case<RisiContraction_10, ccn_b200::RisiContraction_10>("r10", N, C, true, true, PrepAdj());
    PrepDropout train7 = {7, true}, train18 = {18, true}, test5 = {5, false};
    family_case<RisiContraction_18_dropout, ccn_b200::RisiContraction_18_dropout>("dropout7", N, C, true, true, train7);
    family_case<RisiContraction_18_dropout, ccn_b200::RisiContraction_18_dropout>("dropout18", N, C, true, true, train18);
    family_case<RisiContraction_18_dropout, ccn_b200::RisiContraction_18_dropout>("dropout_test_mode", N, C, true, false, test5);
}

// TensorMul and CustomMatMulTensor against the reference classes of the same name (non-zero initial input gradients).
static void scenario_aux(int N, int C) {
    srand(99);
    Tensor3D *A = new Tensor3D(N, N, C), *B = new Tensor3D(N, N, C);
    std::vector<real> a0(A->size), b0(B->size);
    for (int i = 0; i < A->size; ++i) {
        A->value[i] = uniform();
        B->value[i] = uniform();
        a0[i] = uniform();
        b0[i] = uniform();
    }
    TensorMul *rt = new TensorMul(A, B);
    ccn_b200::TensorMul *mt = new ccn_b200::TensorMul(A, B);
    rt->forward();
    mt->forward();
    check("aux", "tensor_mul_forward", max_diff(mt->value, rt->value, rt->size) / max_abs(rt->value, rt->size), 1e-4);
    for (int i = 0; i < rt->size; ++i) rt->gradient[i] = mt->gradient[i] = uniform();
    std::memcpy(A->gradient, &a0[0], sizeof(real) * A->size);
    std::memcpy(B->gradient, &b0[0], sizeof(real) * B->size);
    rt->backward();
    std::vector<real> wa(A->gradient, A->gradient + A->size), wb(B->gradient, B->gradient + B->size);
    std::memcpy(A->gradient, &a0[0], sizeof(real) * A->size);
    std::memcpy(B->gradient, &b0[0], sizeof(real) * B->size);
    mt->backward();
    check("aux", "tensor_mul_backward_first", max_diff(A->gradient, &wa[0], wa.size()) / max_abs(&wa[0], wa.size()), 1e-4);
    check("aux", "tensor_mul_backward_second", max_diff(B->gradient, &wb[0], wb.size()) / max_abs(&wb[0], wb.size()), 1e-4);
    mt->release();

    Matrix *K = new Matrix(C, 18 * C);
    Tensor3D *X = new Tensor3D(N, N, 18 * C);
    for (int i = 0; i < K->size; ++i) K->value[i] = 0.05 * uniform();
    for (int i = 0; i < X->size; ++i) X->value[i] = uniform();
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(X->gradient, 0, sizeof(real) * X->size);
    CustomMatMulTensor *rc = new CustomMatMulTensor(K, X);
    ccn_b200::CustomMatMulTensor *mc = new ccn_b200::CustomMatMulTensor(K, X);
    rc->forward();
    mc->forward();
    check("aux", "custom_matmul_tensor_forward", max_diff(mc->value, rc->value, rc->size) / max_abs(rc->value, rc->size), 1e-4);
    for (int i = 0; i < rc->size; ++i) rc->gradient[i] = mc->gradient[i] = uniform();
    rc->backward();
    std::vector<real> wk(K->gradient, K->gradient + K->size), wx(X->gradient, X->gradient + X->size);
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(X->gradient, 0, sizeof(real) * X->size);
    mc->backward();
    check("aux", "custom_matmul_tensor_backward_first", max_diff(K->gradient, &wk[0], wk.size()) / max_abs(&wk[0], wk.size()), 1e-4);
    check("aux", "custom_matmul_tensor_backward_second", max_diff(X->gradient, &wx[0], wx.size()) / max_abs(&wx[0], wx.size()), 1e-4);
    mc->release();
}

// ccn_b200::RisiContraction_50 vs the reference's N^6 loops (RisiContraction_50.h), signed real adjacency, += semantics.
static void scenario_r50(int N, int C) {
    srand(555);
    std::vector<Tensor3D *> tensors(N);
    const size_t slab = (size_t)N * N * C;
    std::vector<real> g0(slab * N);
    for (int i = 0; i < N; ++i) {
        tensors[i] = new Tensor3D(N, N, C);
        for (size_t j = 0; j < slab; ++j) {
            tensors[i]->value[j] = uniform();
            g0[i * slab + j] = uniform();
        }
    }
    Matrix *adj = new Matrix(N, N);
    for (int i = 0; i < adj->size; ++i) adj->value[i] = uniform();
    RisiContraction_50 *truth = new RisiContraction_50(N, C);
    ccn_b200::RisiContraction_50 *ours = new ccn_b200::RisiContraction_50(N, C);
    for (int i = 0; i < N; ++i) {
        truth->add_tensor(tensors[i]);
        ours->add_tensor(tensors[i]);
    }
    truth->set_adjacency(adj);
    ours->set_adjacency(adj);
    truth->forward();
    ours->forward();
    check("r50", "forward", slab_err(ours->value, truth->value, (size_t)N * N, 50, C), 1e-4);
    for (int i = 0; i < truth->size; ++i) truth->gradient[i] = ours->gradient[i] = uniform();
    std::vector<real> want(slab * N), got(slab * N);
    for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
    truth->backward();
    for (int a = 0; a < N; ++a) std::memcpy(&want[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    for (int a = 0; a < N; ++a) std::memcpy(tensors[a]->gradient, &g0[a * slab], sizeof(real) * slab);
    ours->backward();
    for (int a = 0; a < N; ++a) std::memcpy(&got[a * slab], tensors[a]->gradient, sizeof(real) * slab);
    check("r50", "backward", max_diff(&got[0], &want[0], want.size()) / max_abs(&want[0], want.size()), 1e-4);
    ours->release();
}

// A whole level at once: six vertices with receptive fields of different sizes through ccn_b200::LevelBatch (one
// contraction launch + one mix launch per direction) against six independent reference chains sharing K and b.
static void scenario_batch(int C, int P) {
    srand(4242);
    const int sizes[6] = {5, 9, 3, 9, 1, 7};
    const int maxN = 9;
    Matrix *K = new Matrix(18 * C, P);
    Vector *b = new Vector(P);
    for (int i = 0; i < K->size; ++i) K->value[i] = 0.05 * uniform();
    for (int i = 0; i < b->size; ++i) b->value[i] = 0.5 * uniform();
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(b->gradient, 0, sizeof(real) * b->size);
    ccn_b200::LevelBatch *lvl = new ccn_b200::LevelBatch(maxN, C, P);
    lvl->set_weights(K, b);
    ccn_b200::Executor ref, mine;
    std::vector<std::vector<Tensor3D *> > tensors(6);
    std::vector<Tensor3D *> outs(6);
    std::vector<LeakyReLU3D *> racts(6);
    for (int v = 0; v < 6; ++v) {
        const int n = sizes[v];
        for (int a = 0; a < n; ++a) {
            Tensor3D *t = new Tensor3D(n, n, C);
            for (int j = 0; j < t->size; ++j) t->value[j] = uniform();
            std::memset(t->gradient, 0, sizeof(real) * t->size);
            tensors[v].push_back(t);
        }
        Matrix *adj = new Matrix(n, n);
        for (int i = 0; i < n; ++i)
            for (int j = i; j < n; ++j) {
                const real w = (i == j) ? 1 : ((rand() % 3 == 0) ? 1 : 0);
                adj->value[adj->index(i, j)] = w;
                adj->value[adj->index(j, i)] = w;
            }
        RisiContraction_18 *r_con = new RisiContraction_18(n, C);
        for (int a = 0; a < n; ++a) r_con->add_tensor(tensors[v][a]);
        r_con->set_adjacency(adj);
        Reshape2D *r_2d = new Reshape2D(r_con, n * n, 18 * C);
        MatMul *r_mm = new MatMul(r_2d, K);
        Reshape3D *r_3d = new Reshape3D(r_mm, n, n, P);
        VectorAddTensor *r_add = new VectorAddTensor(b, r_3d);
        racts[v] = new LeakyReLU3D(r_add);
        ref.add(r_con);
        ref.add(r_2d);
        ref.add(r_mm);
        ref.add(r_3d);
        ref.add(r_add);
        ref.add(racts[v]);
        outs[v] = new Tensor3D(n, n, P);
        lvl->add(outs[v], tensors[v], adj);
    }
    mine.add(lvl, ccn_b200::LEVELBATCH_B200);
    ref.forward();
    mine.forward();
    double worst = 0;
    for (int v = 0; v < 6; ++v)
        worst = std::max(worst, max_diff(outs[v]->value, racts[v]->value, racts[v]->size) / max_abs(racts[v]->value, racts[v]->size));
    check("batch", "forward", worst, 1e-4);

    std::vector<std::vector<real> > gz(6);
    for (int v = 0; v < 6; ++v) {
        gz[v].resize(outs[v]->size);
        for (size_t i = 0; i < gz[v].size(); ++i) gz[v][i] = uniform();
        std::memcpy(racts[v]->gradient, &gz[v][0], sizeof(real) * gz[v].size());
    }
    ref.backward();
    std::vector<real> want_K(K->gradient, K->gradient + K->size), want_b(b->gradient, b->gradient + b->size);
    std::vector<std::vector<real> > want_T;
    for (int v = 0; v < 6; ++v)
        for (size_t a = 0; a < tensors[v].size(); ++a) {
            want_T.push_back(std::vector<real>(tensors[v][a]->gradient, tensors[v][a]->gradient + tensors[v][a]->size));
            std::memset(tensors[v][a]->gradient, 0, sizeof(real) * tensors[v][a]->size);
        }
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(b->gradient, 0, sizeof(real) * b->size);
    for (int v = 0; v < 6; ++v) std::memcpy(outs[v]->gradient, &gz[v][0], sizeof(real) * gz[v].size());
    mine.backward();
    double wt = 0, scale = 0;
    size_t idx = 0;
    for (int v = 0; v < 6; ++v)
        for (size_t a = 0; a < tensors[v].size(); ++a, ++idx) {
            wt = std::max(wt, max_diff(tensors[v][a]->gradient, &want_T[idx][0], want_T[idx].size()));
            scale = std::max(scale, max_abs(&want_T[idx][0], want_T[idx].size()));
        }
    check("batch", "backward_tensors", wt / scale, 1e-4);
    check("batch", "backward_K", max_diff(K->gradient, &want_K[0], want_K.size()) / max_abs(&want_K[0], want_K.size()), 1e-4);
    check("batch", "backward_b", max_diff(b->gradient, &want_b[0], want_b.size()) / max_abs(&want_b[0], want_b.size()), 1e-4);
    lvl->release();
}

// bench N C P: wall time of ccn_b200::LevelBatch forward + backward for a batch of B vertices with full receptive fields, all
// inputs and outputs in the reference's host arrays (plain new[]): what a model built on the drop-in classes pays per level.
// Prints one JSON line (picked up by bench.py's `facade` figure).  Not a parity scenario.
#include <chrono>
static void scenario_bench(int N, int C, int P) {
    srand(7);
    const int B = 64, reps = 3;
    Matrix *K = new Matrix(18 * C, P);
    Vector *b = new Vector(P);
    for (int i = 0; i < K->size; ++i) K->value[i] = 0.05 * uniform();
    for (int i = 0; i < b->size; ++i) b->value[i] = 0.5 * uniform();
    std::memset(K->gradient, 0, sizeof(real) * K->size);
    std::memset(b->gradient, 0, sizeof(real) * b->size);
    ccn_b200::LevelBatch *lvl = new ccn_b200::LevelBatch(N, C, P);
    lvl->set_weights(K, b);
    std::vector<Tensor3D *> outs(B);
    for (int v = 0; v < B; ++v) {
        std::vector<Tensor3D *> tensors;
        for (int a = 0; a < N; ++a) {
            Tensor3D *t = new Tensor3D(N, N, C);
            for (int j = 0; j < t->size; ++j) t->value[j] = uniform();
            std::memset(t->gradient, 0, sizeof(real) * t->size);
            tensors.push_back(t);
        }
        Matrix *adj = new Matrix(N, N);
        for (int i = 0; i < N; ++i)
            for (int j = i; j < N; ++j) {
                const real w = (i == j) ? 1 : ((rand() % 10 == 0) ? 1 : 0);
                adj->value[adj->index(i, j)] = w;
                adj->value[adj->index(j, i)] = w;
            }
        outs[v] = new Tensor3D(N, N, P);
        lvl->add(outs[v], tensors, adj);
    }
    double best = 1e30;
    for (int r = 0; r < reps + 1; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        lvl->forward();
        for (int v = 0; v < B; ++v)
            for (int j = 0; j < outs[v]->size; ++j) outs[v]->gradient[j] = 1;
        lvl->backward();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0 && s < best) best = s;
    }
    std::printf("{\"what\": \"ccn_b200::LevelBatch forward + backward from host arrays (C++ drop-in classes, pageable new[] memory), "
                "%d vertices with full receptive fields N=%d, C=%d -> %d\", \"instances\": %d, \"s_per_step\": %.6f, "
                "\"value\": %.1f, \"unit\": \"contractions/s\"}\n",
                B, N, C, P, B, best, B / best);
    lvl->release();
}

int main(int argc, char **argv) {
    const std::string what = argc > 1 ? argv[1] : "all";
    const int a1 = argc > 2 ? std::atoi(argv[2]) : 0, a2 = argc > 3 ? std::atoi(argv[3]) : 0, a3 = argc > 4 ? std::atoi(argv[4]) : 0;
    std::printf("facade real=%s\n", sizeof(real) == 4 ? "float" : "double");
    if (what == "contract") scenario_contract(a1, a2);
    else if (what == "hostapi") scenario_hostapi(a1, a2);
    else if (what == "matmul") scenario_matmul();
    else if (what == "level") scenario_level(a1, a2, a3);
    else if (what == "batch") scenario_batch(a1, a2);
    else if (what == "aux") scenario_aux(a1, a2);
    else if (what == "r50") scenario_r50(a1, a2);
    else if (what == "threads") scenario_threads(a1, a2);
    else if (what == "family") scenario_family(a1, a2);
    else if (what == "bench") scenario_bench(a1, a2, a3);
    else {
        scenario_contract(8, 4);    // BASELINE.json configs[0]
        scenario_contract(12, 32);  // fused kernels
        scenario_hostapi(10, 6);
        scenario_hostapi(16, 64);
        scenario_matmul();
        scenario_level(8, 4, 4);
        scenario_level(16, 32, 32);
        scenario_r50(6, 4);
        scenario_aux(6, 4);
        scenario_aux(16, 32);    // CustomMatMulTensor on the tensor-core kernels
        scenario_batch(4, 4);    // generic kernels + SIMT mix
        scenario_batch(32, 32);  // fused kernels + tensor-core mix, ragged vertex batch
        scenario_threads(12, 32);  // four replicas, four host threads, four streams
        scenario_family(6, 4);     // RisiContraction_4 / _10 / _18_dropout (N^5 reference loops; keep N small)
    }
    std::printf("facade failures=%d\n", failures);
    return failures == 0 ? 0 : 1;
}
