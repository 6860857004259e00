// ccn_ops_b200.h -- header-compatible C++ operator classes over the C-ABI of include/ccn_b200.h.
//
// This is the reference-facing side of the drop-in boundary.  GraphFlow is header-only C++11 whose operators are
// classes derived from its storage types (Vector / Matrix / Tensor3D / Tensor4D, host arrays value[] / gradient[])
// with `setParameter(...)`, `forward()`, `backward()` and are dispatched by a type-tag ladder
// (GraphFlow/GraphFlow.h:182-184, 190-727).  The classes below present the same constructors, public fields and
// methods as the reference's GPU operators so that a model header can use them in place of
//
//     GraphFlow_gpu/RisiContraction_18_gpu.h:846-1804   -> ccn_b200::RisiContraction_18_gpu
//     GraphFlow/RisiContraction_18.h:23-883             -> ccn_b200::RisiContraction_18   (add_tensor API, stack fused)
//     GraphFlow/StackTensor3D.h:25-97                   -> ccn_b200::StackTensor3D        (device-resident stack)
//     GraphFlow/RisiContraction_50.h:22-822             -> ccn_b200::RisiContraction_50   (all 50 contractions)
//     GraphFlow_gpu/MatMul_gpu.h:113-505                -> ccn_b200::MatMul_gpu
//     the per-vertex chain of GraphFlow_gpu/SMP_beta_gpu.h:584-616
//         (stack -> contract -> Reshape2D -> MatMul -> Reshape3D -> VectorAddTensor -> LeakyReLU3D)
//                                                       -> ccn_b200::CCNLevel             (one fused, device-resident op)
//     the `for v` loop around that chain (SMP_beta.h:576-618)       -> ccn_b200::LevelBatch           (one op per LEVEL: the vertex batch)
//     GraphFlow/TensorMul.h:25-94, GraphFlow/CustomMatMulTensor.h:25-93 -> ccn_b200::TensorMul, ccn_b200::CustomMatMulTensor
//     GraphFlow/GraphFlow.h:176-1337 (add / clear / forward / backward)  -> ccn_b200::Executor
//
// It derives from the reference's OWN storage headers: include this file with one of the reference trees on the
// include path (-I<GraphFlow>/GraphFlow for double, -I<GraphFlow>/GraphFlow_32bit for float).  Nothing of the
// reference is copied here.  The kernels compute in fp32; with the double tree the host arrays are converted on
// the way in and out.
//
// There is NO CPU path: the reference's complexity threshold and forward_CPU()/backward_CPU() fallbacks
// (RisiContraction_18_gpu.h:961-984, MatMul_gpu.h:203-223) do not exist here -- every forward()/backward() runs the
// sm_100a kernels, and any failure prints the C-ABI error and aborts, like the reference's assert(err == cudaSuccess).
//
// Threading: one op instance is used by one host thread at a time (the reference's rule, SMP_beta.h:722-729); every
// host thread gets its own ccn_ctx (thread_local), so replicas on different threads / streams run concurrently.
// GraphFlow::~GraphFlow deletes ops through Entity* with a non-virtual destructor (GraphFlow.h:1332-1336), so device
// buffers are also freed by an explicit release().
#ifndef GRAPHFLOW_B200_CCN_OPS_B200_H_INCLUDED
#define GRAPHFLOW_B200_CCN_OPS_B200_H_INCLUDED

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "Matrix.h"    // reference storage types, from the include path
#include "Tensor3D.h"
#include "Tensor4D.h"

#include "../ccn_b200.h"

#ifndef __DRIVER_TYPES_H__  // same typedef as cuda_runtime.h, so that set_gpu_stream(cudaStream_t) keeps its signature
struct CUstream_st;
typedef struct CUstream_st *cudaStream_t;
#endif

namespace ccn_b200 {

// element type of the reference tree on the include path: double (GraphFlow/) or float (GraphFlow_32bit/)
typedef std::remove_pointer<decltype(Vector::value)>::type real;

inline void die(ccn_ctx *ctx, int rc, const char *what) {
    std::fprintf(stderr, "ccn_b200: %s failed: %s (%s)\n", what, ccn_status_string(rc), ctx ? ccn_last_error(ctx) : "");
    std::abort();
}
#define CCN_B200_CHECK(ctx, call)                  \
    do {                                           \
        int rc__ = (call);                         \
        if (rc__ != CCN_OK) ::ccn_b200::die((ctx), rc__, #call); \
    } while (0)

// One context per host thread, created on first use on device $CCN_B200_DEVICE (default 0).
struct ThreadContext {
    ccn_ctx *ctx;
    ThreadContext() : ctx(NULL) {}
    ~ThreadContext() {
        if (ctx) ccn_ctx_destroy(ctx);
    }
};
inline ccn_ctx *context() {
    static thread_local ThreadContext tc;
    if (!tc.ctx) {
        const char *env = std::getenv("CCN_B200_DEVICE");
        int rc = ccn_ctx_create(&tc.ctx, env ? std::atoi(env) : 0);
        if (rc != CCN_OK) die(NULL, rc, "ccn_ctx_create");
    }
    return tc.ctx;
}

// A float device array with (for the double tree) a host conversion buffer.
class DeviceArray {
public:
    DeviceArray() : dev(NULL), cap(0) {}
    float *dev;
    size_t cap;  // floats

    void reserve(size_t n) {
        if (n <= cap) return;
        ccn_ctx *ctx = context();
        if (dev) {
            CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
            CCN_B200_CHECK(ctx, ccn_device_free(ctx, dev));
        }
        void *p = NULL;
        CCN_B200_CHECK(ctx, ccn_device_alloc(ctx, &p, n * sizeof(float)));
        dev = static_cast<float *>(p);
        cap = n;
    }
    void zero(size_t n, void *st) {
        reserve(n);
        ccn_ctx *ctx = context();
        CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, dev, n * sizeof(float), st));
    }
    // host[0..n) -> dev[off..off+n)
    void upload(const real *host, size_t n, size_t off, void *st) {
        ccn_ctx *ctx = context();
        const float *src = as_float(host, n, off);
        CCN_B200_CHECK(ctx, ccn_h2d(ctx, dev + off, src, n * sizeof(float), st));
    }
    // dev[off..off+n) -> host (overwrite).  Synchronises the stream.
    void download(real *host, size_t n, size_t off, void *st) {
        ccn_ctx *ctx = context();
        float *dst = float_target(host, n, off);
        CCN_B200_CHECK(ctx, ccn_d2h(ctx, dst, dev + off, n * sizeof(float), st));
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, st));
        if (!std::is_same<real, float>::value)
            for (size_t i = 0; i < n; ++i) host[i] = (real)dst[i];
    }
    // host[0..n) += dev[off..off+n)   (the reference's `+=` into an input's gradient).  Synchronises the stream.
    void download_add(real *host, size_t n, size_t off, void *st) {
        ccn_ctx *ctx = context();
        if (stage.size() < off + n) stage.resize(off + n);
        CCN_B200_CHECK(ctx, ccn_d2h(ctx, &stage[off], dev + off, n * sizeof(float), st));
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, st));
        for (size_t i = 0; i < n; ++i) host[i] += (real)stage[off + i];
    }
    void release() {
        if (dev) {
            ccn_ctx *ctx = context();
            ccn_stream_synchronize(ctx, NULL);
            ccn_device_free(ctx, dev);
        }
        dev = NULL;
        cap = 0;
        std::vector<float>().swap(stage);
    }

private:
    std::vector<float> stage;  // double tree: converted copy; both trees: landing area of download_add
    const float *as_float(const real *host, size_t n, size_t off) {
        if (std::is_same<real, float>::value) return reinterpret_cast<const float *>(host);
        if (stage.size() < off + n) stage.resize(off + n);  // distinct offsets keep in-flight uploads apart
        for (size_t i = 0; i < n; ++i) stage[off + i] = (float)host[i];
        return &stage[off];
    }
    float *float_target(real *host, size_t n, size_t off) {
        if (std::is_same<real, float>::value) return reinterpret_cast<float *>(host);
        if (stage.size() < off + n) stage.resize(off + n);
        return &stage[off];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// StackTensor3D (GraphFlow/StackTensor3D.h:25-97): same API; the stack is assembled directly in device memory (one
// H2D per neighbour tensor), so a ccn_b200::RisiContraction_18_gpu wired to it never touches a host copy of T.
// ---------------------------------------------------------------------------------------------------------------
class StackTensor3D : public Tensor4D {
public:
    StackTensor3D(int nRows, int nColumns, int nChanels1, int nChanels2)
        : Tensor4D(nRows, nColumns, nChanels1, nChanels2), mirror_host_value(true), device_fresh(false),
          grad_on_device(false), stream(NULL) {
        tensors.clear();
    }
    void setParameter(int nRows, int nColumns, int nChanels1, int nChanels2) {
        this->nRows = nRows;
        this->nColumns = nColumns;
        this->nChanels1 = nChanels1;
        this->nChanels2 = nChanels2;
        size = nRows * nColumns * nChanels1 * nChanels2;
        tensors.clear();
        device_fresh = false;
    }
    void add_tensor(Tensor3D *tensor) {
        assert(tensor->nRows == nColumns);
        assert(tensor->nColumns == nChanels1);
        assert(tensor->nDepth == nChanels2);
        tensors.push_back(tensor);
    }
    void clear() { tensors.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }
    void turn_off_gpu_stream() { stream = NULL; }

    void forward() {  // StackTensor3D.h:54-72
        assert((int)tensors.size() == nRows);
        const size_t slab = (size_t)nColumns * nChanels1 * nChanels2;
        d_value.reserve((size_t)size);
        for (int row = 0; row < nRows; ++row) {
            d_value.upload(tensors[row]->value, slab, row * slab, stream);
            if (mirror_host_value) std::memcpy(value + row * slab, tensors[row]->value, slab * sizeof(real));
        }
        ccn_ctx *ctx = context();
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, stream));
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
        device_fresh = true;
        grad_on_device = false;
    }
    void backward() {  // StackTensor3D.h:74-90: tensors[row]->gradient += gradient (host part + device part)
        assert((int)tensors.size() == nRows);
        const size_t slab = (size_t)nColumns * nChanels1 * nChanels2;
        for (int row = 0; row < nRows; ++row) {
            real *g = tensors[row]->gradient;
            const real *mine = gradient + row * slab;
            for (size_t j = 0; j < slab; ++j) g[j] += mine[j];
            if (grad_on_device) d_grad.download_add(g, slab, row * slab, stream);
        }
    }
    void release() {
        d_value.release();
        d_grad.release();
    }

    std::vector<Tensor3D *> tensors;
    bool mirror_host_value;  // keep value[] valid for consumers that are not ccn_b200 ops (default: yes)
    // device side, used by ccn_b200::RisiContraction_18_gpu
    DeviceArray d_value, d_grad;
    bool device_fresh, grad_on_device;
    cudaStream_t stream;
};

// ---------------------------------------------------------------------------------------------------------------
// RisiContraction_18_gpu (GraphFlow_gpu/RisiContraction_18_gpu.h:846-1804): same constructors, setParameter,
// forward / backward / forward_GPU / backward_GPU, set_gpu_stream / turn_off_gpu_stream and public fields.
// ---------------------------------------------------------------------------------------------------------------
class RisiContraction_18_gpu : public Tensor3D {
public:
    RisiContraction_18_gpu(int max_N, int max_nChanels) : Tensor3D(max_N, max_N, nContractions * max_nChanels) { init(); }
    RisiContraction_18_gpu(Tensor4D *tensor, Matrix *adj)
        : Tensor3D(tensor->nRows, tensor->nRows, nContractions * tensor->nChanels2) {
        init();
        setParameter(tensor, adj);
    }
    RisiContraction_18_gpu(StackTensor3D *stack, Matrix *adj)
        : Tensor3D(stack->nRows, stack->nRows, nContractions * stack->nChanels2) {
        init();
        setParameter(stack, adj);
    }

    void setParameter(Tensor4D *tensor, Matrix *adj) {  // RisiContraction_18_gpu.h:920-936
        this->tensor = tensor;
        this->adj = adj;
        this->stack = NULL;
        N = tensor->nRows;
        nChanels = tensor->nChanels2;
        assert(N == tensor->nColumns);
        assert(N == tensor->nChanels1);
        assert(N == adj->nRows);
        assert(N == adj->nColumns);
        nRows = N;
        nColumns = N;
        nDepth = nChanels * nContractions;
        size = nRows * nColumns * nDepth;
    }
    // device-resident variant: T is read from (and its gradient accumulated into) the stack's device buffers
    void setParameter(StackTensor3D *stack, Matrix *adj) {
        setParameter(static_cast<Tensor4D *>(stack), adj);
        this->stack = stack;
    }

    void set_gpu_stream(cudaStream_t stream) {  // :947-950
        use_gpu_stream = true;
        this->stream = stream;
    }
    void turn_off_gpu_stream() { use_gpu_stream = false; }  // :953-955

    void forward() { forward_GPU(); }    // no complexity threshold, no CPU path
    void backward() { backward_GPU(); }

    void forward_GPU() {  // replaces :1509-1568
        ccn_ctx *ctx = context();
        void *st = use_gpu_stream ? (void *)stream : NULL;
        const size_t szT = (size_t)N * N * N * nChanels, szA = (size_t)N * N, szO = (size_t)size;
        const float *Tdev;
        if (stack && stack->device_fresh) {
            Tdev = stack->d_value.dev;
        } else {
            d_T.reserve(szT);
            d_T.upload(tensor->value, szT, 0, st);
            Tdev = d_T.dev;
        }
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, st);
        d_out.reserve(szO);
        CCN_B200_CHECK(ctx, ccn_contract18_forward(ctx, Tdev, NULL, d_adj.dev, d_out.dev, NULL, N, nChanels, 1, (int64_t)szT,
                                                   (int64_t)szA, (int64_t)szO, CCN_ADJ_POSITIVE_PART, st));
        d_out.download(value, szO, 0, st);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;  // :1565-1567
    }

    void backward_GPU() {  // replaces :1639-1689; `+=` into tensor->gradient (:683)
        ccn_ctx *ctx = context();
        void *st = use_gpu_stream ? (void *)stream : NULL;
        const size_t szT = (size_t)N * N * N * nChanels, szA = (size_t)N * N, szO = (size_t)size;
        d_gout.reserve(szO);
        d_gout.upload(gradient, szO, 0, st);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, st);
        if (stack) {  // accumulate on the device; StackTensor3D::backward brings it home
            stack->d_grad.reserve(szT);
            const float beta = stack->grad_on_device ? 1.0f : 0.0f;
            CCN_B200_CHECK(ctx, ccn_contract18_backward(ctx, d_gout.dev, d_adj.dev, stack->d_grad.dev, NULL, NULL, N, nChanels,
                                                        1, (int64_t)szO, (int64_t)szA, (int64_t)szT, CCN_ADJ_POSITIVE_PART, beta, st));
            CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, st));
            stack->grad_on_device = true;
        } else {
            d_gT.reserve(szT);
            CCN_B200_CHECK(ctx, ccn_contract18_backward(ctx, d_gout.dev, d_adj.dev, d_gT.dev, NULL, NULL, N, nChanels, 1,
                                                        (int64_t)szO, (int64_t)szA, (int64_t)szT, CCN_ADJ_POSITIVE_PART, 0.0f, st));
            d_gT.download_add(tensor->gradient, szT, 0, st);
        }
    }

    void release() {
        d_T.release();
        d_gT.release();
        d_adj.release();
        d_out.release();
        d_gout.release();
    }

    int N;
    int nChanels;
    Tensor4D *tensor;
    Matrix *adj;
    StackTensor3D *stack;  // non-NULL when wired to a device-resident stack
    bool use_gpu_stream;
    cudaStream_t stream;
    static const int nContractions = 18;  // :1749

private:
    void init() {
        tensor = NULL;
        adj = NULL;
        stack = NULL;
        N = nChanels = 0;
        use_gpu_stream = false;
        stream = NULL;
    }
    DeviceArray d_T, d_gT, d_adj, d_out, d_gout;
};

// ---------------------------------------------------------------------------------------------------------------
// RisiContraction_18 (GraphFlow/RisiContraction_18.h:23-883): the add_tensor / set_adjacency API used by the models
// that never had a GPU variant (SMP_2D_ver8.h, SMP_omega_physics.h).  The stack is fused into the upload.
// ---------------------------------------------------------------------------------------------------------------
class RisiContraction_18 : public Tensor3D {
public:
    RisiContraction_18(int max_nRows, int max_nColumns, int max_nDepth) : Tensor3D(max_nRows, max_nColumns, max_nDepth) {
        N = nChanels = 0;
        adj = NULL;
        stream = NULL;
    }
    RisiContraction_18(int N, int nChanels) : Tensor3D(N, N, nContractions * nChanels) {
        this->N = N;
        this->nChanels = nChanels;
        adj = NULL;
        stream = NULL;
    }
    void setParameter(int N, int nChanels) {  // RisiContraction_18.h:36-47
        this->N = N;
        this->nChanels = nChanels;
        nRows = N;
        nColumns = N;
        nDepth = nChanels * nContractions;
        size = nRows * nColumns * nDepth;
        tensors.clear();
    }
    void add_tensor(Tensor3D *tensor) {  // :49-55
        assert(tensor->nRows == N);
        assert(tensor->nColumns == N);
        assert(tensor->nDepth == nChanels);
        tensors.push_back(tensor);
    }
    void set_adjacency(Matrix *adj) {  // :57-61
        assert(adj->nRows == N);
        assert(adj->nColumns == N);
        this->adj = adj;
    }
    void clear() { tensors.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }

    void forward() {  // replaces :73-331
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_T.reserve(szT);
        for (int a = 0; a < N; ++a) d_T.upload(tensors[a]->value, slab, a * slab, stream);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, stream);
        d_out.reserve(szO);
        CCN_B200_CHECK(ctx, ccn_contract18_forward(ctx, d_T.dev, NULL, d_adj.dev, d_out.dev, NULL, N, nChanels, 1, (int64_t)szT,
                                                   (int64_t)szA, (int64_t)szO, CCN_ADJ_POSITIVE_PART, stream));
        d_out.download(value, szO, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;  // :327-329
    }
    void backward() {  // replaces :333-560
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_gout.reserve(szO);
        d_gout.upload(gradient, szO, 0, stream);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, stream);
        d_gT.reserve(szT);
        CCN_B200_CHECK(ctx, ccn_contract18_backward(ctx, d_gout.dev, d_adj.dev, d_gT.dev, NULL, NULL, N, nChanels, 1, (int64_t)szO,
                                                    (int64_t)szA, (int64_t)szT, CCN_ADJ_POSITIVE_PART, 0.0f, stream));
        for (int a = 0; a < N; ++a) d_gT.download_add(tensors[a]->gradient, slab, a * slab, stream);
    }
    void release() {
        d_T.release();
        d_gT.release();
        d_adj.release();
        d_out.release();
        d_gout.release();
    }

    int N;
    int nChanels;
    std::vector<Tensor3D *> tensors;
    Matrix *adj;
    cudaStream_t stream;
    static const int nContractions = 18;

private:
    DeviceArray d_T, d_gT, d_adj, d_out, d_gout;
};

// ---------------------------------------------------------------------------------------------------------------
// RisiContraction_50 (GraphFlow/RisiContraction_50.h:22-822): all 50 contractions, same add_tensor / set_adjacency API;
// the reference multiplies by the raw adjacency entry here (value_at, :63-65).
// ---------------------------------------------------------------------------------------------------------------
class RisiContraction_50 : public Tensor3D {
public:
    RisiContraction_50(int max_nRows, int max_nColumns, int max_nDepth) : Tensor3D(max_nRows, max_nColumns, max_nDepth) {
        N = nChanels = 0;
        adj = NULL;
        stream = NULL;
    }
    RisiContraction_50(int N, int nChanels) : Tensor3D(N, N, nContractions * nChanels) {
        this->N = N;
        this->nChanels = nChanels;
        adj = NULL;
        stream = NULL;
    }
    void setParameter(int N, int nChanels) {  // RisiContraction_50.h:36-47
        this->N = N;
        this->nChanels = nChanels;
        nRows = N;
        nColumns = N;
        nDepth = nChanels * nContractions;
        size = nRows * nColumns * nDepth;
        tensors.clear();
    }
    void add_tensor(Tensor3D *tensor) {
        assert(tensor->nRows == N);
        assert(tensor->nColumns == N);
        assert(tensor->nDepth == nChanels);
        tensors.push_back(tensor);
    }
    void set_adjacency(Matrix *adj) {
        assert(adj->nRows == N);
        assert(adj->nColumns == N);
        this->adj = adj;
    }
    void clear() { tensors.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }

    void forward() {  // replaces RisiContraction_50.h:73-441
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_T.reserve(szT);
        for (int a = 0; a < N; ++a) d_T.upload(tensors[a]->value, slab, a * slab, stream);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, stream);
        d_out.reserve(szO);
        CCN_B200_CHECK(ctx, ccn_contract50_forward(ctx, d_T.dev, NULL, d_adj.dev, d_out.dev, NULL, N, nChanels, 1, (int64_t)szT,
                                                   (int64_t)szA, (int64_t)szO, CCN_ADJ_RAW, stream));
        d_out.download(value, szO, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
    }
    void backward() {  // replaces RisiContraction_50.h:443-802
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_gout.reserve(szO);
        d_gout.upload(gradient, szO, 0, stream);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, stream);
        d_gT.reserve(szT);
        CCN_B200_CHECK(ctx, ccn_contract50_backward(ctx, d_gout.dev, d_adj.dev, d_gT.dev, NULL, NULL, N, nChanels, 1, (int64_t)szO,
                                                    (int64_t)szA, (int64_t)szT, CCN_ADJ_RAW, 0.0f, stream));
        for (int a = 0; a < N; ++a) d_gT.download_add(tensors[a]->gradient, slab, a * slab, stream);
    }
    void release() {
        d_T.release();
        d_gT.release();
        d_adj.release();
        d_out.release();
        d_gout.release();
    }

    int N;
    int nChanels;
    std::vector<Tensor3D *> tensors;
    Matrix *adj;
    cudaStream_t stream;
    static const int nContractions = 50;

private:
    DeviceArray d_T, d_gT, d_adj, d_out, d_gout;
};

// ---------------------------------------------------------------------------------------------------------------
// The other members of the contraction family, same add_tensor / set_adjacency / forward / backward API:
//   RisiContraction_4          (GraphFlow/RisiContraction_4.h:24-190; no adjacency)           SMP_gamma*
//   RisiContraction_10         (GraphFlow/RisiContraction_10.h:23-244; raw adjacency)
//   RisiContraction_18_dropout (GraphFlow/RisiContraction_18_dropout.h:22-813)                SMP_sigma*
// all through ccn_contract_family_forward / _backward.
// ---------------------------------------------------------------------------------------------------------------
template <int VARIANT>
class ContractionFamilyOp : public Tensor3D {
public:
    ContractionFamilyOp(int max_nRows, int max_nColumns, int max_nDepth) : Tensor3D(max_nRows, max_nColumns, max_nDepth) {
        N = nChanels = 0;
        adj = NULL;
        stream = NULL;
    }
    ContractionFamilyOp(int N, int nChanels) : Tensor3D(N, N, VARIANT * nChanels) {
        this->N = N;
        this->nChanels = nChanels;
        adj = NULL;
        stream = NULL;
    }
    void setParameter(int N, int nChanels) {
        this->N = N;
        this->nChanels = nChanels;
        nRows = N;
        nColumns = N;
        nDepth = nChanels * nContractions;
        size = nRows * nColumns * nDepth;
        tensors.clear();
    }
    void add_tensor(Tensor3D *tensor) {
        assert(tensor->nRows == N);
        assert(tensor->nColumns == N);
        assert(tensor->nDepth == nChanels);
        tensors.push_back(tensor);
    }
    void clear() { tensors.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }
    void release() {
        d_T.release();
        d_gT.release();
        d_adj.release();
        d_out.release();
        d_gout.release();
    }

    int N;
    int nChanels;
    std::vector<Tensor3D *> tensors;
    Matrix *adj;
    cudaStream_t stream;
    static const int nContractions = VARIANT;

protected:
    void run_forward(uint64_t keep_mask, int adj_mode, float out_scale) {
        assert((int)tensors.size() == N);
        assert(VARIANT == 4 || adj != NULL);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_T.reserve(szT);
        for (int a = 0; a < N; ++a) d_T.upload(tensors[a]->value, slab, a * slab, stream);
        if (VARIANT != 4) {
            d_adj.reserve(szA);
            d_adj.upload(adj->value, szA, 0, stream);
        }
        d_out.reserve(szO);
        CCN_B200_CHECK(ctx, ccn_contract_family_forward(ctx, VARIANT, keep_mask, d_T.dev, NULL, VARIANT != 4 ? d_adj.dev : NULL,
                                                        d_out.dev, NULL, N, nChanels, 1, (int64_t)szT, (int64_t)szA, (int64_t)szO,
                                                        adj_mode, out_scale, stream));
        d_out.download(value, szO, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
    }
    void run_backward(uint64_t keep_mask, int adj_mode) {
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * nChanels, szT = slab * N, szA = (size_t)N * N, szO = (size_t)size;
        d_gout.reserve(szO);
        d_gout.upload(gradient, szO, 0, stream);
        if (VARIANT != 4) {
            d_adj.reserve(szA);
            d_adj.upload(adj->value, szA, 0, stream);
        }
        d_gT.reserve(szT);
        CCN_B200_CHECK(ctx, ccn_contract_family_backward(ctx, VARIANT, keep_mask, d_gout.dev, VARIANT != 4 ? d_adj.dev : NULL, d_gT.dev,
                                                         NULL, NULL, N, nChanels, 1, (int64_t)szO, (int64_t)szA, (int64_t)szT, adj_mode,
                                                         0.0f, stream));
        for (int a = 0; a < N; ++a) d_gT.download_add(tensors[a]->gradient, slab, a * slab, stream);
    }

private:
    DeviceArray d_T, d_gT, d_adj, d_out, d_gout;
};

class RisiContraction_4 : public ContractionFamilyOp<4> {
public:
    RisiContraction_4(int max_nRows, int max_nColumns, int max_nDepth) : ContractionFamilyOp<4>(max_nRows, max_nColumns, max_nDepth) {}
    RisiContraction_4(int N, int nChanels) : ContractionFamilyOp<4>(N, nChanels) {}
    void forward() { run_forward(~0ull, CCN_ADJ_RAW, 1.0f); }   // replaces RisiContraction_4.h:68-123
    void backward() { run_backward(~0ull, CCN_ADJ_RAW); }       // replaces :125-180
};

class RisiContraction_10 : public ContractionFamilyOp<10> {
public:
    RisiContraction_10(int max_nRows, int max_nColumns, int max_nDepth) : ContractionFamilyOp<10>(max_nRows, max_nColumns, max_nDepth) {}
    RisiContraction_10(int N, int nChanels) : ContractionFamilyOp<10>(N, nChanels) {}
    void set_adjacency(Matrix *adj) {
        assert(adj->nRows == N);
        assert(adj->nColumns == N);
        this->adj = adj;
    }
    void forward() { run_forward(~0ull, CCN_ADJ_RAW, 1.0f); }   // replaces RisiContraction_10.h:72-150 (value_at: raw adj)
    void backward() { run_backward(~0ull, CCN_ADJ_RAW); }       // replaces :152-230
};

class RisiContraction_18_dropout : public ContractionFamilyOp<18> {
public:
    RisiContraction_18_dropout(int max_nRows, int max_nColumns, int max_nDepth)
        : ContractionFamilyOp<18>(max_nRows, max_nColumns, max_nDepth) {
        init();
    }
    RisiContraction_18_dropout(int N, int nChanels) : ContractionFamilyOp<18>(N, nChanels) { init(); }
    void setContractions(int nKept) {  // :50-55
        assert(nKept > 0);
        assert(nKept <= nContractions);
        this->nKept = nKept;
    }
    void setTrainMode() { mode = true; }
    void setTestMode() { mode = false; }
    void setMode(bool mode) { this->mode = mode; }
    void set_adjacency(Matrix *adj) {
        assert(adj->nRows == N);
        assert(adj->nColumns == N);
        this->adj = adj;
    }
    void forward() {  // replaces :104-478
        assert(nKept > 0);
        if (mode) {  // the same rand() draws as the reference (:113-126), so a seeded run keeps the same slabs
            for (int i = 0; i < nContractions; ++i) use[i] = false;
            for (int i = 0; i < nKept; ++i) {
                while (true) {
                    const int j = rand() % nContractions;
                    if (!use[j]) {
                        use[j] = true;
                        break;
                    }
                }
            }
        } else {
            for (int i = 0; i < nContractions; ++i) use[i] = true;
        }
        run_forward(mask(), CCN_ADJ_POSITIVE_PART, mode ? 1.0f : (float)((double)nKept / (double)nContractions));  // :467-472
    }
    void backward() {  // replaces :480-797
        assert(nKept > 0);
        assert(mode == true);  // :485
        run_backward(mask(), CCN_ADJ_POSITIVE_PART);
    }

    bool mode;  // true = train
    bool *use;
    int nKept;

private:
    void init() {
        nKept = 0;
        use = new bool[nContractions];
        for (int i = 0; i < nContractions; ++i) use[i] = false;
        mode = true;
    }
    uint64_t mask() const {
        uint64_t m = 0;
        for (int i = 0; i < nContractions; ++i)
            if (use[i]) m |= (uint64_t)1 << i;
        return m;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// MatMul_gpu (GraphFlow_gpu/MatMul_gpu.h:113-505): value = first * second; backward `+=` into both inputs.
// ---------------------------------------------------------------------------------------------------------------
class MatMul_gpu : public Matrix {
public:
    MatMul_gpu(int max_first_nRows, int max_first_nColumns, int max_second_nRows, int max_second_nColumns)
        : Matrix(max_first_nRows, max_second_nColumns) {
        assert(max_first_nColumns == max_second_nRows);
        first = second = NULL;
        use_gpu_stream = false;
        stream = NULL;
    }
    MatMul_gpu(Matrix *first, Matrix *second) : Matrix(first->nRows, second->nColumns) {
        use_gpu_stream = false;
        stream = NULL;
        setParameter(first, second);
    }
    void setParameter(Matrix *first, Matrix *second) {  // MatMul_gpu.h:167-181
        assert(first->nColumns == second->nRows);
        this->first = first;
        this->second = second;
        nRows = first->nRows;
        nColumns = second->nColumns;
        size = nRows * nColumns;
    }
    void set_gpu_stream(cudaStream_t stream) {
        use_gpu_stream = true;
        this->stream = stream;
    }
    void turn_off_gpu_stream() { use_gpu_stream = false; }

    void forward() { forward_GPU(); }
    void backward() { backward_GPU(); }

    void forward_GPU() {  // replaces :273-313
        ccn_ctx *ctx = context();
        void *st = use_gpu_stream ? (void *)stream : NULL;
        const size_t M = first->nRows, K = first->nColumns, P = second->nColumns;
        d_X.reserve(M * K);
        d_X.upload(first->value, M * K, 0, st);
        d_W.reserve(K * P);
        d_W.upload(second->value, K * P, 0, st);
        d_Y.reserve(M * P);
        CCN_B200_CHECK(ctx, ccn_mix_forward(ctx, d_X.dev, d_W.dev, NULL, d_Y.dev, NULL, (int64_t)M, (int)K, (int)P, 0.0f, st));
        d_Y.download(value, M * P, 0, st);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;  // :310-312
    }
    void backward_GPU() {  // replaces :412-466
        ccn_ctx *ctx = context();
        void *st = use_gpu_stream ? (void *)stream : NULL;
        const size_t M = first->nRows, K = first->nColumns, P = second->nColumns;
        d_X.reserve(M * K);
        d_X.upload(first->value, M * K, 0, st);
        d_W.reserve(K * P);
        d_W.upload(second->value, K * P, 0, st);
        d_gY.reserve(M * P);
        d_gY.upload(gradient, M * P, 0, st);
        d_gX.reserve(M * K);
        d_gW.zero(K * P, st);
        CCN_B200_CHECK(ctx, ccn_mix_backward(ctx, d_X.dev, d_W.dev, NULL, NULL, d_gY.dev, d_gX.dev, d_gW.dev, NULL, (int64_t)M,
                                             (int)K, (int)P, 0.0f, 0.0f, st));
        d_gX.download_add(first->gradient, M * K, 0, st);
        d_gW.download_add(second->gradient, K * P, 0, st);
    }
    void release() {
        d_X.release();
        d_W.release();
        d_Y.release();
        d_gY.release();
        d_gX.release();
        d_gW.release();
    }

    Matrix *first;
    Matrix *second;
    bool use_gpu_stream;
    cudaStream_t stream;

private:
    DeviceArray d_X, d_W, d_Y, d_gY, d_gX, d_gW;
};

// ---------------------------------------------------------------------------------------------------------------
// TensorMul (GraphFlow/TensorMul.h:25-94): per-channel [R x K] . [K x C] product, the feature mix of SMP_2D v1-5
// (SMP_2D.h:573).  Same constructors / setParameter / forward / backward (+= into both inputs' gradients).
// ---------------------------------------------------------------------------------------------------------------
class TensorMul : public Tensor3D {
public:
    TensorMul(int max_nRows, int max_nColumns, int max_nDepth) : Tensor3D(max_nRows, max_nColumns, max_nDepth) {
        first = second = NULL;
        stream = NULL;
    }
    TensorMul(Tensor3D *first, Tensor3D *second) : Tensor3D(first->nRows, second->nColumns, first->nDepth) {
        stream = NULL;
        setParameter(first, second);
    }
    void setParameter(Tensor3D *first, Tensor3D *second) {  // TensorMul.h:34-46
        assert(first->nColumns == second->nRows);
        assert(first->nDepth == second->nDepth);
        this->first = first;
        this->second = second;
        nRows = first->nRows;
        nColumns = second->nColumns;
        nDepth = first->nDepth;
        size = nRows * nColumns * nDepth;
    }
    void set_gpu_stream(cudaStream_t s) { stream = s; }
    void forward() {  // replaces TensorMul.h:48-70
        ccn_ctx *ctx = context();
        const size_t sa = (size_t)first->size, sb = (size_t)second->size, so = (size_t)size;
        d_A.reserve(sa);
        d_A.upload(first->value, sa, 0, stream);
        d_B.reserve(sb);
        d_B.upload(second->value, sb, 0, stream);
        d_out.reserve(so);
        CCN_B200_CHECK(ctx, ccn_tensor_mul_forward(ctx, d_A.dev, d_B.dev, d_out.dev, first->nRows, first->nColumns, second->nColumns, nDepth,
                                                   1, stream));
        d_out.download(value, so, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
    }
    void backward() {  // replaces TensorMul.h:72-86
        ccn_ctx *ctx = context();
        const size_t sa = (size_t)first->size, sb = (size_t)second->size, so = (size_t)size;
        d_A.reserve(sa);
        d_A.upload(first->value, sa, 0, stream);
        d_B.reserve(sb);
        d_B.upload(second->value, sb, 0, stream);
        d_g.reserve(so);
        d_g.upload(gradient, so, 0, stream);
        d_gA.reserve(sa);
        d_gB.reserve(sb);
        CCN_B200_CHECK(ctx, ccn_tensor_mul_backward(ctx, d_A.dev, d_B.dev, d_g.dev, d_gA.dev, d_gB.dev, first->nRows, first->nColumns,
                                                    second->nColumns, nDepth, 1, 0.0f, stream));
        d_gA.download_add(first->gradient, sa, 0, stream);
        d_gB.download_add(second->gradient, sb, 0, stream);
    }
    void release() {
        DeviceArray *all[] = {&d_A, &d_B, &d_out, &d_g, &d_gA, &d_gB};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
    }
    Tensor3D *first;
    Tensor3D *second;
    cudaStream_t stream;

private:
    DeviceArray d_A, d_B, d_out, d_g, d_gA, d_gB;
};

// ---------------------------------------------------------------------------------------------------------------
// CustomMatMulTensor (GraphFlow/CustomMatMulTensor.h:25-93; SMP_2D_ver8.h:526-527): value[i,j,k] = sum_v first[k,v]
// second[i,j,v] -- the feature mix with the weights stored [C_out, 18 C].  Runs on the tensor-core mix kernels.
// ---------------------------------------------------------------------------------------------------------------
class CustomMatMulTensor : public Tensor3D {
public:
    CustomMatMulTensor(int max_nRows, int max_nColumns, int max_nDepth) : Tensor3D(max_nRows, max_nColumns, max_nDepth) {
        first = NULL;
        second = NULL;
        stream = NULL;
    }
    CustomMatMulTensor(Matrix *first, Tensor3D *second) : Tensor3D(second->nRows, second->nColumns, first->nRows) {
        stream = NULL;
        setParameter(first, second);
    }
    void setParameter(Matrix *first, Tensor3D *second) {  // CustomMatMulTensor.h:34-45
        assert(first->nColumns == second->nDepth);
        this->first = first;
        this->second = second;
        nRows = second->nRows;
        nColumns = second->nColumns;
        nDepth = first->nRows;
        size = nRows * nColumns * nDepth;
    }
    void set_gpu_stream(cudaStream_t s) { stream = s; }
    void forward() {  // replaces CustomMatMulTensor.h:47-67
        ccn_ctx *ctx = context();
        const size_t sk = (size_t)first->size, sx = (size_t)second->size, sy = (size_t)size;
        d_K.reserve(sk);
        d_K.upload(first->value, sk, 0, stream);
        d_X.reserve(sx);
        d_X.upload(second->value, sx, 0, stream);
        d_Y.reserve(sy);
        CCN_B200_CHECK(ctx, ccn_custom_matmul_tensor_forward(ctx, d_K.dev, d_X.dev, d_Y.dev, (int64_t)nRows * nColumns, second->nDepth, nDepth,
                                                             stream));
        d_Y.download(value, sy, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
    }
    void backward() {  // replaces CustomMatMulTensor.h:69-85
        ccn_ctx *ctx = context();
        const size_t sk = (size_t)first->size, sx = (size_t)second->size, sy = (size_t)size;
        d_K.reserve(sk);
        d_K.upload(first->value, sk, 0, stream);
        d_X.reserve(sx);
        d_X.upload(second->value, sx, 0, stream);
        d_gY.reserve(sy);
        d_gY.upload(gradient, sy, 0, stream);
        d_gK.zero(sk, stream);
        d_gX.reserve(sx);
        CCN_B200_CHECK(ctx, ccn_custom_matmul_tensor_backward(ctx, d_K.dev, d_X.dev, d_gY.dev, d_gK.dev, d_gX.dev, (int64_t)nRows * nColumns,
                                                              second->nDepth, nDepth, 0.0f, stream));
        d_gK.download_add(first->gradient, sk, 0, stream);
        d_gX.download_add(second->gradient, sx, 0, stream);
    }
    void release() {
        DeviceArray *all[] = {&d_K, &d_X, &d_Y, &d_gY, &d_gK, &d_gX};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
    }
    Matrix *first;
    Tensor3D *second;
    cudaStream_t stream;

private:
    DeviceArray d_K, d_X, d_Y, d_gY, d_gK, d_gX;
};

// ---------------------------------------------------------------------------------------------------------------
// CCNLevel: one vertex of one CCN level as a single device-resident operator.  Replaces the seven-op chain
//   StackTensor3D -> RisiContraction_18(_gpu) -> Reshape2D -> MatMul(_gpu)(., K) -> Reshape3D -> VectorAddTensor(b, .)
//   -> LeakyReLU3D                                     (SMP_beta.h:600-616, GraphFlow_gpu/SMP_beta_gpu.h:584-616)
// value [N, N, C_out] = lrelu(contract18(stack(tensors), adj) . K + b).  One upload of the neighbour tensors, one
// download of the activation; the [N,N,18C] contraction output and the pre-activation never leave the device.
// backward(): `+=` into tensors[a]->gradient, K->gradient, b->gradient.
// ---------------------------------------------------------------------------------------------------------------
class CCNLevel : public Tensor3D {
public:
    CCNLevel(int max_N, int max_C_in, int max_C_out) : Tensor3D(max_N, max_N, max_C_out) {
        N = C_in = C_out = 0;
        adj = NULL;
        K = NULL;
        b = NULL;
        alpha = 0.01f;  // LeakyReLU3D.h:31
        stream = NULL;
    }
    void setParameter(int N, int C_in, Matrix *K, Vector *b) {
        assert(K->nRows == nContractions * C_in);
        assert(b->size == K->nColumns);
        this->N = N;
        this->C_in = C_in;
        this->C_out = K->nColumns;
        this->K = K;
        this->b = b;
        nRows = N;
        nColumns = N;
        nDepth = C_out;
        size = nRows * nColumns * nDepth;
        tensors.clear();
    }
    void add_tensor(Tensor3D *tensor) {
        assert(tensor->nRows == N);
        assert(tensor->nColumns == N);
        assert(tensor->nDepth == C_in);
        tensors.push_back(tensor);
    }
    void set_adjacency(Matrix *adj) {
        assert(adj->nRows == N);
        assert(adj->nColumns == N);
        this->adj = adj;
    }
    void clear() { tensors.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }

    void forward() {
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * C_in, szT = slab * N, szA = (size_t)N * N, M = (size_t)N * N;
        const size_t Kd = (size_t)nContractions * C_in, szX = M * Kd, szY = M * C_out;
        d_T.reserve(szT);
        for (int a = 0; a < N; ++a) d_T.upload(tensors[a]->value, slab, a * slab, stream);
        d_adj.reserve(szA);
        d_adj.upload(adj->value, szA, 0, stream);
        d_K.reserve(Kd * C_out);
        d_K.upload(K->value, Kd * C_out, 0, stream);
        d_b.reserve(C_out);
        d_b.upload(b->value, C_out, 0, stream);
        d_X.reserve(szX);
        d_Y.reserve(szY);
        d_Z.reserve(szY);
        CCN_B200_CHECK(ctx, ccn_contract18_forward(ctx, d_T.dev, NULL, d_adj.dev, d_X.dev, NULL, N, C_in, 1, (int64_t)szT,
                                                   (int64_t)szA, (int64_t)szX, CCN_ADJ_POSITIVE_PART, stream));
        CCN_B200_CHECK(ctx, ccn_mix_forward(ctx, d_X.dev, d_K.dev, d_b.dev, d_Y.dev, d_Z.dev, (int64_t)M, (int)Kd, C_out, alpha, stream));
        d_Z.download(value, szY, 0, stream);
        for (int i = 0; i < size; ++i) gradient[i] = 0.0;
    }
    void backward() {
        assert((int)tensors.size() == N);
        ccn_ctx *ctx = context();
        const size_t slab = (size_t)N * N * C_in, szT = slab * N, szA = (size_t)N * N, M = (size_t)N * N;
        const size_t Kd = (size_t)nContractions * C_in, szX = M * Kd, szY = M * C_out;
        d_gZ.reserve(szY);
        d_gZ.upload(gradient, szY, 0, stream);
        d_gX.reserve(szX);
        d_gK.zero(Kd * C_out, stream);
        d_gb.zero(C_out, stream);
        d_gT.reserve(szT);
        CCN_B200_CHECK(ctx, ccn_mix_backward(ctx, d_X.dev, d_K.dev, d_b.dev, d_Y.dev, d_gZ.dev, d_gX.dev, d_gK.dev, d_gb.dev, (int64_t)M,
                                             (int)Kd, C_out, alpha, 0.0f, stream));
        CCN_B200_CHECK(ctx, ccn_contract18_backward(ctx, d_gX.dev, d_adj.dev, d_gT.dev, NULL, NULL, N, C_in, 1, (int64_t)szX,
                                                    (int64_t)szA, (int64_t)szT, CCN_ADJ_POSITIVE_PART, 0.0f, stream));
        for (int a = 0; a < N; ++a) d_gT.download_add(tensors[a]->gradient, slab, a * slab, stream);
        d_gK.download_add(K->gradient, Kd * C_out, 0, stream);
        d_gb.download_add(b->gradient, C_out, 0, stream);
    }
    void release() {
        DeviceArray *all[] = {&d_T, &d_adj, &d_K, &d_b, &d_X, &d_Y, &d_Z, &d_gZ, &d_gX, &d_gK, &d_gb, &d_gT};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
    }

    int N, C_in, C_out;
    std::vector<Tensor3D *> tensors;
    Matrix *adj;
    Matrix *K;
    Vector *b;
    float alpha;
    cudaStream_t stream;
    static const int nContractions = 18;

private:
    DeviceArray d_T, d_adj, d_K, d_b, d_X, d_Y, d_Z, d_gZ, d_gX, d_gK, d_gb, d_gT;
};

// ---------------------------------------------------------------------------------------------------------------
// LevelBatch: ALL vertices of one CCN level (of one graph or of a whole mini-batch of graphs) as ONE operator: one
// contraction launch, one feature-mix launch per direction for the whole vertex batch instead of the reference's
// seven host-driven ops per vertex (SMP_beta.h:576-618 loops `for v` and adds contract[v], reshape2D[v], represent[v],
// reshape3D[v], add[v], f[v] to the graph one vertex at a time).  Vertices may have different receptive-field sizes
// n_v <= max_N (ragged batch).  Outputs are ordinary Tensor3D objects owned by the caller, so the next level's
// promotion ops (MatTensorMul / TensorMatMul) and the read-out head consume them unchanged.
//
//   LevelBatch lvl(max_N, C_in, C_out);
//   lvl.set_weights(K, b);                               // K [18 C_in, C_out], b [C_out]   (shared by the level)
//   for v: lvl.add(f_v /*Tensor3D [n_v,n_v,C_out], caller-owned*/, quadratic_v /*n_v tensors [n_v,n_v,C_in]*/, adj_v);
//   graph.add(&lvl, LEVELBATCH_B200);   ->  forward(): f_v->value filled, f_v->gradient zeroed
//                                           backward(): reads f_v->gradient; += into the inputs', K's and b's gradients
// ---------------------------------------------------------------------------------------------------------------
class LevelBatch {
public:
    LevelBatch(int max_N, int C_in, int C_out) : max_N(max_N), C_in(C_in), C_out(C_out), K(NULL), b(NULL), alpha(0.01f), stream(NULL) {}
    void set_weights(Matrix *K, Vector *b) {
        assert(K->nRows == nContractions * C_in && K->nColumns == C_out && b->size == C_out);
        this->K = K;
        this->b = b;
    }
    void clear() { items.clear(); }
    void set_gpu_stream(cudaStream_t s) { stream = s; }
    int add(Tensor3D *out, const std::vector<Tensor3D *> &tensors, Matrix *adj) {
        const int n = (int)tensors.size();
        assert(n >= 1 && n <= max_N && adj->nRows == n && adj->nColumns == n);
        assert(out->nRows == n && out->nColumns == n && out->nDepth == C_out);
        for (int a = 0; a < n; ++a) assert(tensors[a]->nRows == n && tensors[a]->nColumns == n && tensors[a]->nDepth == C_in);
        Item it;
        it.out = out;
        it.tensors = tensors;
        it.adj = adj;
        items.push_back(it);
        return (int)items.size() - 1;
    }
    size_t size() const { return items.size(); }

    void forward() {
        const size_t B = items.size();
        if (B == 0) return;
        ccn_ctx *ctx = context();
        const size_t N = max_N, sT = N * N * N * C_in, sA = N * N, Kd = (size_t)nContractions * C_in, sX = N * N * Kd, sY = N * N * C_out;
        d_T.reserve(B * sT);
        d_adj.reserve(B * sA);
        d_X.zero(B * sX, stream);  // the rows between a small instance and the next one must read as zero in the mix
        d_Y.reserve(B * sY);
        d_Z.reserve(B * sY);
        d_K.reserve(Kd * C_out);
        d_b.reserve(C_out);
        d_n.reserve(B);
        n_host.resize(B);
        for (size_t i = 0; i < B; ++i) {
            const int n = (int)items[i].tensors.size();
            n_host[i] = n;
            const size_t slab = (size_t)n * n * C_in;
            for (int a = 0; a < n; ++a) d_T.upload(items[i].tensors[a]->value, slab, i * sT + a * slab, stream);
            d_adj.upload(items[i].adj->value, (size_t)n * n, i * sA, stream);
        }
        CCN_B200_CHECK(ctx, ccn_h2d(ctx, d_n.dev, &n_host[0], B * sizeof(int32_t), stream));
        d_K.upload(K->value, Kd * C_out, 0, stream);
        d_b.upload(b->value, C_out, 0, stream);
        CCN_B200_CHECK(ctx, ccn_contract18_forward(ctx, d_T.dev, NULL, d_adj.dev, d_X.dev, reinterpret_cast<const int32_t *>(d_n.dev), max_N,
                                                   C_in, (int64_t)B, (int64_t)sT, (int64_t)sA, (int64_t)sX, CCN_ADJ_POSITIVE_PART, stream));
        CCN_B200_CHECK(ctx, ccn_mix_forward(ctx, d_X.dev, d_K.dev, d_b.dev, d_Y.dev, d_Z.dev, (int64_t)(B * N * N), (int)Kd, C_out, alpha,
                                            stream));
        for (size_t i = 0; i < B; ++i) {
            Tensor3D *o = items[i].out;
            d_Z.download(o->value, (size_t)o->size, i * sY, stream);
            for (int j = 0; j < o->size; ++j) o->gradient[j] = 0.0;
        }
    }

    void backward() {
        const size_t B = items.size();
        if (B == 0) return;
        ccn_ctx *ctx = context();
        const size_t N = max_N, sT = N * N * N * C_in, sA = N * N, Kd = (size_t)nContractions * C_in, sX = N * N * Kd, sY = N * N * C_out;
        d_gZ.zero(B * sY, stream);  // rows of the padding carry no gradient
        for (size_t i = 0; i < B; ++i) d_gZ.upload(items[i].out->gradient, (size_t)items[i].out->size, i * sY, stream);
        d_gX.reserve(B * sX);
        d_gK.zero(Kd * C_out, stream);
        d_gb.zero(C_out, stream);
        d_gT.reserve(B * sT);
        CCN_B200_CHECK(ctx, ccn_mix_backward(ctx, d_X.dev, d_K.dev, d_b.dev, d_Y.dev, d_gZ.dev, d_gX.dev, d_gK.dev, d_gb.dev,
                                             (int64_t)(B * N * N), (int)Kd, C_out, alpha, 0.0f, stream));
        CCN_B200_CHECK(ctx, ccn_contract18_backward(ctx, d_gX.dev, d_adj.dev, d_gT.dev, NULL, reinterpret_cast<const int32_t *>(d_n.dev), max_N,
                                                    C_in, (int64_t)B, (int64_t)sX, (int64_t)sA, (int64_t)sT, CCN_ADJ_POSITIVE_PART, 0.0f, stream));
        for (size_t i = 0; i < B; ++i) {
            const int n = n_host[i];
            const size_t slab = (size_t)n * n * C_in;
            for (int a = 0; a < n; ++a) d_gT.download_add(items[i].tensors[a]->gradient, slab, i * sT + a * slab, stream);
        }
        d_gK.download_add(K->gradient, Kd * C_out, 0, stream);
        d_gb.download_add(b->gradient, C_out, 0, stream);
    }
    void release() {
        DeviceArray *all[] = {&d_T, &d_adj, &d_K, &d_b, &d_X, &d_Y, &d_Z, &d_gZ, &d_gX, &d_gK, &d_gb, &d_gT, &d_n};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
    }

    int max_N, C_in, C_out;
    Matrix *K;
    Vector *b;
    float alpha;
    cudaStream_t stream;
    static const int nContractions = 18;

private:
    struct Item {
        Tensor3D *out;
        std::vector<Tensor3D *> tensors;
        Matrix *adj;
    };
    std::vector<Item> items;
    std::vector<int32_t> n_host;
    DeviceArray d_T, d_adj, d_K, d_b, d_X, d_Y, d_Z, d_gZ, d_gX, d_gK, d_gb, d_gT, d_n;
};

// ---------------------------------------------------------------------------------------------------------------
// Executor: the add / clear / forward / backward surface of GraphFlow (GraphFlow.h:176-188, 190-727, 729-1266) as an
// ordered list of type-erased entries, so that the new operators run in the same topology as unmodified reference
// operators without patching the reference's tag ladder (unknown tags would be skipped silently there).
// ---------------------------------------------------------------------------------------------------------------
enum {  // tags of the new operators; the reference uses 0..101 (GraphFlow_gpu/GraphFlow.h:96-180)
    RISICONTRACTION_18_B200 = 200,
    STACKTENSOR3D_B200 = 201,
    MATMUL_B200 = 202,
    CCNLEVEL_B200 = 203,
    RISICONTRACTION_18_HOSTAPI_B200 = 204,
    LEVELBATCH_B200 = 205
};

class Executor {
    struct IOp {
        virtual void forward() = 0;
        virtual void backward() = 0;
        virtual ~IOp() {}
    };
    template <class T>
    struct RefOp : IOp {  // adapts any op with forward()/backward(), reference or ours
        T *p;
        explicit RefOp(T *p_) : p(p_) {}
        void forward() { p->forward(); }
        void backward() { p->backward(); }
    };

public:
    ~Executor() { clear(); }
    template <class T>
    void add(T *op, int tag = -1) {
        ops.push_back(new RefOp<T>(op));
        tags.push_back(tag);
    }
    void clear() {
        for (size_t i = 0; i < ops.size(); ++i) delete ops[i];
        ops.clear();
        tags.clear();
    }
    void forward() {
        for (size_t i = 0; i < ops.size(); ++i) ops[i]->forward();
    }
    void backward() {
        for (size_t i = ops.size(); i-- > 0;) ops[i]->backward();
    }
    size_t size() const { return ops.size(); }
    std::vector<int> tags;

private:
    std::vector<IOp *> ops;
};

}  // namespace ccn_b200

// Opt-in header-compatible spelling: with -DCCN_B200_DROP_IN (and the reference's own *_gpu headers NOT included)
// a model header written against GraphFlow_gpu/ compiles against this implementation unchanged.
#ifdef CCN_B200_DROP_IN
typedef ccn_b200::RisiContraction_18_gpu RisiContraction_18_gpu;
typedef ccn_b200::MatMul_gpu MatMul_gpu;
#endif

namespace ccn_b200 {
// the reference's threaded stack (StackTensor3D_thread.h:24-189) has the same interface as StackTensor3D
typedef StackTensor3D StackTensor3D_thread;
}  // namespace ccn_b200

#endif  // GRAPHFLOW_B200_CCN_OPS_B200_H_INCLUDED
