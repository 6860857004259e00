// profiles/scatter_gather_probe.cu -- EXPERIMENT (not product code): is the promotion backward cheaper as a GATHER than as a
// SCATTER?  DESIGN.md section 9 item 2a.
//
// The fused backward adds every element of gT_v[a, b, c, f] into gf[w = phi_v[a]][pos(b), pos(c), f] with RED.ADD.F32: N^3 C
// reductions per instance, bound by the L2 reduction rate.  The alternative keeps phase 1 (the N^2 C planes U, g6 | V, G10 of every
// instance, 1 MB per instance, L2-resident per graph) and lets a tile (graph, source vertex w, row block) SUM
//     gT_v[a_v(w), b, c] = U_v[a,b] + V_v[b,c] + g6_v[a,b] r_v[c] + r_v[a] G10_v[b,c]
// over the vertices v of the graph in registers, reading V_v / G10_v through the inverse position tables, and write gf[w] once.
// This probe times both memory patterns on synthetic planes (full fields: every v contains every w; one random permutation per
// instance): the scatter emulation does the same arithmetic as the gather one and the same number of reductions as the real kernel.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/sgp profiles/scatter_gather_probe.cu && /tmp/sgp [graphs]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

constexpr int N = 32, C = 64, TB = 4, THREADS = TB * C;
#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e__ = (x);                                                     \
        if (e__ != cudaSuccess) {                                                  \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__));         \
            std::exit(1);                                                          \
        }                                                                          \
    } while (0)

// planes of instance i: P[i][k][x][y][f], k = 0: U[a,b], 1: g6[a,b], 2: V[b,c], 3: G10[b,c]
__device__ __forceinline__ const float *plane(const float *P, int64_t inst, int k) { return P + (inst * 4 + k) * (int64_t)N * N * C; }

// gather: CTA = (graph, w, row block of TB source rows i'); thread = (row i', channel f)
__global__ void __launch_bounds__(THREADS) k_gather(const float *__restrict__ P, const float *__restrict__ r, const unsigned char *__restrict__ inv,
                                                    float *__restrict__ gf) {
    __shared__ unsigned char s_inv[N][N];  // inverse position table of every instance v of the graph: s_inv[v][i'] = b
    __shared__ float s_r[N][N];
    const int g = blockIdx.z, w = blockIdx.y, i0 = blockIdx.x * TB;
    const int f = threadIdx.x % C, il = threadIdx.x / C, ip = i0 + il;
    for (int i = threadIdx.x; i < N * N; i += THREADS) {
        s_inv[i / N][i % N] = inv[(int64_t)g * N * N + i];
        s_r[i / N][i % N] = r[(int64_t)g * N * N + i];
    }
    __syncthreads();
    float acc[N];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j] = 0.f;
    for (int v = 0; v < N; ++v) {
        const int64_t inst = (int64_t)g * N + v;
        const int a = s_inv[v][w];  // the slot of w in v's field (full fields: some permutation)
        const int b = s_inv[v][ip];
        const float u = plane(P, inst, 0)[((int64_t)a * N + b) * C + f], g6 = plane(P, inst, 1)[((int64_t)a * N + b) * C + f];
        const float ra = s_r[v][a];
        const float *Vb = plane(P, inst, 2) + (int64_t)b * N * C + f, *Gb = plane(P, inst, 3) + (int64_t)b * N * C + f;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int c = s_inv[v][j];
            acc[j] += u + __ldg(Vb + c * C) + g6 * s_r[v][c] + ra * __ldg(Gb + c * C);
        }
    }
    float *o = gf + (((int64_t)g * N + w) * N + ip) * (int64_t)N * C + f;
#pragma unroll
    for (int j = 0; j < N; ++j) __stcs(o + j * C, acc[j]);
}

// scatter: CTA = (graph, v, row block of TB rows b); thread = (row b, channel f); for every slab a: N reductions per thread
__global__ void __launch_bounds__(THREADS) k_scatter(const float *__restrict__ P, const float *__restrict__ r, const unsigned char *__restrict__ pos,
                                                     float *__restrict__ gf) {
    __shared__ unsigned char s_pos[N];
    __shared__ float s_r[N];
    const int g = blockIdx.z, v = blockIdx.y, b0 = blockIdx.x * TB;
    const int f = threadIdx.x % C, bl = threadIdx.x / C, b = b0 + bl;
    const int64_t inst = (int64_t)g * N + v;
    if (threadIdx.x < N) {
        s_pos[threadIdx.x] = pos[inst * N + threadIdx.x];
        s_r[threadIdx.x] = r[inst * N + threadIdx.x];
    }
    __syncthreads();
    float V[N], G10[N];
    const float *Vb = plane(P, inst, 2) + (int64_t)b * N * C + f, *Gb = plane(P, inst, 3) + (int64_t)b * N * C + f;
#pragma unroll
    for (int c = 0; c < N; ++c) V[c] = __ldg(Vb + c * C), G10[c] = __ldg(Gb + c * C);
    const int pb = s_pos[b];
    for (int a = 0; a < N; ++a) {
        const float u = plane(P, inst, 0)[((int64_t)a * N + b) * C + f], g6 = plane(P, inst, 1)[((int64_t)a * N + b) * C + f];
        const float ra = s_r[a];
        const int w = s_pos[a];  // slab a comes from vertex w
        float *o = gf + (((int64_t)g * N + w) * N + pb) * (int64_t)N * C + f;
#pragma unroll
        for (int c = 0; c < N; ++c) atomicAdd(o + s_pos[c] * C, u + V[c] + g6 * s_r[c] + ra * G10[c]);
    }
}

int main(int argc, char **argv) {
    const int G = argc > 1 ? std::atoi(argv[1]) : 64;
    const int64_t inst = (int64_t)G * N;
    const size_t pbytes = (size_t)inst * 4 * N * N * C * sizeof(float), obytes = (size_t)inst * N * N * C * sizeof(float);
    float *P, *r, *gf1, *gf2;
    unsigned char *pos, *inv;
    CK(cudaMalloc(&P, pbytes));
    CK(cudaMalloc(&r, inst * N * sizeof(float)));
    CK(cudaMalloc(&gf1, obytes));
    CK(cudaMalloc(&gf2, obytes));
    CK(cudaMalloc(&pos, inst * N));
    CK(cudaMalloc(&inv, inst * N));
    std::vector<float> hP((size_t)inst * 4 * N * N * C), hr(inst * N);
    std::vector<unsigned char> hpos(inst * N), hinv(inst * N);
    srand(1);
    for (auto &x : hP) x = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &x : hr) x = (float)(rand() % 5);
    for (int64_t i = 0; i < inst; ++i) {
        for (int k = 0; k < N; ++k) hpos[i * N + k] = (unsigned char)k;
        for (int k = N - 1; k > 0; --k) std::swap(hpos[i * N + k], hpos[i * N + rand() % (k + 1)]);
        for (int k = 0; k < N; ++k) hinv[i * N + hpos[i * N + k]] = (unsigned char)k;
    }
    CK(cudaMemcpy(P, hP.data(), pbytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(r, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pos, hpos.data(), hpos.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(inv, hinv.data(), hinv.size(), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const dim3 grid(N / TB, N, G);
    float ms_g = 0, ms_s = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k_gather<<<grid, THREADS>>>(P, r, inv, gf1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_g, e0, e1));
        CK(cudaMemset(gf2, 0, obytes));
        CK(cudaEventRecord(e0));
        k_scatter<<<grid, THREADS>>>(P, r, pos, gf2);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_s, e0, e1));
    }
    CK(cudaGetLastError());
    // the two kernels compute the same sums (different summation order)
    std::vector<float> h1(1 << 16), h2(1 << 16);
    CK(cudaMemcpy(h1.data(), gf1, h1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), gf2, h2.size() * 4, cudaMemcpyDeviceToHost));
    double md = 0, mx = 0;
    for (size_t i = 0; i < h1.size(); ++i) {
        md = std::max(md, (double)std::fabs(h1[i] - h2[i]));
        mx = std::max(mx, (double)std::fabs(h1[i]));
    }
    std::printf("{\"graphs\": %d, \"instances\": %lld, \"gather_ms\": %.3f, \"scatter_ms\": %.3f, \"gather_ms_per_512\": %.3f, "
                "\"scatter_ms_per_512\": %.3f, \"max_rel_diff\": %.2e}\n",
                G, (long long)inst, ms_g, ms_s, ms_g * 512.0 / inst, ms_s * 512.0 / inst, md / mx);
    return 0;
}
