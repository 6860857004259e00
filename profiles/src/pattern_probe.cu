// pattern_probe.cu -- what HBM bandwidth does the fused forward kernel's ACCESS PATTERN admit, with no arithmetic at all?
// Each CTA (256 threads, 2 per SM, ticketed tiles like k_fwd_fused) pulls 1 MiB of T as 32 bulk copies of 32 KiB through a
// 3-stage mbarrier ring (never touching the data), then writes its 576 KiB of `out`:
//   mode 0  the kernel's pattern: 18 slabs x 128 cells, each store instruction one 128-byte piece at a 4608-byte stride
//   mode 1  the same bytes written contiguously
//   mode 2  reads only        mode 3  scattered writes only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/_build/pattern_probe profiles/src/pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

constexpr int N = 32, C = 64, TB = 4, STAGES = 3, CHUNK = TB * N * C;  // floats per bulk copy (32 KiB)

__global__ void __launch_bounds__(256, 2) k_probe(const float *T, float *out, int *ticket, int tiles, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *ring = reinterpret_cast<float *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * CHUNK * 4);
    __shared__ int work;
    const int tid = threadIdx.x;
    if (tid == 0) work = atomicAdd(ticket, 1);
    __syncthreads();
    const int w = work;
    if (w >= tiles) return;
    const int inst = w / 8, tile = w % 8;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float *Ti = T + (size_t)inst * N * N * N * C + (size_t)tile * CHUNK;
    if (mode != 3) {
        if (tid == 0)
            for (int s = 0; s < STAGES; ++s) {
                mbar_expect(&full[s], CHUNK * 4);
                bulk_g2s(ring + s * CHUNK, Ti + (size_t)s * N * N * C, CHUNK * 4, &full[s]);
            }
        for (int a = 0; a < N; ++a) {
            mbar_wait(&full[a % STAGES], (a / STAGES) & 1);
            __syncthreads();
            if (tid == 0 && a + STAGES < N) {
                mbar_expect(&full[a % STAGES], CHUNK * 4);
                bulk_g2s(ring + (a % STAGES) * CHUNK, Ti + (size_t)(a + STAGES) * N * N * C, CHUNK * 4, &full[a % STAGES]);
            }
        }
    }
    if (mode == 2) return;
    const int f = tid % C, bl = tid / C;
    float *o = out + (size_t)inst * N * N * 18 * C + (size_t)(tile * TB) * N * 18 * C;
    const float v = (float)tid;
    if (mode == 0 || mode == 3) {
        for (int k = 0; k < 18; ++k)
            for (int y = 0; y < N; ++y) __stcs(o + ((size_t)(bl * N + y) * 18 + k) * C + f, v);
    } else {
        for (int i = 0; i < TB * N * 18 * C / 256; ++i) __stcs(o + (size_t)i * 256 + tid, v);
    }
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 512;
    float *T, *out;
    int *ticket;
    cudaMalloc(&T, (size_t)B * N * N * N * C * 4);
    cudaMalloc(&out, (size_t)B * N * N * 18 * C * 4);
    cudaMalloc(&ticket, 4);
    cudaMemset(T, 0, (size_t)B * N * N * N * C * 4);
    const size_t smem_bytes = STAGES * CHUNK * 4 + 64;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    const char *names[4] = {"read + scattered 128B writes (kernel pattern)", "read + contiguous writes", "read only", "scattered writes only"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(ticket, 0, 4);
            cudaEventRecord(e0);
            k_probe<<<B * 8, 256, smem_bytes>>>(T, out, ticket, B * 8, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double rd = mode == 3 ? 0 : (double)B * N * N * N * C * 4, wr = mode == 2 ? 0 : (double)B * N * N * 18 * C * 4;
        printf("mode %d  %-48s %.3f ms  %.0f GB/s  (%s)\n", mode, names[mode], best, (rd + wr) / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
