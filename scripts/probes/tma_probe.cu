// tma_probe.cu -- the static update kernel's traffic (read A, V; write A, V, O; float4 each) with the
// loads staged through shared memory by 1-D bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP) in a
// STAGES-deep ring per CTA, results stored with ordinary STG.128 (variant 1) or staged and bulk-stored
// (variant 2). Does it beat the plain LDG/STG walk of order_probe.cu (5.7 - 6.1 TB/s)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/tma_probe scripts/probes/tma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(phase)
        : "memory");
}

template <int STAGES, int STORE_BULK>
__global__ void __launch_bounds__(256) walk(float4 *A, float4 *V, float4 *O, size_t n_tiles, int contiguous) {
    extern __shared__ __align__(128) uint8_t smem[];
    float4 *sA = (float4 *)smem;                      // [STAGES][256]
    float4 *sV = sA + STAGES * 256;                   // [STAGES][256]
    float4 *sO = sV + STAGES * 256;                   // [2][3][256] output staging (STORE_BULK)
    uint64_t *bar = (uint64_t *)(sO + (STORE_BULK ? 2 * 3 * 256 : 0));
    const size_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
    auto tile_of = [&](size_t k) { return contiguous ? (size_t)blockIdx.x * per + k : (size_t)blockIdx.x + k * gridDim.x; };
    size_t mine = 0;
    while (mine < per && tile_of(mine) < n_tiles) mine++;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t k) {
        const int s = (int)(k % STAGES);
        const size_t i = tile_of(k) * 256;
        mbar_expect(&bar[s], 2 * 4096);
        bulk_load(sA + s * 256, A + i, 4096, &bar[s]);
        bulk_load(sV + s * 256, V + i, 4096, &bar[s]);
    };
    if (threadIdx.x == 0)
        for (size_t k = 0; k < (size_t)STAGES && k < mine; k++) issue(k);
    for (size_t k = 0; k < mine; k++) {
        const int s = (int)(k % STAGES);
        mbar_wait(&bar[s], (uint32_t)((k / STAGES) & 1));
        float4 a = sA[s * 256 + threadIdx.x], v = sV[s * 256 + threadIdx.x];
        a.x += v.x * 0.016f; a.y += v.y * 0.016f; a.z += v.z * 0.016f; a.w += 0.016f;
        v.x *= 0.99f; v.y = v.y * 0.99f - 0.1f; v.z *= 0.99f;
        const float4 o = make_float4(a.w, a.w * 0.5f, 1.0f - a.w, 1.0f);
        const size_t i = tile_of(k) * 256 + threadIdx.x;
        if (STORE_BULK) {
            const int b = (int)(k & 1);
            if (k >= 2) { // the bulk stores that last read this staging buffer have finished reading it
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncthreads();
            }
            sO[(b * 3 + 0) * 256 + threadIdx.x] = a;
            sO[(b * 3 + 1) * 256 + threadIdx.x] = v;
            sO[(b * 3 + 2) * 256 + threadIdx.x] = o;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads(); // (also: every thread has read stage s)
            if (threadIdx.x == 0) {
                const size_t t0 = tile_of(k) * 256;
                bulk_store(A + t0, sO + (b * 3 + 0) * 256, 4096);
                bulk_store(V + t0, sO + (b * 3 + 1) * 256, 4096);
                bulk_store(O + t0, sO + (b * 3 + 2) * 256, 4096);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (k + STAGES < mine) issue(k + STAGES);
            }
        } else {
            A[i] = a;
            V[i] = v;
            O[i] = o;
            __syncthreads(); // every thread has read stage s
            if (threadIdx.x == 0 && k + STAGES < mine) issue(k + STAGES);
        }
    }
    if (STORE_BULK && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int STAGES, int STORE_BULK>
void run(float4 *A, float4 *V, float4 *O, size_t n, int ctas_per_sm) {
    const size_t n_tiles = n / 256;
    const int smem = STAGES * 2 * 4096 + (STORE_BULK ? 2 * 3 * 4096 : 0) + STAGES * 8 + 64;
    cudaFuncSetAttribute(walk<STAGES, STORE_BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int contiguous = 0; contiguous < 2; contiguous++) {
        const int grid = 148 * ctas_per_sm;
        for (int w = 0; w < 3; w++) walk<STAGES, STORE_BULK><<<grid, 256, smem>>>(A, V, O, n_tiles, contiguous);
        cudaEventRecord(e0);
        const int reps = 20;
        for (int r = 0; r < reps; r++) walk<STAGES, STORE_BULK><<<grid, 256, smem>>>(A, V, O, n_tiles, contiguous);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("stages %d %-11s %d CTAs/SM %-18s %.4f ms  %.0f GB/s  (%s)\n", STAGES, STORE_BULK ? "bulk stores" : "STG stores", ctas_per_sm,
               contiguous ? "contiguous shares" : "grid-stride tiles", ms, n * 80.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    const size_t n = 10u * 1000u * 1000u / 256u * 256u;
    float4 *A, *V, *O;
    cudaMalloc(&A, n * 16); cudaMalloc(&V, n * 16); cudaMalloc(&O, n * 16);
    cudaMemset(A, 0, n * 16); cudaMemset(V, 0, n * 16);
    for (int c : {2, 3, 4, 5}) {
        run<4, 0>(A, V, O, n, c);
        run<4, 1>(A, V, O, n, c);
    }
    run<8, 0>(A, V, O, n, 2);
    run<8, 1>(A, V, O, n, 2);
    run<2, 1>(A, V, O, n, 5);
    return 0;
}
