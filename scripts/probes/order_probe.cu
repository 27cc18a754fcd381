// order_probe.cu -- does the ORDER in which a persistent grid walks three float4 arrays matter to HBM?
// Same traffic as the static update kernel (read A, V; write A, V, O: 32 B in, 48 B out per element),
// trivial arithmetic. (a) grid-stride tiles: CTA b takes tile b, b+grid, ... (the whole grid sweeps one
// moving window); (b) contiguous share per CTA: 740 independent sequential streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/order_probe scripts/probes/order_probe.cu && /tmp/order_probe
#include <cstdio>
#include <cuda_runtime.h>

// team > 0: teams of `team` consecutive CTAs share one contiguous range and deal its tiles round robin
__global__ void __launch_bounds__(256, 5) walk(float4 *A, float4 *V, float4 *O, size_t n_tiles, int contiguous, int team) {
    const size_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
    for (size_t k = 0; k < per; k++) {
        size_t tile = contiguous ? (size_t)blockIdx.x * per + k : (size_t)blockIdx.x + k * gridDim.x;
        if (team > 0) tile = (size_t)(blockIdx.x / team) * per * team + (blockIdx.x % team) + k * team;
        if (tile >= n_tiles) break;
        const size_t i = tile * 256 + threadIdx.x;
        float4 a = A[i], v = V[i];
        a.x += v.x * 0.016f; a.y += v.y * 0.016f; a.z += v.z * 0.016f; a.w += 0.016f;
        v.x *= 0.99f; v.y = v.y * 0.99f - 0.1f; v.z *= 0.99f;
        A[i] = a;
        V[i] = v;
        O[i] = make_float4(a.w, a.w * 0.5f, 1.0f - a.w, 1.0f);
    }
}

int main() {
    const size_t n = 10u * 1000u * 1000u / 256u * 256u, n_tiles = n / 256;
    float4 *A, *V, *O;
    cudaMalloc(&A, n * 16); cudaMalloc(&V, n * 16); cudaMalloc(&O, n * 16);
    cudaMemset(A, 0, n * 16); cudaMemset(V, 0, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {740, 592}) {
        for (int contiguous = 0; contiguous < 2; contiguous++) {
            for (int w = 0; w < 3; w++) walk<<<grid, 256>>>(A, V, O, n_tiles, contiguous, 0);
            cudaEventRecord(e0);
            const int reps = 20;
            for (int r = 0; r < reps; r++) walk<<<grid, 256>>>(A, V, O, n_tiles, contiguous, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            printf("grid %4d %-22s %.4f ms  %.0f GB/s\n", grid, contiguous ? "contiguous shares" : "grid-stride tiles", ms, n * 80.0 / (ms * 1e-3) / 1e9);
        }
        for (int team : {2, 4, 8, 16, 37, 74, 148}) {
            if (grid % team) continue;
            for (int w = 0; w < 3; w++) walk<<<grid, 256>>>(A, V, O, n_tiles, 1, team);
            cudaEventRecord(e0);
            const int reps = 20;
            for (int r = 0; r < reps; r++) walk<<<grid, 256>>>(A, V, O, n_tiles, 1, team);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            printf("grid %4d teams of %-13d %.4f ms  %.0f GB/s\n", grid, team, ms, n * 80.0 / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
