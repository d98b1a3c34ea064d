// microbench.cu -- measured denominators that MEASURED_PEAKS.json lacks: FP64 FMA peak, FP64 global
// atomic (RED) throughput for scattered / run-of-4 / coalesced addresses, shared-memory RMW throughput.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_dfma(double *out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: scattered (hash), 1: runs of 4 consecutive doubles per 4 lanes, 2: fully coalesced per warp, 3: all lanes same address
__global__ void k_red(double *buf, size_t n, int iters, int mode) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    size_t warp = tid >> 5;
    uint64_t h = warp * 0x9E3779B97F4A7C15ull + 12345;
    for (int it = 0; it < iters; ++it) {
        h = h * 6364136223846793005ull + 1442695040888963407ull;
        size_t idx;
        if (mode == 0) { uint64_t hl = (h ^ (lane * 0xD6E8FEB86659FD93ull)) * 0xFF51AFD7ED558CCDull; idx = (hl >> 20) % n; }
        else if (mode == 1) { uint64_t hl = (h ^ ((lane >> 2) * 0xD6E8FEB86659FD93ull)) * 0xFF51AFD7ED558CCDull; idx = ((hl >> 20) % (n / 4)) * 4 + (lane & 3); }
        else if (mode == 2) { idx = ((h >> 20) % (n / 32)) * 32 + lane; }
        else { idx = ((h >> 20) % n); }
        atomicAdd(buf + idx, 1.0);
    }
}

__global__ void k_smem_rmw(double *out, int iters) {
    extern __shared__ double s[];
    const int n = 16 * 1024;   // 128 KB
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 0;
    __syncthreads();
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        // 4 independent conflict-free RMWs per iteration
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int j = (idx + u * 1024 + it * 32) & (n - 1);
            s[j] += 1.0;
        }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    {
        const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
        k_dfma<<<blocks, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0)); k_dfma<<<blocks, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf(", \"fp64_fma_tflops\": %.2f", flops / (best * 1e-3) / 1e12);
    }
    {
        const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 5000;
        k_dmma<<<blocks, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0)); k_dmma<<<blocks, threads>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        // one m8n8k4 MMA per warp = 8*8*4 FMA = 512 flop
        double flops = 512.0 * 8 * iters * (double)blocks * threads / 32;
        printf(", \"fp64_mma884_tflops\": %.2f", flops / (best * 1e-3) / 1e12);
    }
    {
        const size_t n = 1ull << 27;   // 1 GiB of doubles
        double *buf; CK(cudaMalloc(&buf, n * sizeof(double))); CK(cudaMemset(buf, 0, n * sizeof(double)));
        const int blocks = p.multiProcessorCount * 16, threads = 256, iters = 200;
        const char *names[4] = {"scattered", "runs_of_4", "coalesced_warp", "same_address_per_warp"};
        for (int mode = 0; mode < 4; ++mode) {
            k_red<<<blocks, threads>>>(buf, n, 10, mode); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); k_red<<<blocks, threads>>>(buf, n, iters, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf(", \"red_f64_%s_Gops\": %.2f", names[mode], (double)blocks * threads * iters / (ms * 1e-3) / 1e9);
        }
        // small footprint (L2 resident): 64 MB
        const size_t n2 = 1ull << 23;
        for (int mode = 0; mode < 3; ++mode) {
            k_red<<<blocks, threads>>>(buf, n2, 10, mode); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); k_red<<<blocks, threads>>>(buf, n2, iters, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf(", \"red_f64_l2_%s_Gops\": %.2f", names[mode], (double)blocks * threads * iters / (ms * 1e-3) / 1e9);
        }
        CK(cudaFree(buf));
    }
    {
        const int blocks = p.multiProcessorCount, threads = 1024, iters = 4000;
        CK(cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        k_smem_rmw<<<blocks, threads, 128 * 1024>>>(out, 10); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k_smem_rmw<<<blocks, threads, 128 * 1024>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"smem_f64_rmw_Gops\": %.2f", 4.0 * blocks * threads * iters / (ms * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}
