// microbench_fp64.cu -- FP64 pipe of B200 per instruction type and at the occupancy of the IMLS kernel.
//   throughput: DFMA / DMUL / DADD / DSETP+FSEL, warp instructions per cycle and SM sub-partition (8 independent chains, 8 warps)
//   latency   : one warp per sub-partition, ONE dependent chain -> cycles per dependent instruction
//   occupancy : W single-warp blocks per SM (W = 4 .. 16) with ILP 1, 2, 4 -> fraction of the DFMA peak reached
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int OP, int ILP>
__global__ void k_op(double *out, int iters, double b, double c) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i + 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8 / ILP; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = fma(a[i], b, c);
                else if (OP == 1) a[i] = a[i] * b;
                else if (OP == 2) a[i] = a[i] + c;
                else if (OP == 3) a[i] = a[i] > 0.5 ? a[i] * b : c;          // DSETP + DMUL + FSELx2
                else if (OP == 4) a[i] = fma(a[i], b, c) * b + c;            // DFMA, DMUL?, DADD mix as the compiler emits it (contracted)
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(t1 - t0) * 1e-30;
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long *>(out)[1 << 20] = t1 - t0;
}

template <int OP, int ILP>
static void run(const char *name, double *out, int blocks, int threads, int sms, double clk_ghz) {
    const int iters = 20000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_op<OP, ILP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
    float best = 1e30f, ms;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k_op<OP, ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    long long cyc; CK(cudaMemcpy(&cyc, reinterpret_cast<long long *>(out) + (1 << 20), 8, cudaMemcpyDeviceToHost));
    const double ops_per_thread = 8.0 * iters;                    // source-level operations (OP 3, 4: several instructions each)
    const double warp_ops = ops_per_thread * blocks * threads / 32.0;
    const double per_smsp_cycle = warp_ops / (sms * 4.0) / (double)cyc;
    printf("{\"name\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ilp\": %d, \"ms\": %.3f, \"cycles\": %lld, \"ops_per_cycle_per_subpartition\": %.4f, \"cycles_per_op_per_warp\": %.2f}\n",
           name, blocks, threads, ILP, best, cyc, per_smsp_cycle, (double)cyc / ops_per_thread);
}

__global__ void k_where(int *rec, int iters, double *out) {
    double a = threadIdx.x + 1.0;
    for (int it = 0; it < iters; ++it) a = fma(a, 1.0000001, 1e-9);      // long enough for every block of the grid to be resident at once
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) { rec[2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))] = (int)smid; rec[2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) + 1] = (int)warpid; }
    if (a == 0.123) out[0] = a;
}

static void where(const char *name, int blocks, int threads, int sms, double *out) {
    const int nw = blocks * threads / 32;
    int *rec; CK(cudaMalloc(&rec, sizeof(int) * 2 * nw));
    k_where<<<blocks, threads>>>(rec, 200000, out); CK(cudaDeviceSynchronize());
    int *h = (int *)malloc(sizeof(int) * 2 * nw);
    CK(cudaMemcpy(h, rec, sizeof(int) * 2 * nw, cudaMemcpyDeviceToHost));
    // warps per (SM, warpid mod 4): histogram of the per-sub-partition counts
    int *cnt = (int *)calloc(sms * 4 + 4, sizeof(int));
    for (int i = 0; i < nw; ++i) if (h[2 * i] < sms) cnt[h[2 * i] * 4 + (h[2 * i + 1] & 3)]++;
    int hist[33] = {0};
    for (int i = 0; i < sms * 4; ++i) hist[cnt[i] > 32 ? 32 : cnt[i]]++;
    printf("{\"name\": \"%s\", \"blocks\": %d, \"threads\": %d, \"subpartitions_with_k_warps\": {", name, blocks, threads);
    bool first = true;
    for (int k = 0; k <= 32; ++k) if (hist[k]) { printf("%s\"%d\": %d", first ? "" : ", ", k, hist[k]); first = false; }
    printf("}}\n");
    free(h); free(cnt); cudaFree(rec);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * ((1 << 20) + 16) + sms * 64 * 1024));
    // throughput: 8 warps per sub-partition, 8 independent chains
    run<0, 8>("dfma_throughput", out, sms * 4, 256, sms, 0);
    run<1, 8>("dmul_throughput", out, sms * 4, 256, sms, 0);
    run<2, 8>("dadd_throughput", out, sms * 4, 256, sms, 0);
    run<3, 8>("dsetp_dmul_fsel_throughput", out, sms * 4, 256, sms, 0);
    run<4, 8>("dfma_dfma_pair_throughput", out, sms * 4, 256, sms, 0);
    // latency: one warp per sub-partition, one chain
    run<0, 1>("dfma_latency", out, sms, 128, sms, 0);
    run<1, 1>("dmul_latency", out, sms, 128, sms, 0);
    run<2, 1>("dadd_latency", out, sms, 128, sms, 0);
    // one warp per sub-partition, ILP 2 / 4 / 8
    run<0, 2>("dfma_1warp_ilp2", out, sms, 128, sms, 0);
    run<0, 4>("dfma_1warp_ilp4", out, sms, 128, sms, 0);
    run<0, 8>("dfma_1warp_ilp8", out, sms, 128, sms, 0);
    // the IMLS occupancy: 12 single-warp blocks per SM = 3 per sub-partition
    run<0, 1>("dfma_3warps_ilp1", out, sms * 12, 32, sms, 0);
    run<0, 2>("dfma_3warps_ilp2", out, sms * 12, 32, sms, 0);
    run<0, 4>("dfma_3warps_ilp4", out, sms * 12, 32, sms, 0);
    run<1, 2>("dmul_3warps_ilp2", out, sms * 12, 32, sms, 0);
    run<2, 2>("dadd_3warps_ilp2", out, sms * 12, 32, sms, 0);
    // the same 3 warps per sub-partition as 128-thread blocks (the hardware spreads a block's warps over the four sub-partitions)
    run<0, 1>("dfma_3warps_ilp1_blocks_of_4_warps", out, sms * 3, 128, sms, 0);
    run<0, 2>("dfma_3warps_ilp2_blocks_of_4_warps", out, sms * 3, 128, sms, 0);
    run<0, 4>("dfma_3warps_ilp4_blocks_of_4_warps", out, sms * 3, 128, sms, 0);
    where("placement_single_warp_blocks_12_per_sm", sms * 12, 32, sms, out);
    where("placement_4_warp_blocks_3_per_sm", sms * 3, 128, sms, out);
    where("placement_2_warp_blocks_6_per_sm", sms * 6, 64, sms, out);
    return 0;
}
