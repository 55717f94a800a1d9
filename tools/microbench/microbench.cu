// tools/microbench/microbench.cu — measured denominators for the rooflines in DESIGN.md:
//   (1) IMAD.WIDE.U32 issue rate (independent chains, no carries)
//   (2) carry-chained IMAD.WIDE.U32.X rows (what a Montgomery product issues)
//   (3) Fq Montgomery products per second (the unit the MSM/NTT kernels are counted in)
//   (4) random 64-byte gathers per second from tables of 0.25 .. 48 GiB (fixed-base window tables),
//       lanes fully random vs. lanes inside one 2 MiB slab per warp-step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../libzkp_b200/csrc/field.cuh"
using namespace lzkp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_imad_wide(uint64_t *out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = threadIdx.x + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a + j), "r"(b));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_imad_chain(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t X[8], Y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { X[j] = threadIdx.x + j; Y[j] = threadIdx.x * 3 + j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            X[7] += row_mad(X, a, a + 1, a + 2, a + 3, b + u);
            Y[7] += row_mad(Y, a + 4, a + 5, a + 6, a + 7, b + u);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= X[j] ^ Y[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_montmul(Fq *out, const Fq *in, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x[ILP], y = in[t];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = in[t + j + 1];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = x[j] * y;
    }
    Fq s = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) s = s + x[j];
    out[t] = s;
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// mode 0: every lane anywhere in the table; mode 1: warp picks a random 2 MiB slab, lanes random inside it
__global__ void __launch_bounds__(256) k_gather(const uint4 *tbl, uint64_t n_entries, int iters, int mode, uint4 *out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t warp = t >> 5;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint64_t slab = (2u << 20) / 64;     // entries per 2 MiB
    const uint64_t n_slabs = n_entries / slab;
    for (int i = 0; i < iters; i++) {
        uint64_t e;
        if (mode == 0) {
            uint64_t h = ((uint64_t)mix(t * 2654435761u + i) << 32) | mix(t + 0x9e3779b9u * i);
            e = h % n_entries;
        } else {
            uint64_t s = (((uint64_t)mix(warp * 2654435761u + i) << 16) ^ mix(warp + 77u * i)) % n_slabs;
            e = s * slab + (mix(t * 40503u + i) % slab);
        }
        const uint4 *p = tbl + e * 4;
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
        acc.x ^= v0.x ^ v1.x ^ v2.x ^ v3.x; acc.y ^= v0.y ^ v1.y ^ v2.y ^ v3.y;
        acc.z ^= v0.z ^ v1.z ^ v2.z ^ v3.z; acc.w ^= v0.w ^ v1.w ^ v2.w ^ v3.w;
    }
    out[t] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char **argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void *buf; CK(cudaMalloc(&buf, 64 << 20));
    CK(cudaMemset(buf, 1, 64 << 20));
    const int blocks = sms * 8, threads = 256;
    {
        int iters = 2000;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k_imad_wide<<<blocks, threads>>>((uint64_t *)buf, 12345, 6789, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            double ops = (double)blocks * threads * iters * 16 * 8;
            printf("{\"bench\": \"imad_wide_independent\", \"rep\": %d, \"ms\": %.3f, \"Tops_per_s\": %.3f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", rep,
                   time_ms(e0, e1), ops / time_ms(e0, e1) / 1e9, ops / (time_ms(e0, e1) * 1e-3) / sms / 1.965e9);
        }
    }
    {
        int iters = 2000;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k_imad_chain<<<blocks, threads>>>((uint32_t *)buf, 12345, 6789, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            double ops = (double)blocks * threads * iters * 8 * 2 * 4;
            printf("{\"bench\": \"imad_wide_carry_rows\", \"rep\": %d, \"ms\": %.3f, \"Tops_per_s\": %.3f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", rep,
                   time_ms(e0, e1), ops / time_ms(e0, e1) / 1e9, ops / (time_ms(e0, e1) * 1e-3) / sms / 1.965e9);
        }
    }
    {
        int iters = 1000;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k_montmul<1><<<blocks, threads>>>((Fq *)buf, (const Fq *)buf, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            double ops = (double)blocks * threads * iters;
            printf("{\"bench\": \"fq_montmul_ilp1\", \"rep\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", rep, time_ms(e0, e1), ops / time_ms(e0, e1) / 1e6);
            cudaEventRecord(e0); k_montmul<2><<<blocks, threads>>>((Fq *)buf, (const Fq *)buf, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            printf("{\"bench\": \"fq_montmul_ilp2\", \"rep\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", rep, time_ms(e0, e1), 2 * ops / time_ms(e0, e1) / 1e6);
        }
        for (int bl = 1; bl <= 16; bl *= 2) {      // occupancy sweep: blocks per SM
            cudaEventRecord(e0); k_montmul<1><<<sms * bl, 128>>>((Fq *)buf, (const Fq *)buf, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            double ops = (double)sms * bl * 128 * iters;
            printf("{\"bench\": \"fq_montmul_occupancy\", \"warps_per_sm\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", bl * 4, time_ms(e0, e1), ops / time_ms(e0, e1) / 1e6);
        }
    }
    {
        double gib[] = {0.25, 2, 16, 48};
        uint4 *out; CK(cudaMalloc(&out, (size_t)sms * 16 * 256 * 16));
        for (double g : gib) {
            size_t bytes = (size_t)(g * (1ull << 30));
            void *tbl;
            if (cudaMalloc(&tbl, bytes) != cudaSuccess) { printf("{\"bench\": \"gather\", \"gib\": %.2f, \"error\": \"alloc\"}\n", g); cudaGetLastError(); continue; }
            CK(cudaMemset(tbl, 3, bytes));
            for (int mode = 0; mode < 2; mode++) {
                int iters = 256;
                int gb = sms * 16;
                k_gather<<<gb, 256>>>((const uint4 *)tbl, bytes / 64, 16, mode, out);
                cudaEventRecord(e0); k_gather<<<gb, 256>>>((const uint4 *)tbl, bytes / 64, iters, mode, out); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                double n = (double)gb * 256 * iters;
                printf("{\"bench\": \"gather64B\", \"gib\": %.2f, \"mode\": \"%s\", \"ms\": %.3f, \"Glookups_per_s\": %.3f, \"GBps\": %.1f}\n", g,
                       mode ? "slab2MiB_per_warp" : "fully_random", time_ms(e0, e1), n / time_ms(e0, e1) / 1e6, n * 64 / time_ms(e0, e1) / 1e6);
            }
            CK(cudaFree(tbl));
        }
    }
    return 0;
}
