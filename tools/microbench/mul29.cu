// tools/microbench/mul29.cu — does an unsaturated-limb Montgomery product (9 x 29-bit limbs, 64-bit column
// accumulators, carry-free IMAD.WIDE at full issue rate) beat the 8 x 32-bit CIOS product whose carry-chained
// IMAD.WIDE.X rows issue at half rate?  Prints products/s for both.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mul29 mul29.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../libzkp_b200/csrc/field.cuh"
using namespace lzkp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr uint32_t M29 = (1u << 29) - 1;
// q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 in 29-bit limbs
__device__ __constant__ uint32_t P29c[9];
struct F29 { uint32_t l[9]; };

template <bool CONSTP>
__device__ __forceinline__ F29 mul29(const F29 &a, const F29 &b, const uint32_t (&P)[9], uint32_t pinv) {
    uint64_t t[18];
#pragma unroll
    for (int k = 0; k < 18; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) t[i + j] += (uint64_t)a.l[i] * b.l[j];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint32_t m = ((uint32_t)t[i] * pinv) & M29;
#pragma unroll
        for (int j = 0; j < 9; j++) t[i + j] += (uint64_t)m * P[j];
        t[i + 1] += t[i] >> 29;
    }
    F29 r;
#pragma unroll
    for (int k = 9; k < 17; k++) { r.l[k - 9] = (uint32_t)t[k] & M29; t[k + 1] += t[k] >> 29; }
    r.l[8] = (uint32_t)t[17];
    return r;
}

__global__ void __launch_bounds__(256) k_mul29(F29 *out, const F29 *in, int iters, uint32_t pinv) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t P[9];
#pragma unroll
    for (int i = 0; i < 9; i++) P[i] = P29c[i];
    F29 x = in[t], y = in[t + 1];
    for (int i = 0; i < iters; i++) x = mul29<false>(x, y, P, pinv);
    out[t] = x;
}
__global__ void __launch_bounds__(256) k_mul32(Fq *out, const Fq *in, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[t], y = in[t + 1];
    for (int i = 0; i < iters; i++) x = x * y;
    out[t] = x;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    // p in 29-bit limbs and -p^-1 mod 2^29 (computed here with 128-bit host arithmetic on the low limb)
    unsigned __int128 lo = ((unsigned __int128)0x97816a916871ca8dULL << 64) | 0x3c208c16d87cfd47ULL;
    unsigned __int128 hi = ((unsigned __int128)0x30644e72e131a029ULL << 64) | 0xb85045b68181585dULL;
    uint32_t P[9];
    for (int i = 0; i < 9; i++) {
        int bit = 29 * i;
        unsigned __int128 v;
        if (bit < 128) { v = lo >> bit; if (bit + 29 > 128) v |= hi << (128 - bit); }
        else v = hi >> (bit - 128);
        P[i] = (uint32_t)v & M29;
    }
    uint32_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - P[0] * inv;       // p^-1 mod 2^32
    uint32_t pinv = (0u - inv) & M29;
    CK(cudaMemcpyToSymbol(P29c, P, sizeof(P)));
    const int blocks = sms * 8, threads = 256, iters = 2000;
    void *in, *out;
    CK(cudaMalloc(&in, (size_t)(blocks * threads + 2) * 36)); CK(cudaMalloc(&out, (size_t)(blocks * threads + 2) * 36));
    CK(cudaMemset(in, 0x15, (size_t)(blocks * threads + 2) * 36));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k_mul29<<<blocks, threads>>>((F29 *)out, (const F29 *)in, iters, pinv); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"bench\": \"montmul_9x29_carryfree\", \"rep\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", rep, ms, (double)blocks * threads * iters / ms / 1e6);
        cudaEventRecord(e0); k_mul32<<<blocks, threads>>>((Fq *)out, (const Fq *)in, iters); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"bench\": \"montmul_8x32_cios\", \"rep\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", rep, ms, (double)blocks * threads * iters / ms / 1e6);
    }
    return 0;
}
