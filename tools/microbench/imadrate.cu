// tools/microbench/imadrate.cu — issue rates of the integer multiply-add forms a big-integer product can be built
// from, with the multiplicands in VECTOR registers (loaded per thread) and, for comparison, warp-uniform (kernel
// parameters, which ptxas keeps in uniform registers / the constant bank).  Prints lane-operations per SM per clock
// at 1965 MHz.  Forms: IMAD.WIDE.U32 (64-bit accumulate, no flags) | with carry-out predicate | .X carry-in+out |
// IMAD (32-bit lo) | IMAD.HI.U32 | DFMA | IADD3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o imadrate imadrate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum { WIDE = 0, WIDE_COUT = 1, WIDE_X = 2, LO = 3, HI = 4, DFMA = 5, IADD = 6, WIDE_ZC = 7 };

template <int F>
__device__ __forceinline__ void op(uint32_t &lo, uint32_t &hi, uint32_t &c, uint32_t a, uint32_t b) {
    if (F == WIDE) {
        uint64_t acc = ((uint64_t)hi << 32) | lo;
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
        lo = (uint32_t)acc; hi = (uint32_t)(acc >> 32);
    } else if (F == WIDE_ZC) {      // 64-bit product, accumulator operand is a zero-extended 32-bit word
        uint64_t acc = lo;
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
        lo = (uint32_t)acc; hi ^= (uint32_t)(acc >> 32);
    } else if (F == WIDE_COUT) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(c) : "r"(a), "r"(b));
    } else if (F == LO) {
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
    } else if (F == HI) {
        asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
    } else if (F == IADD) {
        asm("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(lo) : "r"(a), "r"(b));
    }
}

template <int F, bool VEC>
__global__ void __launch_bounds__(256) k_rate(uint32_t *out, const uint32_t *in, uint32_t ua, uint32_t ub, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { a[j] = VEC ? in[t * 8 + j] : ua + j; b[j] = VEC ? in[t * 8 + 4 + j] : ub + j; }
    uint32_t lo[8], hi[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = t + j; hi[j] = t * 3 + j; c[j] = 0; }
    if (F == WIDE_X) {
        // two rows of four pairs, each row one carry chain (what operator* issues)
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                    : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[u]));
                asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                    : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[u]));
            }
        }
    } else if (F == DFMA) {
        double x[8], y = __longlong_as_double(0x3ff0000000000000ll | a[0]), z = __longlong_as_double(0x3fe0000000000000ll | b[0]);
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = (double)(t + j);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = __fma_rz(x[j], y, z);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) lo[j] ^= (uint32_t)__double_as_longlong(x[j]);
    } else {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int j = 0; j < 8; j++)
                    // flag-free forms: ptxas hoists loop-invariant products, so one multiplicand is another chain's low word
                    op<F>(lo[j], hi[j], c[j], (F == WIDE || F == WIDE_ZC) ? lo[(j + 3) & 7] : a[(u + j) & 3], b[u]);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j] ^ c[j];
    out[t] = s;
}

template <int F, bool VEC>
static void run(const char *name, int sms, uint32_t *out, const uint32_t *in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, iters = 4000;
    k_rate<F, VEC><<<blocks, 256>>>(out, in, 12345, 6789, 10);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k_rate<F, VEC><<<blocks, 256>>>(out, in, 12345, 6789, iters); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = fminf(best, ms);
    }
    double ops = (double)blocks * 256 * iters * 32 ;
    printf("{\"bench\": \"%s\", \"operands\": \"%s\", \"ms\": %.3f, \"Tops_per_s\": %.3f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", name,
           VEC ? "vector" : "uniform", best, ops / best / 1e9, ops / (best * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    uint32_t *in, *out;
    size_t n = (size_t)sms * 8 * 256;
    CK(cudaMalloc(&in, n * 32)); CK(cudaMalloc(&out, n * 4));
    CK(cudaMemset(in, 0x5a, n * 32));
    run<WIDE, false>("imad_wide", sms, out, in);          run<WIDE, true>("imad_wide", sms, out, in);
    run<WIDE_ZC, false>("imad_wide_c32", sms, out, in);   run<WIDE_ZC, true>("imad_wide_c32", sms, out, in);
    run<WIDE_COUT, false>("imad_wide_carry_out", sms, out, in); run<WIDE_COUT, true>("imad_wide_carry_out", sms, out, in);
    run<WIDE_X, false>("imad_wide_x_rows", sms, out, in); run<WIDE_X, true>("imad_wide_x_rows", sms, out, in);
    run<LO, false>("imad_lo", sms, out, in);              run<LO, true>("imad_lo", sms, out, in);
    run<HI, false>("imad_hi", sms, out, in);              run<HI, true>("imad_hi", sms, out, in);
    run<DFMA, true>("dfma_rz", sms, out, in);
    run<IADD, true>("iadd3", sms, out, in);
    return 0;
}
