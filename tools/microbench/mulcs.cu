// tools/microbench/mulcs.cu — can a Montgomery product avoid the half-rate carry-IN form IMAD.WIDE.U32.X?
//
// The shipped product (field.cuh, operator*) is operand scanning: every row is one mad.lo.cc / madc.hi.cc chain,
// which ptxas turns into IMAD.WIDE.U32.X (carry in AND out) — measured at 27.9 per SM per clock against 59.3 for
// a plain IMAD.WIDE.U32 (profiles/r1_microbench.jsonl).  Here the product is column scanning ("FIPS"): a column
// accumulator of three words (lo, hi, c); each limb product is  IMAD.WIDE.U32 (lo,hi), P = a*b + (lo,hi)
// (carry OUT only) and the carries of two products are folded into c by one IADD3.X on the ALU pipe.
// Variants measured (all checked bit-for-bit against operator* first):
//   cs1   one accumulator, a*b and m*p products interleaved per column
//   cs2   two accumulators (a*b | m*p) per column, joined at the column end
//   sqr   dedicated squaring, 36 + 64 limb products
//   rate  the bare (IMAD.WIDE P-out, IADD3.X) pattern on independent accumulators
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mulcs mulcs.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../libzkp_b200/csrc/field.cuh"
using namespace lzkp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// (lo, hi) += a*b, no carry out wanted (caller knows it cannot overflow)
__device__ __forceinline__ void mac2(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

template <class P>
__device__ __forceinline__ Fp<P> cs_mul1(const Fp<P> &a, const Fp<P> &b) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, m[8], r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) mac3(t0, t1, t2, a.l[i], b.l[k - i]);
#pragma unroll
        for (int i = 0; i < k; i++) mac3(t0, t1, t2, m[i], P::MOD(k - i));
        m[k] = t0 * P::INV;
        mac3(t0, t1, t2, m[k], P::MOD(0));
        t0 = t1; t1 = t2; t2 = 0;
    }
#pragma unroll
    for (int k = 8; k < 15; k++) {
#pragma unroll
        for (int i = k - 7; i < 8; i++) mac3(t0, t1, t2, a.l[i], b.l[k - i]);
#pragma unroll
        for (int i = k - 7; i < 8; i++) mac3(t0, t1, t2, m[i], P::MOD(k - i));
        r[k - 8] = t0; t0 = t1; t1 = t2; t2 = 0;
    }
    r[7] = t0;
    return Fp<P>::reduce_once(r);
}

template <class P>
__device__ __forceinline__ Fp<P> cs_mul2(const Fp<P> &a, const Fp<P> &b) {
    // u = a*b columns, v = m*p columns; their low words cancel mod 2^32 once m[k]*p[0] is in
    uint32_t u0 = 0, u1 = 0, u2 = 0, v0 = 0, v1 = 0, v2 = 0, m[8], r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; i <= k; i++) mac3(u0, u1, u2, a.l[i], b.l[k - i]);
#pragma unroll
        for (int i = 0; i < k; i++) mac3(v0, v1, v2, m[i], P::MOD(k - i));
        m[k] = (u0 + v0) * P::INV;
        mac3(v0, v1, v2, m[k], P::MOD(0));
        // u0 + v0 == 0 mod 2^32: carry into the next column is 1 unless both are 0
        uint32_t cy = u0 != 0;
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(u1), "+r"(u2) : "r"(cy));
        u0 = u1; u1 = u2; u2 = 0;
        v0 = v1; v1 = v2; v2 = 0;
    }
#pragma unroll
    for (int k = 8; k < 15; k++) {
#pragma unroll
        for (int i = k - 7; i < 8; i++) mac3(u0, u1, u2, a.l[i], b.l[k - i]);
#pragma unroll
        for (int i = k - 7; i < 8; i++) mac3(v0, v1, v2, m[i], P::MOD(k - i));
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
            : "+r"(u0), "+r"(u1), "+r"(u2) : "r"(v0), "r"(v1), "r"(v2));
        r[k - 8] = u0; u0 = u1; u1 = u2; u2 = 0;
        v0 = 0; v1 = 0; v2 = 0;
    }
    r[7] = u0;
    return Fp<P>::reduce_once(r);
}

// a^2: row i multiplies a_i by (a_i, e_{i+1}, d_{i+2}, ..., d_7) where d = 2a (fits 8 limbs, a < 2^255) and
// e_{i+1} = d_{i+1} with bit 0 cleared (that bit is a_i's top bit, which belongs to the a_i * a_i term's row).
template <class P>
__device__ __forceinline__ Fp<P> cs_sqr(const Fp<P> &a) {
    uint32_t d[8], e[8];
    d[0] = a.l[0] << 1;
#pragma unroll
    for (int i = 1; i < 8; i++) d[i] = (a.l[i] << 1) | (a.l[i - 1] >> 31);
#pragma unroll
    for (int i = 1; i < 8; i++) e[i] = a.l[i] << 1;
    uint32_t t0 = 0, t1 = 0, t2 = 0, m[8], r[8];
#pragma unroll
    for (int k = 0; k < 15; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int j = k - i;
            if (j < i || j > 7) continue;
            if (j == i) mac3(t0, t1, t2, a.l[i], a.l[i]);
            else if (j == i + 1) mac3(t0, t1, t2, a.l[i], e[j]);
            else mac3(t0, t1, t2, a.l[i], d[j]);
        }
        if (k < 8) {
#pragma unroll
            for (int i = 0; i < k; i++) mac3(t0, t1, t2, m[i], P::MOD(k - i));
            m[k] = t0 * P::INV;
            mac3(t0, t1, t2, m[k], P::MOD(0));
        } else {
#pragma unroll
            for (int i = k - 7; i < 8; i++) mac3(t0, t1, t2, m[i], P::MOD(k - i));
            r[k - 8] = t0;
        }
        t0 = t1; t1 = t2; t2 = 0;
    }
    r[7] = t0;
    return Fp<P>::reduce_once(r);
}

template <int V>
__device__ __forceinline__ Fq mulv(const Fq &a, const Fq &b) {
    if (V == 0) return a * b;
    if (V == 1) return cs_mul1<FqParams>(a, b);
    if (V == 2) return cs_mul2<FqParams>(a, b);
    if (V == 4) { uint32_t T[16]; mul_wide_k(T, a.l, b.l); return Fq::reduce_wide(T); }      // Karatsuba + reduction
    if (V == 5) { uint32_t T[16]; mul_wide(T, a.l, b.l); return Fq::reduce_wide(T); }        // schoolbook + reduction
    return cs_sqr<FqParams>(a);
}

template <int V, int ILP>
__global__ void __launch_bounds__(128) k_mul(Fq *out, const Fq *in, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x[ILP], y = in[t];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = in[t + j + 1];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = mulv<V>(x[j], y);
    }
    Fq s = x[0];
#pragma unroll
    for (int j = 1; j < ILP; j++) s = s + x[j];
    out[t] = s;
}

__global__ void k_check(const Fq *in, int n, int *bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Fq a = in[t], b = in[t + 1];
    for (int r = 0; r < 8; r++) {
        Fq w = a * b;
        if (cs_mul1<FqParams>(a, b) != w) atomicAdd(bad, 1);
        if (cs_mul2<FqParams>(a, b) != w) atomicAdd(bad + 1, 1);
        if (cs_sqr<FqParams>(a) != a * a) atomicAdd(bad + 2, 1);
        if (mulv<4>(a, b) != w) atomicAdd(bad + 3, 1);
        a = w; b = b + a;
    }
}

__global__ void __launch_bounds__(256) k_rate(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t lo[4], hi[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { lo[j] = threadIdx.x + j; hi[j] = threadIdx.x * 3 + j; c[j] = 0; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < 4; j++) mac3(lo[j], hi[j], c[j], a + u, b + j);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) s ^= lo[j] ^ hi[j] ^ c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Field inversions per second (binary extended GCD, field.cuh): what ONE shared inversion of a batch-affine addition
// scheme costs, in units of Montgomery products.
__global__ void __launch_bounds__(128) k_inv(Fq *out, const Fq *in, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[t], y = in[t + 1];
    for (int i = 0; i < iters; i++) x = x.inverse() + y;
    out[t] = x;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

template <int V, int ILP>
static void run(const char *name, int sms, void *buf, cudaEvent_t e0, cudaEvent_t e1) {
    const int iters = 1000;
    for (int bl : {2, 3, 4, 8, 16}) {     // CTAs of 128 threads per SM: 8 .. 64 warps per SM
        k_mul<V, ILP><<<sms * bl, 128>>>((Fq *)buf, (const Fq *)buf, 10);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k_mul<V, ILP><<<sms * bl, 128>>>((Fq *)buf, (const Fq *)buf, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        double ops = (double)sms * bl * 128 * iters * ILP;
        printf("{\"bench\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"Gmul_per_s\": %.2f}\n", name, ILP, bl * 4, best, ops / best / 1e6);
    }
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void *buf; CK(cudaMalloc(&buf, 64 << 20));
    // operands: reduced field elements (top limb below 0x30000000) from a simple generator
    {
        size_t n = (64 << 20) / 4;
        uint32_t *h = (uint32_t *)malloc(n * 4);
        uint64_t s = 0x9e3779b97f4a7c15ull;
        for (size_t i = 0; i < n; i++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = (uint32_t)(s >> 16);
            if (i % 8 == 7) h[i] &= 0x1fffffffu;
        }
        CK(cudaMemcpy(buf, h, n * 4, cudaMemcpyHostToDevice));
        free(h);
    }
    int *bad; CK(cudaMalloc(&bad, 16)); CK(cudaMemset(bad, 0, 16));
    k_check<<<1024, 256>>>((const Fq *)buf, 1024 * 256, bad);
    int hb[4]; CK(cudaMemcpy(hb, bad, 16, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"vs operator*\", \"cs1_bad\": %d, \"cs2_bad\": %d, \"sqr_bad\": %d, \"karatsuba_bad\": %d}\n", hb[0], hb[1], hb[2], hb[3]);
    {
        const int blocks = sms * 8, iters = 2000;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); k_rate<<<blocks, 256>>>((uint32_t *)buf + (32 << 18), 12345, 6789, iters); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            double ops = (double)blocks * 256 * iters * 16 * 4;
            printf("{\"bench\": \"imad_wide_carry_out_only\", \"rep\": %d, \"ms\": %.3f, \"Tops_per_s\": %.3f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", rep,
                   time_ms(e0, e1), ops / time_ms(e0, e1) / 1e9, ops / (time_ms(e0, e1) * 1e-3) / sms / 1.965e9);
        }
    }
    for (int bl : {4, 16}) {
        k_inv<<<sms * bl, 128>>>((Fq *)buf, (const Fq *)buf, 1);
        cudaEventRecord(e0); k_inv<<<sms * bl, 128>>>((Fq *)buf, (const Fq *)buf, 20); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        double ops = (double)sms * bl * 128 * 20;
        printf("{\"bench\": \"fq_inverse_binary_gcd\", \"warps_per_sm\": %d, \"ms\": %.3f, \"Ginv_per_s\": %.4f}\n", bl * 4, time_ms(e0, e1), ops / time_ms(e0, e1) / 1e6);
    }
    run<0, 1>("cios_shipped", sms, buf, e0, e1);
    run<4, 1>("karatsuba_wide_plus_reduce", sms, buf, e0, e1);
    run<5, 1>("schoolbook_wide_plus_reduce", sms, buf, e0, e1);
    run<4, 2>("karatsuba_wide_plus_reduce", sms, buf, e0, e1);
    run<1, 1>("cs1", sms, buf, e0, e1);
    run<2, 1>("cs2", sms, buf, e0, e1);
    run<3, 1>("cs_sqr", sms, buf, e0, e1);
    run<0, 2>("cios_shipped", sms, buf, e0, e1);
    run<1, 2>("cs1", sms, buf, e0, e1);
    run<2, 2>("cs2", sms, buf, e0, e1);
    return 0;
}
