// ubench.cu -- FP32 issue-rate microbenchmark for the bit-exact (non-FMA) FIR.
// Measures warp-instructions/clk/SM for: FFMA, FMUL+FADD (scalar), and the packed
// f32x2 forms (FFMA2 used as exact mul: a*b+(-0), exact add: a*1+c).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

__global__ void k_ffma(float *out, float a, float b)
{
    float acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = __fmaf_rn(acc[i], a, b);
    float s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mul_add(float *out, float a, float b)
{
    float acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = __fadd_rn(__fmul_rn(acc[i], a), b);
    float s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// exact mul and exact add through two FFMA2 per pair of lanes
__global__ void k_ffma2_muladd(float *out, float a, float b)
{
    unsigned long long acc[NACC];
    const unsigned long long A = pack(a, a), B = pack(b, b), NZ = pack(-0.0f, -0.0f),
                             ONE = pack(1.0f, 1.0f);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            const unsigned long long p = fma2(acc[i], A, NZ);  // exact product
            acc[i] = fma2(p, ONE, B);                          // exact sum
        }
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}

__global__ void k_ffma2(float *out, float a, float b)
{
    unsigned long long acc[NACC];
    const unsigned long long A = pack(a, a), B = pack(b, b);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma2(acc[i], A, B);
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}

__device__ __forceinline__ unsigned long long add2p(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2p(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__global__ void k_fadd2(float *out, float a, float b)
{
    unsigned long long acc[NACC];
    const unsigned long long B = pack(b, b);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = add2p(acc[i], B);
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
__global__ void k_fmul2(float *out, float a, float b)
{
    unsigned long long acc[NACC];
    const unsigned long long A = pack(a, a);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = mul2p(acc[i], A);
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
// products shared by two adds (as in the transposed-form Z phase): can ptxas still contract?
__global__ void k_mul2_add2_shared(float *out, float a, float b)
{
    unsigned long long acc[NACC];
    const unsigned long long A = pack(a, a);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS / 2; it++) {
        unsigned long long p[NACC / 2];
#pragma unroll
        for (int i = 0; i < NACC / 2; i++) p[i] = mul2p(acc[2 * i] , A);
#pragma unroll
        for (int i = 0; i < NACC / 2; i++) {
            acc[2 * i] = add2p(acc[2 * i + 1], p[i]);
            acc[2 * i + 1] = add2p(acc[2 * i], p[(i + 1) % (NACC / 2)]);
        }
    }
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
__global__ void k_ffma2_rt(float *out, float a, float b, float nz, float one)
{
    unsigned long long acc[NACC];
    const unsigned long long A = pack(a, a), B = pack(b, b), NZ = pack(nz, nz), ONE = pack(one, one);
    for (int i = 0; i < NACC; i++) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            const unsigned long long p = fma2(acc[i], A, NZ);
            acc[i] = fma2(p, ONE, B);
        }
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}

template <typename F>
static void run(const char *name, F kern, double lane_ops_per_thread, int sms, double clk_mhz)
{
    float *out;
    const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, threads>>>(out, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) kern<<<blocks, threads>>>(out, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double ops = lane_ops_per_thread * blocks * threads;
    printf("%-18s %8.3f ms  %8.2f T lane-results/s  (%.1f results/clk/SM at %.0f MHz)\n", name, ms,
           ops / ms / 1e9, ops / (ms * 1e-3) / (clk_mhz * 1e6) / sms, clk_mhz);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double mhz = p.clockRate / 1e3;
    printf("device %s, %d SMs, %.0f MHz nominal\n", p.name, p.multiProcessorCount, mhz);
    const double n = (double)ITERS * NACC;
    // "results" = one (mul, add) pair applied to one float
    run("FFMA", k_ffma, n, p.multiProcessorCount, mhz);
    run("FMUL+FADD", k_mul_add, n, p.multiProcessorCount, mhz);
    run("FFMA2 (fused x2)", k_ffma2, 2 * n, p.multiProcessorCount, mhz);
    run("2xFFMA2 literal", k_ffma2_muladd, 2 * n, p.multiProcessorCount, mhz);
    run("FADD2 only", k_fadd2, 2 * n, p.multiProcessorCount, mhz);
    run("FMUL2 only", k_fmul2, 2 * n, p.multiProcessorCount, mhz);
    run("FMUL2+2xFADD2", k_mul2_add2_shared, 2 * n, p.multiProcessorCount, mhz);
    {   // run-time constants: the two FFMA2 cannot be folded
        float *out;
        const int blocks = p.multiProcessorCount * 8, threads = 256;
        cudaMalloc(&out, (size_t)blocks * threads * 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_ffma2_rt<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, -0.0f, 1.0f);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; r++) k_ffma2_rt<<<blocks, threads>>>(out, 1.0000001f, 1e-9f, -0.0f, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        const double ops = 2 * n * blocks * threads;
        printf("%-18s %8.3f ms  %8.2f T lane-results/s  (%.1f results/clk/SM at %.0f MHz)\n",
               "2xFFMA2 runtime", ms, ops / ms / 1e9, ops / (ms * 1e-3) / (mhz * 1e6) / p.multiProcessorCount, mhz);
    }
    return 0;
}
