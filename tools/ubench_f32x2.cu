// Issue-slot cost of the packed FP32 instructions (FFMA2 / FMUL2) on sm_100a, measured next to scalar FFMA and mixed with ALU / SFU work.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/ubench_f32x2 tools/ubench_f32x2.cu && tools/build/ubench_f32x2
//
// Each warp runs a register-only loop of independent instructions and reports cycles per loop body (clock64 around the loop, slowest warp of
// the SM's resident set).  One line per (mix, warps per scheduler):  slots = warp-instructions per body, cyc = cycles per body per scheduler.
// If FFMA2 held the issue port for one cycle and the FMA pipe for two, "8 FFMA2 + 8 IADD3" would cost 16 cycles; if it holds the issue port
// for two, 24.  Results of the round-2 run are quoted in profiles/history_r02.md section 10.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

#define FFMA2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define FFMA(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define IADD(acc, a) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc) : "r"(a))
#define LOP(acc, a) asm volatile("xor.b32 %0, %0, %1;" : "+r"(acc) : "r"(a))      // a = another accumulator: a constant would cancel over two iterations
#define MUFU(acc) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc))
#define FMUL(acc, a) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(acc) : "f"(a))
#define FMNMX(acc, a) asm volatile("max.f32 %0, %0, %1;" : "+f"(acc) : "f"(a))
#define LDS(acc, addr) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(acc) : "r"(addr))

// MIX: 0 = 16 FFMA, 1 = 8 FFMA2, 2 = 8 FFMA2 + 8 IADD, 3 = 16 FFMA + 8 IADD, 4 = 8 FFMA2 + 4 MUFU, 5 = 16 FFMA + 4 MUFU, 6 = 8 IADD + 8 LOP,
//      7 = 8 FFMA2 + 8 FFMA, 8 = 4 FFMA2 + 12 IADD, 9 = 8 FFMA2 x 8 FMUL alternating, 10 = 8 FFMA2 then 8 FFMA (two runs), 11 = 8 FFMA2 x 8 FMNMX alternating,
//      12 = 8 FFMA2 x 8 LOP alternating, 13 = 8 FFMA2 then 8 LOP (two runs), 14 = 16 FFMA x 8 LOP alternating, 15 = 24 FFMA, 16 = 8 FFMA2 x 8 LDS alternating,
//      17 = 16 FFMA x 8 FMNMX, 18 = runs of 4: (4 FFMA2, 4 FFMA) x 2,
//      19 / 20 / 21 / 22 = 16 FFMA / 8 FFMA2 / 8 (FFMA2, FFMA) alternating / 8 FFMA2 then 8 FFMA with three DISTINCT register operands each (no operand reuse)
template <int MIX> __global__ void __launch_bounds__(1024) body(float seed, unsigned long long *cycles, float *sink) {
    unsigned long long p[8];
    float f[16];
    int k[12];
    const unsigned long long a2 = (static_cast<unsigned long long>(__float_as_uint(seed)) << 32) | __float_as_uint(seed * 0.5f);
    const unsigned long long b2 = (static_cast<unsigned long long>(__float_as_uint(seed * 0.25f)) << 32) | __float_as_uint(seed * 0.125f);
    const int ka = static_cast<int>(seed * 7.0f) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = a2 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = seed + i;
#pragma unroll
    for (int i = 0; i < 12; ++i) k[i] = ka + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < kIters; ++it) {
        if constexpr (MIX == 0 || MIX == 3 || MIX == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                FFMA(f[i], seed, f[(i + 8) & 15]);
                if (MIX == 3 && (i & 1)) LOP(k[i >> 1], k[((i >> 1) + 1) & 7]);
                if (MIX == 5 && (i & 3) == 3) MUFU(f[i]);
            }
        } else if constexpr (MIX == 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                IADD(k[i], ka);
                LOP(k[(i + 4) & 7], k[(i + 5) & 7]);
            }
        } else if constexpr (MIX >= 19) {
            if constexpr (MIX == 19) {
#pragma unroll
                for (int i = 0; i < 16; ++i) FFMA(f[i], f[(i + 3) & 15], f[(i + 7) & 15]);
            } else if constexpr (MIX == 20) {
#pragma unroll
                for (int i = 0; i < 8; ++i) FFMA2(p[i], p[(i + 3) & 7], p[(i + 5) & 7]);
            } else if constexpr (MIX == 21) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    FFMA2(p[i], p[(i + 3) & 7], p[(i + 5) & 7]);
                    FFMA(f[i], f[(i + 3) & 15], f[(i + 7) & 15]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) FFMA2(p[i], p[(i + 3) & 7], p[(i + 5) & 7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) FFMA(f[i], f[(i + 3) & 15], f[(i + 7) & 15]);
            }
        } else if constexpr (MIX >= 9) {
            const unsigned sa = (threadIdx.x & 31) * 4;
            if constexpr (MIX == 9 || MIX == 11 || MIX == 12 || MIX == 16) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    FFMA2(p[i], a2, b2);
                    if (MIX == 9) FMUL(f[i], seed);
                    if (MIX == 11) FMNMX(f[i], f[(i + 1) & 7]);
                    if (MIX == 12) LOP(k[i], k[(i + 1) & 7]);
                    if (MIX == 16) LDS(f[i], sa + 128 * i);
                }
            } else if constexpr (MIX == 10 || MIX == 13) {
#pragma unroll
                for (int i = 0; i < 8; ++i) FFMA2(p[i], a2, b2);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (MIX == 10) FFMA(f[i], seed, f[i + 8]);
                    if (MIX == 13) LOP(k[i], k[(i + 1) & 7]);
                }
            } else if constexpr (MIX == 14 || MIX == 17) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    FFMA(f[i], seed, f[(i + 8) & 15]);
                    if (MIX == 14 && (i & 1)) LOP(k[i >> 1], k[((i >> 1) + 1) & 7]);
                    if (MIX == 17 && (i & 1)) FMNMX(f[i], f[(i + 2) & 15]);
                }
            } else if constexpr (MIX == 15) {
#pragma unroll
                for (int i = 0; i < 24; ++i) FFMA(f[i & 15], seed, f[(i + 8) & 15]);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) FFMA2(p[4 * h + i], a2, b2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) FFMA(f[4 * h + i], seed, f[4 * h + i + 8]);
                }
            }
        } else {
            constexpr int n2 = MIX == 8 ? 4 : 8;
#pragma unroll
            for (int i = 0; i < n2; ++i) {
                FFMA2(p[i], a2, b2);
                if (MIX == 2) LOP(k[i], k[(i + 1) & 7]);
                if (MIX == 4 && (i & 1)) MUFU(f[i]);
                if (MIX == 7) FFMA(f[i], seed, f[i + 8]);
                if (MIX == 8) { LOP(k[3 * i], k[3 * i + 1]); LOP(k[3 * i + 1], k[3 * i + 2]); LOP(k[3 * i + 2], k[(3 * i + 3) % 12]); }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32));
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) s += static_cast<float>(k[i]);
    if (s == 123.456f) sink[0] = s;
    if ((threadIdx.x & 31) == 0) atomicMax(cycles + blockIdx.x, static_cast<unsigned long long>(t1 - t0));
}

template <int MIX> static void run(const char *name, int slots, unsigned long long *d_cyc, float *d_sink, int sms) {
    for (int warps_per_sched : {1, 4}) {
        const int threads = warps_per_sched * 4 * 32;
        cudaMemset(d_cyc, 0, sizeof(unsigned long long) * sms);
        body<MIX><<<sms, threads>>>(1.0009765625f, d_cyc, d_sink);
        cudaMemset(d_cyc, 0, sizeof(unsigned long long) * sms);
        body<MIX><<<sms, threads>>>(1.0009765625f, d_cyc, d_sink);
        unsigned long long h[1024];
        cudaMemcpy(h, d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cyc_per_body = static_cast<double>(mx) / kIters;
        printf("%-28s warps/sched %d  slots/body %2d  cycles/body/warp %7.2f  cycles per body per scheduler %6.2f  (%.2f cycles per instruction)\n", name,
               warps_per_sched, slots, cyc_per_body, cyc_per_body / warps_per_sched, cyc_per_body / warps_per_sched / slots);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *d_cyc;
    float *d_sink;
    cudaMalloc(&d_cyc, sizeof(unsigned long long) * 1024);
    cudaMalloc(&d_sink, 4);
    run<0>("16 FFMA", 16, d_cyc, d_sink, sms);
    run<1>("8 FFMA2", 8, d_cyc, d_sink, sms);
    run<7>("8 FFMA2 + 8 FFMA", 16, d_cyc, d_sink, sms);
    run<2>("8 (FFMA2, LOP3) [= mix 12]", 16, d_cyc, d_sink, sms);
    run<8>("4 (FFMA2, 3 LOP3)", 16, d_cyc, d_sink, sms);
    run<3>("8 (FFMA, FFMA, LOP3)", 24, d_cyc, d_sink, sms);
    run<4>("8 FFMA2 + 4 MUFU", 12, d_cyc, d_sink, sms);
    run<5>("16 FFMA + 4 MUFU", 20, d_cyc, d_sink, sms);
    run<6>("8 IADD + 8 LOP", 16, d_cyc, d_sink, sms);
    run<9>("8 (FFMA2, FMUL)", 16, d_cyc, d_sink, sms);
    run<10>("8 FFMA2 then 8 FFMA", 16, d_cyc, d_sink, sms);
    run<18>("2 x (4 FFMA2, 4 FFMA)", 16, d_cyc, d_sink, sms);
    run<11>("8 (FFMA2, FMNMX)", 16, d_cyc, d_sink, sms);
    run<12>("8 (FFMA2, LOP3)", 16, d_cyc, d_sink, sms);
    run<13>("8 FFMA2 then 8 LOP3", 16, d_cyc, d_sink, sms);
    run<14>("8 (FFMA, FFMA, LOP3)", 24, d_cyc, d_sink, sms);
    run<17>("8 (FFMA, FFMA, FMNMX)", 24, d_cyc, d_sink, sms);
    run<15>("24 FFMA", 24, d_cyc, d_sink, sms);
    run<16>("8 (FFMA2, LDS)", 16, d_cyc, d_sink, sms);
    run<19>("16 FFMA, 3 distinct regs", 16, d_cyc, d_sink, sms);
    run<20>("8 FFMA2, 3 distinct regs", 8, d_cyc, d_sink, sms);
    run<21>("8 (FFMA2, FFMA), distinct", 16, d_cyc, d_sink, sms);
    run<22>("8 FFMA2 then 8 FFMA, distinct", 16, d_cyc, d_sink, sms);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
