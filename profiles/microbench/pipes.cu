// Issue-rate microbenchmark for the CUDA-core instructions the fused edge kernel is made of (sm_100a).
// Each test: every warp runs ITER iterations of 8 independent dependency chains of one instruction (or a short
// mix); prints warp-instructions per clock per SM sub-partition (SMSP), from clock64() deltas of one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITER = 2048;

#define CHAIN8(body)  body(0) body(1) body(2) body(3) body(4) body(5) body(6) body(7)

template <int OP>
__global__ void __launch_bounds__(1024) bench(float *out, long long *cycles, float seed) {
    float r[16];
    uint32_t u[16];
    uint64_t d[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = seed + i + threadIdx.x; u[i] = (uint32_t)(threadIdx.x * 17 + i); }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d[i]) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
    const float c1 = seed * 0.5f, c2 = seed * 0.25f;
    uint64_t cc;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(c1), "f"(c2));
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (OP == 0) {
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(c1), "f"(c2));
            CHAIN8(B)
#undef B
        } else if (OP == 1) {
#define B(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(cc));
            CHAIN8(B)
#undef B
        } else if (OP == 2) {
#define B(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c1));
            CHAIN8(B)
#undef B
        } else if (OP == 3) {
#define B(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(cc));
            CHAIN8(B)
#undef B
        } else if (OP == 4) {
#define B(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[8]), "r"(u[9]));
            CHAIN8(B)
#undef B
        } else if (OP == 5) {
#define B(i) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[8 + (i & 3)]));
            CHAIN8(B)
#undef B
        } else if (OP == 6) {   // cvt.rn.bf16x2.f32 (F2FP)
#define B(i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(r[i]), "f"(r[8 + i])); r[i] = __uint_as_float(u[i]);
            CHAIN8(B)
#undef B
        } else if (OP == 7) {   // MUFU.EX2
#define B(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
            CHAIN8(B)
#undef B
        } else if (OP == 8) {   // ELU as written in the kernel: x > 0 ? x : ex2(x*log2e) - 1
#define B(i) { float x = r[i], t; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 1.4426950408889634f)); r[i] = (x > 0.f ? x : t - 1.0f) + c1; }
            CHAIN8(B)
#undef B
        } else if (OP == 9) {   // SHFL.BFLY
#define B(i) r[i] = __shfl_xor_sync(0xffffffffu, r[i], 1 + (i & 15));
            CHAIN8(B)
#undef B
        } else if (OP == 10) {  // FFMA + LOP3 interleaved (two pipes)
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(c1), "f"(c2)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[8]), "r"(u[9]));
            CHAIN8(B)
#undef B
        } else if (OP == 11) {  // FFMA2 + LOP3 interleaved
#define B(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(cc)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[8]), "r"(u[9]));
            CHAIN8(B)
#undef B
        } else if (OP == 12) {  // split_bf16x2 as written in the kernel (per PAIR of elements)
#define B(i) { uint32_t hi, lo; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(r[8 + i]), "f"(r[i])); \
               float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u); \
               asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r[8 + i] - hb), "f"(r[i] - ha)); \
               r[i] += __uint_as_float(lo); r[8 + i] += __uint_as_float(hi); }
            CHAIN8(B)
#undef B
        } else if (OP == 13) {  // FMUL
#define B(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(c1));
            CHAIN8(B)
#undef B
        } else if (OP == 14) {  // FSETP + FSEL (select)
#define B(i) r[i] = r[i] > c1 ? r[8 + i] : r[i] + 0.f;
            CHAIN8(B)
#undef B
        } else if (OP == 15) {  // FFMA + FFMA2 interleaved (same pipe?)
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(c1), "f"(c2)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(cc));
            CHAIN8(B)
#undef B
        } else if (OP == 16) {  // FFMA + MUFU interleaved 4:1
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(c1), "f"(c2));
            CHAIN8(B)
#undef B
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[8]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[9]));
        } else if (OP == 17) {  // fmnmx
#define B(i) r[i] = fmaxf(r[i], r[8 + (i & 7)]);
            CHAIN8(B)
#undef B
        } else if (OP == 18) {  // packed bf16 fma (HFMA2.BF16)
#define B(i) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[8]), "r"(u[9]));
            CHAIN8(B)
#undef B
        } else if (OP == 19) {  // shift
#define B(i) asm volatile("shl.b32 %0, %0, 16;" : "+r"(u[i]));
            CHAIN8(B)
#undef B
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += r[i] + __uint_as_float(u[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(d[i])); acc += a + b; }
    if (acc == 123.456f) out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int instr_per_iter, float *out, long long *cyc) {
    for (int threads : {256, 512, 1024}) {
        bench<OP><<<148, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < 148; ++i) mean += (double)h[i];
        mean /= 148;
        const double warps_per_smsp = threads / 32 / 4.0;
        const double ipc = (double)ITER * instr_per_iter * warps_per_smsp / mean;
        printf("%-44s warps/SMSP=%4.1f  %.3f warp-instr/clk/SMSP  (%.2f clk per warp-instr)\n", name, warps_per_smsp, ipc, 1.0 / ipc);
    }
}

int main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA", 8, out, cyc);
    run<1>("FFMA2 (fma.rn.f32x2)", 8, out, cyc);
    run<2>("FADD", 8, out, cyc);
    run<3>("FADD2", 8, out, cyc);
    run<13>("FMUL", 8, out, cyc);
    run<4>("LOP3", 8, out, cyc);
    run<19>("SHL", 8, out, cyc);
    run<5>("PRMT", 8, out, cyc);
    run<17>("FMNMX", 8, out, cyc);
    run<6>("F2FP.BF16.PACK_AB (cvt.rn.bf16x2.f32)", 8, out, cyc);
    run<7>("MUFU.EX2", 8, out, cyc);
    run<9>("SHFL.BFLY", 8, out, cyc);
    run<18>("HFMA2.BF16", 8, out, cyc);
    run<14>("FSETP+FSEL(+FADD) select [count as 8 selects]", 8, out, cyc);
    run<8>("ELU sequence [per element]", 8, out, cyc);
    run<12>("split_bf16x2 [per PAIR]", 8, out, cyc);
    run<10>("FFMA + LOP3 interleaved", 16, out, cyc);
    run<11>("FFMA2 + LOP3 interleaved", 16, out, cyc);
    run<15>("FFMA + FFMA2 interleaved", 16, out, cyc);
    run<16>("8 FFMA + 2 MUFU", 10, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
