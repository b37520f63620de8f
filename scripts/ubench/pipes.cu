// Instruction-throughput microbenchmark for the ops of the node test (sm_100a).  Prints warp-instructions per
// clock per SM for each op at full occupancy.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITER 512
#define NACC 8
template <int OP> __device__ __forceinline__ void step(uint32_t (&r)[NACC], uint32_t k, float fk) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        float f = __uint_as_float(r[i]);
        if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(fk));
        if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(fk));
        if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(fk));
        if (OP == 3) { uint32_t x = r[i]; asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(fk));
        if (OP == 5) { uint32_t x = r[i]; asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 6) { uint32_t x = r[i]; asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 7) { unsigned short h = (unsigned short)r[i]; asm volatile("add.rn.f32.f16 %0, %1, %2;" : "=f"(f) : "h"(h), "f"(fk)); }
        if (OP == 8) { uint32_t x = r[i]; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x)); }
        if (OP == 9) asm volatile("{.reg .pred p; setp.le.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(f) : "f"(fk));
        if (OP == 10) asm volatile("max.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(fk));
        if (OP == 11) { uint32_t x = r[i]; asm volatile("{.reg .b32 t; bfe.u32 t, %1, 8, 8; cvt.rn.f32.u32 %0, t;}" : "=f"(f) : "r"(x)); f = __uint_as_float(__float_as_uint(f) + x); }
        if (OP == 12) { uint32_t x = r[i]; asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.b32 %0, %0, %1, p;}" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 13) { uint32_t x = r[i]; asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 14) { uint32_t x = r[i]; asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(k)); r[i] = x; continue; }
        if (OP == 15) { asm volatile("{.reg .pred p; setp.le.f32 p, %0, %1; @p add.rn.f32 %0, %0, %1;}" : "+f"(f) : "f"(fk)); }
        r[i] = __float_as_uint(f);
    }
}
template <int OP> __device__ __forceinline__ void step2(unsigned long long (&q)[NACC], unsigned long long kk) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        if (OP == 20) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(kk));
        if (OP == 21) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(kk));
        if (OP == 22) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(kk));
    }
}
template <int OP> __global__ void bench(uint32_t* out, long long* cyc, uint32_t k, float fk) {
    uint32_t r[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) r[i] = threadIdx.x * 7 + i + k;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) step<OP>(r, k, fk);
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> __global__ void bench2(uint32_t* out, long long* cyc, uint32_t k, float fk) {
    unsigned long long q[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) q[i] = ((unsigned long long)__float_as_uint(1.0f + i) << 32) | __float_as_uint(2.0f + threadIdx.x);
    unsigned long long kk = ((unsigned long long)__float_as_uint(fk) << 32) | __float_as_uint(fk);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) step2<OP>(q, kk);
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= (uint32_t)q[i] ^ (uint32_t)(q[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// mixed: A then B interleaved, to see whether two ops share a pipe
template <int A, int B> __global__ void mix(uint32_t* out, long long* cyc, uint32_t k, float fk) {
    uint32_t r[NACC], s2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { r[i] = threadIdx.x * 7 + i + k; s2[i] = threadIdx.x * 3 + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) { step<A>(r, k, fk); step<B>(s2, k, fk); }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= r[i] ^ s2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <typename F> void run(const char* name, F launch, int per_iter) {
    const int threads = 1024, blocks = 148;
    uint32_t* out; long long* cyc; cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    launch(blocks, threads, out, cyc); launch(blocks, threads, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warp_inst = (double)(threads / 32) * ITER * per_iter;
    printf("%-28s %8.0f cycles  %6.3f warp-inst/clk/SM  (%5.3f per SMSP)\n", name, avg, warp_inst / avg, warp_inst / avg / 4);
    cudaFree(out); cudaFree(cyc);
}
#define RUN1(OP, NAME) run(NAME, [](int b, int t, uint32_t* o, long long* c) { bench<OP><<<b, t>>>(o, c, 0x4B000000u, 1.0001f); }, NACC)
#define RUN2(OP, NAME) run(NAME, [](int b, int t, uint32_t* o, long long* c) { bench2<OP><<<b, t>>>(o, c, 0x4B000000u, 1.0001f); }, NACC)
#define RUNM(A, B, NAME) run(NAME, [](int b, int t, uint32_t* o, long long* c) { mix<A, B><<<b, t>>>(o, c, 0x4B000000u, 1.0001f); }, 2 * NACC)
int main() {
    RUN1(0, "FADD"); RUN1(1, "FMUL"); RUN1(2, "FFMA"); RUN1(3, "PRMT"); RUN1(4, "FMNMX"); RUN1(10, "FMNMX3"); RUN1(5, "LOP3");
    RUN1(6, "SHF"); RUN1(7, "FHADD (add.f32.f16)"); RUN1(8, "I2FP (cvt.f32.u32)"); RUN1(9, "FSETP+FSEL");
    RUN1(11, "I2F.U8 (+IADD)"); RUN1(12, "SEL"); RUN1(13, "IMAD"); RUN1(14, "IADD3"); RUN1(15, "FSETP + @P FADD");
    RUNM(11, 3, "I2F.U8(+IADD) + PRMT"); RUNM(13, 3, "IMAD + PRMT"); RUNM(12, 3, "SEL + PRMT"); RUNM(4, 2, "FMNMX + FFMA"); RUNM(3, 2, "PRMT + FFMA"); RUNM(7, 2, "FHADD + FFMA");
    RUN2(20, "FADD2"); RUN2(21, "FMUL2"); RUN2(22, "FFMA2");
    RUNM(0, 3, "FADD + PRMT"); RUNM(0, 4, "FADD + FMNMX"); RUNM(3, 4, "PRMT + FMNMX"); RUNM(0, 1, "FADD + FMUL");
    RUNM(7, 3, "FHADD + PRMT"); RUNM(7, 0, "FHADD + FADD"); RUNM(8, 3, "I2FP + PRMT"); RUNM(8, 0, "I2FP + FADD"); RUNM(5, 3, "LOP3 + PRMT");
    return 0;
}
