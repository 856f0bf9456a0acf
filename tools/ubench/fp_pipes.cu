// Issue rate of the un-fused fp32 instructions the exact path is made of (FMUL, FADD, FFMA; register and immediate
// forms) per SM sub-partition on sm_100a.  One block per SM, W warps per scheduler, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_pipes fp_pipes.cu && ./fp_pipes
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void k(float *out, float a, float b, long long *cyc)
{
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a + threadIdx.x + i;
    unsigned ri[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ri[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(a));
            if (OP == 3) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(r[i]));
            if (OP == 4) asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+f"(r[i]));
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(r[i]) : "f"(a));
            if (OP == 6) { if (i & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); }
            if (OP == 7) { if (i & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); else asm volatile("add.u32 %0, %0, %1;" : "+r"(ri[i]) : "r"(ri[(i + 1) & 7])); }
            if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(ri[i]) : "r"(ri[(i + 2) & 7]));
            if (OP == 9) { if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); else asm volatile("add.u32 %0, %0, %1;" : "+r"(ri[i]) : "r"(ri[(i + 1) & 7])); }
            if (OP == 10) { if (i & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); else asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+f"(r[i])); }
            if (OP == 11) { if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(a)); else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b)); }
        }
    }
    const long long t1 = clock64();
    float s = 0; unsigned si = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += r[i]; si += ri[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, float *out, long long *cyc)
{
    printf("%-28s", name);
    for (int w = 1; w <= 8; w *= 2) {
        k<OP><<<148, 128 * w>>>(out, 1.5f, 1.0000001f, cyc);
        cudaDeviceSynchronize();
        k<OP><<<148, 128 * w>>>(out, 1.5f, 1.0000001f, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  W=%d: %.2f cyc/instr/SMSP", w, (double)c / ((double)ITERS * 8 * w));
    }
    printf("\n");
}
int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("FMUL r,r", out, cyc); run<1>("FADD r,r", out, cyc); run<2>("FFMA r,r,r", out, cyc);
    run<3>("FMUL r,imm", out, cyc); run<4>("FADD r,imm", out, cyc); run<5>("FFMA r,imm,r", out, cyc);
    run<6>("FMUL + FADD", out, cyc); run<7>("FMUL + IADD", out, cyc); run<8>("IADD", out, cyc); run<9>("FADD + IADD", out, cyc);
    run<10>("FMUL r,r + FADD r,imm", out, cyc); run<11>("FFMA + FADD", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
