// pipes.cu — issue rates of the instructions the TV-L1 inner loop is made of, on one SM sub-partition with 4 resident warps
// (the occupancy of k_tvl1_pair: 512 threads, 1 CTA per SM).  Each test runs N independent register chains per thread so that
// dependency latency is hidden; result = cycles per warp-instruction per sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 512
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float *out, long long *cyc, float seed) {
    float a[8], b[8];
    unsigned long long A[8], B[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; A[i] = pk(a[i], b[i]); B[i] = pk(b[i], a[i]); }
    const float c0 = seed * 1.0001f, c1 = seed * 0.999f;
    const unsigned long long C0 = pk(c0, c1);
    __shared__ float4 sm[512];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(b[i]));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(C0), "l"(B[i]));
            if (OP == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(C0));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 7) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 8) { float4 v = sm[(threadIdx.x + i) & 511]; a[i] += v.x + v.y + v.z + v.w; }
            if (OP == 9) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
            if (OP == 10) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
            if (OP == 11) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(C0), "l"(B[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (OP == 12) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(C0), "l"(B[i])); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i])); s += a[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, 512>>>(out, cyc, 1.25f); k<OP><<<148, 512>>>(out, cyc, 1.25f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // 16 warps per SM = 4 per sub-partition; instructions per warp = ITER * 8 * per_iter
    printf("%-26s %7.2f cycles per warp-instruction per sub-partition (4 warps resident)\n", name, avg / (ITER * 8.0 * per_iter * 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA (3 regs)", 1); run<1>("FFMA2", 1); run<2>("FADD", 1); run<3>("FADD2", 1); run<4>("FMUL2", 1); run<10>("FMUL", 1);
    run<5>("FMNMX", 1); run<6>("MUFU.RCP", 1); run<7>("MUFU.SQRT", 1); run<8>("LDS.128 (+4 FADD)", 1); run<9>("SHFL.UP", 1);
    run<11>("FFMA2 + MUFU.RCP pair", 2); run<12>("FFMA2 + FMNMX pair", 2);
    return 0;
}
