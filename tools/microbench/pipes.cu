// pipes.cu -- issue-rate microbenchmark for the integer instruction classes the rules kernels are made of (sm_100a).
// Each mode runs ITER iterations of 8 independent dependency chains per thread; prints warp-instructions per cycle per
// SM sub-partition for a full SM (16 warps per scheduler) and for 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Opaque { uint32_t one, c1, c2, sh; };

template <int MODE>
__global__ void __launch_bounds__(512) k(Opaque o, int iters, uint32_t* out, long long* cycles)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 8 + i + o.c1;
    __shared__ uint32_t tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = threadIdx.x * 4u % 128u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                if (MODE == 2) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                                 else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2)); }
                if (MODE == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(o.c1));
                if (MODE == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.c1), "r"(o.sh));
                if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(o.c1));
                if (MODE == 6) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                if (MODE == 7) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                                 else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(o.c1)); }
                if (MODE == 8) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                if (MODE == 9) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                                 else { x[i] = tab[x[i] & 31u]; } }
                if (MODE == 10) { if ((i & 3) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
                                  else asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2)); }
                if (MODE == 11) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&x[i & ~1])) : "r"(o.c1), "r"(o.c2));
                if (MODE == 12) asm volatile("shl.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(o.sh));
                if (MODE == 13) asm volatile("bfe.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.sh), "r"(o.c2));
                if (MODE == 14) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(o.c1));
                if (MODE == 15) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(o.c1), "r"(o.c2));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads)
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 2048 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
    Opaque o{1u, 0x9E3779B9u, 0x7F4A7C15u, 7u};
    const int iters = 2000;
    k<MODE><<<148, threads>>>(o, iters, out, cyc);
    k<MODE><<<148, threads>>>(o, iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double winstr = (double)iters * 32 * (threads / 32);       // warp instructions per SM
    printf("%-28s %4d thr/SM (%2d warps/sched): %.3f warp-instr / cycle / SMSP  (err %s)\n", name, threads, threads / 128, winstr / avg / 4.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

#define RUN(M, N) run<M>(N, 512); run<M>(N, 128);
int main()
{
    RUN(0, "LOP3"); RUN(1, "IMAD"); RUN(2, "LOP3+IMAD 1:1"); RUN(10, "LOP3+IMAD 3:1"); RUN(3, "IMAD.HI"); RUN(7, "LOP3+IMAD.HI 1:1");
    RUN(11, "IMAD.WIDE"); RUN(4, "SHF.R.W"); RUN(12, "SHL"); RUN(13, "BFE(2 instr?)"); RUN(5, "IADD"); RUN(6, "ISETP+SEL (2 instr)");
    RUN(14, "IMNMX"); RUN(15, "PRMT"); RUN(8, "POPC"); RUN(9, "LOP3+LDS 1:1");
    return 0;
}
