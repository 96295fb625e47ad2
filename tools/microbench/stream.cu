// stream.cu -- what the memory system gives for the access pattern of azb_step (4.2 M two-player games, structure of arrays):
// read 17 state words + 1 action byte, write 17 state words + 6 mask words + done + status per game.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream stream.cu && ./stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int W = 17;

template <int MODE>
__global__ void __launch_bounds__(256) k_plain(const uint32_t* __restrict__ s_in, uint32_t* __restrict__ s, const uint8_t* __restrict__ a, uint32_t* __restrict__ m, uint8_t* __restrict__ done, uint8_t* __restrict__ st, int64_t n)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t w[W];
#pragma unroll
    for (int i = 0; i < W; i++) w[i] = MODE == 1 ? __ldcs(s_in + i * n + g) : s_in[i * n + g];
    const uint32_t act = a[g];
    uint32_t x = act;
#pragma unroll
    for (int i = 0; i < W; i++) { x = x * 0x9E3779B9u + w[i]; w[i] ^= x >> 7; }
#pragma unroll
    for (int i = 0; i < W; i++) { if (MODE == 1) __stcs(s + i * n + g, w[i]); else s[i * n + g] = w[i]; }
#pragma unroll
    for (int i = 0; i < 6; i++) { if (MODE == 1) __stcs(m + i * n + g, x + i); else m[i * n + g] = x + i; }
    done[g] = x & 1; st[g] = (x >> 1) & 1;
}

int main()
{
    const int64_t n = 4194304;
    uint32_t *s, *m; uint8_t *a, *d, *st;
    cudaMalloc(&s, W * n * 4); cudaMalloc(&m, 6 * n * 4); cudaMalloc(&a, n); cudaMalloc(&d, n); cudaMalloc(&st, n);
    cudaMemset(s, 1, W * n * 4); cudaMemset(a, 3, n);
    uint32_t* flush; cudaMalloc(&flush, 256 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 12; it++) {
            cudaMemsetAsync(flush, it, 256 << 20);
            cudaEventRecord(e0);
            if (mode == 0) k_plain<0><<<(unsigned)((n + 255) / 256), 256>>>(s, s, a, m, d, st, n);
            else k_plain<1><<<(unsigned)((n + 255) / 256), 256>>>(s, s, a, m, d, st, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("mode %d (%s): avg %.1f us, best %.1f us -> %.0f GB/s algorithmic (163 B/game), err %s\n", mode, mode ? "ld.cs/st.cs" : "plain",
               1e3f * sum / 10, 1e3f * best, 163.0 * n / (sum / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
