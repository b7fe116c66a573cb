// Probe: issue rate of legacy mma.sync.m16n8k16 bf16 (fp32 accumulate) on sm_100a, and of ldmatrix.x4 next to it.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/hmma_rate_probe tools/probes/hmma_rate_probe.cu && /tmp/hmma_rate_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int CHAINS, bool LDSM>
__global__ void probe(int iters, long long* cycles, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;
    __syncthreads();
    float acc[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    uint32_t b[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 2048;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; c += 2) {
            if (LDSM) ldsm_x4(b, base + ((it * CHAINS + c) & 3) * 512);
            mma(acc[c], a, b[0], b[1]);
            mma(acc[c + 1], a, b[2], b[3]);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c) for (int e = 0; e < 4; ++e) s += acc[c][e];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS, bool LDSM>
void run(int warps) {
    long long* d; float* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    cudaFuncSetAttribute(probe<CHAINS, LDSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    probe<CHAINS, LDSM><<<148, warps * 32, 48 * 1024>>>(iters, d, sink);
    probe<CHAINS, LDSM><<<148, warps * 32, 48 * 1024>>>(iters, d, sink);
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)iters * CHAINS * warps / 4.0;
    printf("chains %2d ldmatrix %d warps/SM %2d : %8.2f cycles per HMMA per SMSP  (%.0f MAC/clk/SM)  err=%s\n", CHAINS, (int)LDSM, warps,
           c / per_smsp, 2048.0 * 4.0 * per_smsp / c, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(sink);
}

int main() {
    run<2, false>(4); run<4, false>(4); run<8, false>(4); run<16, false>(4);
    run<2, false>(16); run<4, false>(16); run<8, false>(16); run<16, false>(16);
    run<8, true>(4); run<8, true>(16); run<16, true>(16); run<4, true>(16);
    return 0;
}
