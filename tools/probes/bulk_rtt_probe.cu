// Probe: round-trip time of "consumer releases a slot -> producer thread sees it -> cp.async.bulk global->smem -> consumer sees
// the full barrier", per transfer size, L2-miss (fresh addresses) vs L2-hit (same address again), 148 CTAs at once.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/rtt tools/probes/bulk_rtt_probe.cu && /tmp/rtt
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(dst)), "l"(src), "r"(n), "r"(su32(b)) : "memory");
}

// mode 0: consumer-released slot -> producer warp -> load -> consumer (full handshake)
// mode 1: the consumer thread issues the load itself (pure TMA latency)
__global__ void probe(const uint8_t* src, size_t stride, uint32_t bytes, int iters, int mode, int hit, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full, empty;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&empty, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint8_t* base = src + (size_t)blockIdx.x * stride;
    if (warp == 1) {
        if (lane == 0 && mode == 0)
            for (int i = 0; i < iters; ++i) {
                mbar_wait(&empty, i & 1);
                mbar_expect(&full, bytes);
                bulk(smem, base + (hit ? 0 : (size_t)i * 32768), bytes, &full);
            }
        return;
    }
    long long acc = 0;
    for (int i = 0; i < iters; ++i) {
        __syncwarp();
        const long long t0 = clock64();
        if (lane == 0) {
            if (mode == 0) mbar_arrive(&empty);
            else { mbar_expect(&full, bytes); bulk(smem, base + (hit ? 0 : (size_t)i * 32768), bytes, &full); }
        }
        mbar_wait(&full, i & 1);
        const long long t1 = clock64();
        if (i >= 2) acc += t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc / (iters - 2);
}

int main() {
    const size_t stride = 4u << 20;
    uint8_t* src; long long* out;
    cudaMalloc(&src, stride * 148);
    cudaMemset(src, 1, stride * 148);
    cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const uint32_t sizes[] = {16, 1024, 4096, 14400, 28800};
    for (int mode = 0; mode < 2; ++mode)
        for (int hit = 0; hit < 2; ++hit)
            for (uint32_t b : sizes)
                for (int grid : {1, 148}) {
                    probe<<<grid, 64, 64 * 1024>>>(src, stride, b, 66, mode, hit, out);
                    long long h[148];
                    cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
                    long long s = 0; for (int i = 0; i < grid; ++i) s += h[i];
                    printf("mode %d (%s) %s bytes %6u grid %3d : %7.0f cycles  %s\n", mode, mode ? "self-issued" : "handshake", hit ? "L2-hit " : "L2-miss", b, grid,
                           (double)s / grid, cudaGetErrorString(cudaGetLastError()));
                }
    return 0;
}
