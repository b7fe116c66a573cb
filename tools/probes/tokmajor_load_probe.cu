// Probe for the token-major single-pass image-pool design: can a persistent CTA stream a view
// ([512 channels][225 tokens] bf16, 450-byte row pitch) as WINDOWS of PIECE bytes per channel row — 16-byte cp.async copies
// of the aligned chunks, 512 rows per window — at HBM speed?  Compared with the slab order the shipped kernel uses
// (contiguous 28.8 KB cp.async.bulk copies).  Nothing is computed: each window is only waited for and released.
//
//   slab order:   cp.async.bulk 28 800 B (8 per view), ring of `stages` slots
//   token-major:  windows of 32 / 64 / 128 B per row (15 / 8 / 4 windows per view), cp.async.cg 16 B, STAGES windows in flight
// Rows start at byte 450 c; a window piece starts at the 16-byte aligned address at or below 450 c + PIECE * w (what the
// residue-class trick of imgpool_tc.cu reads), clamped to the view.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tokmajor_load_probe.bin tools/probes/tokmajor_load_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int C = 512, HW = 225, VIEW_BYTES = C * HW * 2, SLAB_BYTES = 64 * HW * 2, THREADS = 512;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(su32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(dst)), "l"(src), "r"(n), "r"(su32(b)) : "memory");
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(dst)), "l"(src) : "memory");
}

// mode 0: one producer thread, consumers only wait on the full barriers (slots are released by a block barrier)
__global__ void __launch_bounds__(THREADS, 1) slab_kernel(const uint8_t* img, int views, int stages, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) mbar_init(full + s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int nv = blockIdx.x < views ? (views - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total = nv * 8;
    unsigned long long acc = 0;
    // prologue: fill the ring
    if (threadIdx.x == 0)
        for (int i = 0; i < stages && i < total; ++i) {
            const uint8_t* src = img + (size_t)(blockIdx.x + (i / 8) * gridDim.x) * VIEW_BYTES + (size_t)(i % 8) * SLAB_BYTES;
            mbar_expect(full + i, SLAB_BYTES);
            bulk(smem + (size_t)i * SLAB_BYTES, src, SLAB_BYTES, full + i);
        }
    for (int i = 0; i < total; ++i) {
        const int s = i % stages;
        mbar_wait(full + s, (i / stages) & 1);
        acc += smem[(size_t)s * SLAB_BYTES + threadIdx.x * 16];
        __syncthreads();                                   // everybody is done with the slot
        const int j = i + stages;
        if (threadIdx.x == 0 && j < total) {
            const uint8_t* src = img + (size_t)(blockIdx.x + (j / 8) * gridDim.x) * VIEW_BYTES + (size_t)(j % 8) * SLAB_BYTES;
            mbar_expect(full + s, SLAB_BYTES);
            bulk(smem + (size_t)s * SLAB_BYTES, src, SLAB_BYTES, full + s);
        }
    }
    if (acc == 0x12345678ull) sink[0] = acc;
}

// token-major windows: every thread copies its share of each window with 16-byte cp.async, STAGES commit groups in flight
// (cp.async.wait_group needs an immediate, hence the template on the stage count)
template <int PIECE, int STAGES>
__global__ void __launch_bounds__(THREADS, 1) window_kernel_s(const uint8_t* img, int views, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int WINDOWS = (480 + PIECE - 1) / PIECE;
    constexpr int CHUNKS = PIECE / 16;
    constexpr int WIN_BYTES = C * PIECE;
    constexpr int PER_THREAD = C * CHUNKS / THREADS;
    const int nv = blockIdx.x < views ? (views - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total = nv * WINDOWS;
    auto issue = [&](int i) {
        if (i < total) {
            const uint8_t* view = img + (size_t)(blockIdx.x + (i / WINDOWS) * gridDim.x) * VIEW_BYTES;
            const int w = i % WINDOWS;
            uint8_t* dst = smem + (size_t)(i % STAGES) * WIN_BYTES;
#pragma unroll
            for (int t = 0; t < PER_THREAD; ++t) {
                const int e = threadIdx.x + t * THREADS;
                const int row = e / CHUNKS, ch = e % CHUNKS;
                long long off = (((long long)row * 450) & ~15ll) + (long long)w * PIECE + ch * 16;
                if (off > VIEW_BYTES - 16) off = VIEW_BYTES - 16;
                cp16(dst + (size_t)row * PIECE + ((ch ^ (row & (CHUNKS - 1))) * 16), view + off);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int i = 0; i < STAGES - 1; ++i) issue(i);
    unsigned long long acc = 0;
    for (int i = 0; i < total; ++i) {
        issue(i + STAGES - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");    // window i has landed (for this thread's copies)
        __syncthreads();                                                          // ... and for everybody's
        acc += smem[(size_t)(i % STAGES) * WIN_BYTES + threadIdx.x * 16];
        __syncthreads();                                                          // slot free for window i + STAGES
    }
    if (acc == 0x12345678ull) sink[0] = acc;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int views = 64 * 196;                               // the bench batch: 2.89 GB >> L2
    const size_t bytes = (size_t)views * VIEW_BYTES;
    uint8_t* img; unsigned long long* sink;
    if (cudaMalloc(&img, bytes + 4096) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(img, 1, bytes + 4096);
    cudaMalloc(&sink, 8);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double gb = (double)bytes / 1e9;
    for (int stages = 4; stages <= 7; ++stages) {
        const size_t smem = (size_t)stages * SLAB_BYTES;
        cudaFuncSetAttribute(slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const float ms = time_ms([&] { slab_kernel<<<sms, THREADS, smem>>>(img, views, stages, sink); }, 5);
        printf("slab order   (cp.async.bulk 28800 B) stages=%d (%3zu KB in flight): %.3f ms  %.0f GB/s   %s\n", stages, smem / 1024, ms, gb / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    }
#define RUN_WIN(PIECE, STAGES)                                                                                                     \
    {                                                                                                                              \
        const size_t smem = (size_t)(STAGES) * C * (PIECE);                                                                        \
        cudaFuncSetAttribute(window_kernel_s<PIECE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
        const float ms = time_ms([&] { window_kernel_s<PIECE, STAGES><<<sms, THREADS, smem>>>(img, views, sink); }, 5);           \
        const int windows = (480 + (PIECE) - 1) / (PIECE);                                                                         \
        printf("token-major  (cp.async 16 B, %3d B per row, %d windows/view) stages=%d (%3zu KB in flight): %.3f ms  %.0f GB/s algorithmic (%.0f GB/s moved)   %s\n", \
               PIECE, windows, STAGES, smem / 1024, ms, gb / (ms * 1e-3), gb * (windows * (PIECE) / 450.0) / (ms * 1e-3), cudaGetErrorString(cudaGetLastError())); \
    }
    RUN_WIN(64, 3) RUN_WIN(64, 4) RUN_WIN(64, 5) RUN_WIN(64, 6)
    RUN_WIN(128, 2) RUN_WIN(128, 3)
    RUN_WIN(32, 6) RUN_WIN(32, 10)
    return 0;
}
