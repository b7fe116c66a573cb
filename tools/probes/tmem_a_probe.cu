// Probe: tcgen05.mma with the A operand in TMEM (written by tcgen05.st 32x32b) and B K-major SWIZZLE_128B in smem.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_a_probe tools/probes/tmem_a_probe.cu && /tmp/tmem_a_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr int M = 128, N = 16, K = 32;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    __shared__ __align__(1024) uint8_t sB[2048];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2048 / 4; i += 128) ((uint32_t*)sB)[i] = 0;
    __syncthreads();
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const int off = (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
        *(__nv_bfloat16*)(sB + off) = B[n * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    // A row of this thread (lane = row) packed two bf16 per column, low half = even k
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) {
        const uint32_t lo = __bfloat16_as_ushort(A[tid * K + 2 * j]), hi = __bfloat16_as_ushort(A[tid * K + 2 * j + 1]);
        r[j] = lo | (hi << 16);
    }
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // smem B written by generic stores -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t db = desc_sw128(smem_u32(sB)) + (uint64_t)((ks * 32) >> 4);
            const uint32_t a_addr = tb + 8 * ks, d_addr = tb + 32;
            uint32_t acc = ks > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_addr), "r"(a_addr), "l"(db), "r"(IDESC), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t daddr = tb + ((uint32_t)(warp * 32) << 16) + 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(daddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * N + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}

int main() {
    static __nv_bfloat16 hA[M * K], hB[N * K];
    static float hD[M * N], ref[M * N];
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hA[m * K + k] = __float2bfloat16((float)((m * 7 + k * 3) % 11 - 5) * 0.25f);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16((float)((n * 5 + k) % 9 - 4) * 0.5f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int k = 0; k < K; ++k) s += __bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
        ref[m * N + n] = s;
    }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs(hD[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("TMEM-A probe: max |err| = %g, mismatches = %d / %d  (D[0][0..3] = %g %g %g %g ; ref %g %g %g %g)\n", maxerr, bad, M * N,
           hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    return bad != 0;
}
