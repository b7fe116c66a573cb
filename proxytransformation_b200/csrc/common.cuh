// Shared helpers for the sm_100a kernels behind include/pt_preshape.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pt_preshape.h"

namespace pt {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline figures).
enum ProfTag {
    PROF_MINMAX = 0, PROF_CENTRES, PROF_BALL_QUERY, PROF_OFFSET_NET, PROF_DROPOUT, PROF_ENCODER, PROF_LAYERNORM,
    PROF_GEMM_F32, PROF_GEMM_TC, PROF_SPLIT, PROF_ATTENTION, PROF_HEADS, PROF_IMG_MEAN, PROF_IMG_POOL, PROF_MARK,
    PROF_COUNT, PROF_COMPACT, PROF_MISC, PROF_GEMM_IMG, PROF_NTAGS
};
struct ProfScope {
    int slot;
    cudaStream_t stream;
    ProfScope(int tag, cudaStream_t s);
    ~ProfScope();
};

#define PT_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            pt::set_error(__VA_ARGS__);       \
            return PT_ERR_INVALID;            \
        }                                     \
    } while (0)

#define PT_CUDA_OK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            pt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PT_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define PT_LAUNCH_CHECK()                  \
    do {                                   \
        pt::count_launch();                \
        PT_CUDA_OK(cudaGetLastError());    \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(FULL, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// squared distance in the reference's evaluation order ((dx*dx)+(dy*dy))+(dz*dz), no FMA contraction
__device__ __forceinline__ float dist2_rn(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// cudaFuncSetAttribute is per device: remember per (call site, device) whether it has been done (a process normally drives
// one GPU, but nothing here may assume it).  `done` is a call-site static array of PT_MAX_DEVICES flags.
constexpr int PT_MAX_DEVICES = 64;
inline bool first_use_on_current_device(bool* done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PT_MAX_DEVICES) return true;     // unknown: always (re)apply
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pt
