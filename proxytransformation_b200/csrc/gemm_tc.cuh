// Internal interface of the tcgen05 3xBF16 GEMM (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace pt {

// C_z[M,N] = act(A[:, z*a_koff_z : +K] W_z[N,K]^T + bias_z) + residual_z   for z in [0, batch)
struct GemmTc {
    int M = 0, N = 0, K = 0, batch = 1;
    // A: bf16 planes [2][a_rows][lda]; the tensor map exposes a_cols columns (reads past a_cols are zero-filled)
    const void* a_split = nullptr;
    int a_rows = 0, a_cols = 0, lda = 0, a_koff_z = 0;
    // W: bf16 planes [2][w_rows][ldw]; batch z uses rows [z*w_row_z, z*w_row_z + N)
    const void* w_split = nullptr;
    int w_rows = 0, ldw = 0, w_row_z = 0;
    const float* bias = nullptr;
    long long bias_off_z = 0;
    const float* residual = nullptr;      // laid out like C
    int act = 0;                          // 0 none, 1 GELU(erf)
    float* C = nullptr;                   // optional fp32 result, element (row, n) of batch z at C + z*c_off_z + row*ldc + n
    int ldc = 0;
    long long c_off_z = 0;
    void* c_split = nullptr;              // optional bf16 hi plane of the result (lo plane cs_plane elements later)
    long long cs_plane = 0;
    int ldcs = 0;
    long long cs_off_z = 0;
    // the split planes as IEEE half instead of bfloat16, of cs_scale * result (cs_scale a power of two that keeps the lo halves out of
    // the half subnormals; the consumer folds 1 / cs_scale).  Used for the w_eff planes of fp16 image features.
    bool cs_fp16 = false;
    float cs_scale = 1.0f;
    // optional TRANSPOSED bf16 hi/lo planes for the output columns n >= ct_col0 (those columns are then written nowhere
    // else): element (row, n) goes to ct_split[(n - ct_col0) * ct_ld + row], lo plane ct_plane elements later.  No bias /
    // activation / residual on these columns.  Used for V^T, the K-major value operand of the tcgen05 attention.
    void* ct_split = nullptr;
    int ct_col0 = 0;
    long long ct_ld = 0, ct_plane = 0;
    // rows come in segments of ct_seg (scenes) that start every ct_seg_pad columns of the transposed planes (0: no segments), so that
    // every segment starts on a 16-byte boundary whatever its length
    int ct_seg = 0, ct_seg_pad = 0;
    int bn = 0;                           // tile width override (32/64/128/256), 0 = auto
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda): 16-bit elements (bf16, or fp16 when
// `fp16`), rank <= 5, dims / box in elements (innermost first), strides in bytes for dims 1..rank-1, SWIZZLE_128B,
// zero fill outside the tensor.
int encode_tensor_map_16bit(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box, bool fp16);

// Profiling tag of the GEMMs launched by the calling thread while the scope lives (the image stage books its projections
// under PROF_GEMM_IMG so that bench.py can report a roofline for the whole stage).
struct GemmProfTagScope {
    int saved;
    explicit GemmProfTagScope(int tag);
    ~GemmProfTagScope();
};

bool gemm_tc_supported(int M, int N, int K);
size_t gemm_tc_ws_bytes(int M, int N, int K);
int launch_gemm_tc_ex(const GemmTc& p, cudaStream_t s);
int launch_gemm_tc(const float* A, const void* w_split, const float* bias, const float* residual, int act, int M, int N,
                   int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s);
int split_rows_bf16(const float* x, long long count, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s);

}  // namespace pt
