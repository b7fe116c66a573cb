// C-ABI glue: error reporting, launch accounting and the multi-kernel ProxyBlock stage (S7).
#include "common.cuh"

#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace pt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s);
int launch_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, int act, int M, int N, int K,
                    float* C, cudaStream_t s);
int launch_proxy_attention(const float* qkv, const float* pt_tok, const uint8_t* mask, int B, int n, int l, int c,
                           int heads, float* o, cudaStream_t s);
// tensor-core path (gemm_tc.cu)
bool gemm_tc_supported(int M, int N, int K);
size_t gemm_tc_ws_bytes(int M, int N, int K);
int launch_gemm_tc(const float* A, const void* w_split, const float* bias, const float* residual, int act, int M, int N,
                   int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s);

static int gemm_any(const float* A, const float* W, const void* w_split, const float* bias, const float* residual, int act,
                    int M, int N, int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s) {
    if (w_split != nullptr && gemm_tc_supported(M, N, K)) return launch_gemm_tc(A, w_split, bias, residual, act, M, N, K, C, ws, ws_bytes, s);
    return launch_gemm_f32(A, W, bias, residual, act, M, N, K, C, s);
}

struct BlockWs {
    float *u, *qkv, *pt, *o, *x1, *h2, *hid, *x2;
    void* gemm;
    size_t gemm_bytes, total;
};

static BlockWs carve_block(void* ws, int B, int n, int l, int c, int hidden) {
    BlockWs r;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = ws ? (void*)((char*)ws + off) : nullptr; off += align_up(bytes, 256); return p; };
    const size_t rows = (size_t)B * n;
    r.u = (float*)take(rows * c * 4);
    r.qkv = (float*)take(rows * 3 * c * 4);
    r.pt = (float*)take((size_t)B * l * c * 4);
    r.o = (float*)take(rows * c * 4);
    r.x1 = (float*)take(rows * c * 4);
    r.h2 = (float*)take(rows * c * 4);
    r.hid = (float*)take(rows * hidden * 4);
    r.x2 = (float*)take(rows * c * 4);
    size_t g = gemm_tc_ws_bytes((int)rows, 3 * c, c);
    size_t g2 = gemm_tc_ws_bytes((int)rows, c, hidden);
    size_t g3 = gemm_tc_ws_bytes(B * l, c, c);
    r.gemm_bytes = g > g2 ? (g > g3 ? g : g3) : (g2 > g3 ? g2 : g3);
    r.gemm = take(r.gemm_bytes);
    r.total = off;
    return r;
}

}  // namespace pt

using namespace pt;

extern "C" int pt_abi_version(void) { return 1; }
extern "C" const char* pt_last_error_string(void) { return g_err; }
extern "C" int64_t pt_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" size_t pt_proxy_block_ws_bytes(int B, int n, int l, int c, int hidden) {
    return carve_block(nullptr, B, n, l, c, hidden).total;
}

extern "C" int pt_proxy_block_fused(const float* x, const float* proxy, const uint8_t* mask, const pt_proxy_block_params* p,
                                    int B, int n, int l, int c, int heads, int hidden, float* out, void* ws, size_t ws_bytes,
                                    pt_stream_t stream) {
    PT_REQUIRE(x && proxy && p && out && ws, "pt_proxy_block_fused: null pointer");
    PT_REQUIRE(B > 0 && n > 0 && l > 0 && c > 0 && heads > 0 && hidden > 0 && c % heads == 0, "pt_proxy_block_fused: bad shape");
    BlockWs w = carve_block(ws, B, n, l, c, hidden);
    if (ws_bytes < w.total) { set_error("pt_proxy_block_fused: workspace %zu < %zu", ws_bytes, w.total); return PT_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int rows = B * n;
    int rc;
    // u = LN1(x) + bias[m]                                           (:274, :212-217)
    if ((rc = launch_layernorm(x, p->ln1_w, p->ln1_b, p->pos_bias, n, rows, c, w.u, s))) return rc;
    // [Q|K|V] = u Wqkv^T                                             (:221)
    if ((rc = gemm_any(w.u, p->qkv_w, p->qkv_w_split, nullptr, nullptr, 0, rows, 3 * c, c, w.qkv, w.gemm, w.gemm_bytes, s))) return rc;
    // Pt = proxy Wp^T + bp                                           (:223)
    if ((rc = gemm_any(proxy, p->pp_w, p->pp_w_split, p->pp_b, nullptr, 0, B * l, c, c, w.pt, w.gemm, w.gemm_bytes, s))) return rc;
    // two-stage proxy attention                                      (:225-252)
    if ((rc = launch_proxy_attention(w.qkv, w.pt, mask, B, n, l, c, heads, w.o, s))) return rc;
    // x1 = x + (o Wo^T + bo)                                         (:255, :274)
    if ((rc = gemm_any(w.o, p->proj_w, p->proj_w_split, p->proj_b, x, 0, rows, c, c, w.x1, w.gemm, w.gemm_bytes, s))) return rc;
    // x2 = x1 + fc2(GELU(fc1(LN2(x1))))                              (:275)
    if ((rc = launch_layernorm(w.x1, p->ln2_w, p->ln2_b, nullptr, 1, rows, c, w.h2, s))) return rc;
    if ((rc = gemm_any(w.h2, p->fc1_w, p->fc1_w_split, p->fc1_b, nullptr, 1, rows, hidden, c, w.hid, w.gemm, w.gemm_bytes, s))) return rc;
    if ((rc = gemm_any(w.hid, p->fc2_w, p->fc2_w_split, p->fc2_b, w.x1, 0, rows, c, hidden, w.x2, w.gemm, w.gemm_bytes, s))) return rc;
    // out = text_norm[i] / img_norm[i] (x2)                          (:443, :452)
    return launch_layernorm(w.x2, p->lno_w, p->lno_b, nullptr, 1, rows, c, out, s);
}

extern "C" size_t pt_gemm_ws_bytes(int M, int N, int K) { return gemm_tc_ws_bytes(M, N, K); }

extern "C" int pt_gemm_nt(const float* A, const float* W, const void* w_split, const float* bias, const float* residual,
                          int act, int M, int N, int K, float* C, void* ws, size_t ws_bytes, pt_stream_t stream) {
    PT_REQUIRE(A && (W || w_split) && C, "pt_gemm_nt: null pointer");
    PT_REQUIRE(act == 0 || act == 1, "pt_gemm_nt: act=%d", act);
    if (w_split != nullptr) {
        PT_REQUIRE(gemm_tc_supported(M, N, K), "pt_gemm_nt: shape M=%d N=%d K=%d unsupported on the tensor-core path", M, N, K);
        return launch_gemm_tc(A, w_split, bias, residual, act, M, N, K, C, ws, ws_bytes, (cudaStream_t)stream);
    }
    return launch_gemm_f32(A, W, bias, residual, act, M, N, K, C, (cudaStream_t)stream);
}
