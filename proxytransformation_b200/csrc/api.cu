// C-ABI glue: error reporting, launch accounting and the multi-kernel ProxyBlock stage (S7).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <atomic>
#include <mutex>
#include <vector>
#include <stdarg.h>
#include <stdlib.h>
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

// ---- profiling: (tag, start, stop) event records, filled only while enabled
static const char* const g_prof_names[PROF_NTAGS] = {
    "minmax_partial", "centres", "ball_query", "offset_net", "cluster_dropout", "point_encoder", "layernorm", "gemm_f32",
    "gemm_tc_3xbf16", "split_bf16", "proxy_attention", "heads", "img_mean", "img_pool", "scatter_mark", "scatter_count",
    "scatter_compact", "misc", "gemm_img_3xbf16"};
struct ProfRec { int tag; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
constexpr size_t PROF_MAX = 1 << 17;

ProfScope::ProfScope(int tag, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.size() >= PROF_MAX) return;
    ProfRec r;
    r.tag = tag;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[slot].b, stream);
}

int launch_layernorm(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                     float* out, cudaStream_t s);
int launch_layernorm_split(const float* x, const float* w, const float* b, const float* add, int add_rows, int rows, int c,
                           float* out, __nv_bfloat16* out_hi, long long out_plane, cudaStream_t s);
bool proxy_attention_mma_supported(int n, int l, int c, int heads);
int launch_proxy_attention_mma(const float* qkv, const float* pt_tok, const uint8_t* mask, int B, int n, int l, int c, int heads,
                               float* o, void* o_split, long long o_plane, cudaStream_t s);
bool proxy_attention_tc_supported(int n, int l, int c, int heads);
int launch_proxy_attention_tc(const void* qk_split, long long qk_plane, int ldq, const void* vt_split, long long vt_plane, long long ldv, int vt_seg,
                              const void* pt_split, long long pt_plane, const uint8_t* mask, int B, int n, int l, int c, int heads,
                              float* o, void* o_split, long long o_plane, cudaStream_t s);
int launch_gemm_f32(const float* A, const float* W, const float* bias, const float* residual, int act, int M, int N, int K,
                    float* C, cudaStream_t s);
int launch_proxy_attention(const float* qkv, const float* pt_tok, const uint8_t* mask, int B, int n, int l, int c,
                           int heads, float* o, cudaStream_t s);

static int gemm_any(const float* A, const float* W, const void* w_split, const float* bias, const float* residual, int act,
                    int M, int N, int K, float* C, void* ws, size_t ws_bytes, cudaStream_t s) {
    if (w_split != nullptr && gemm_tc_supported(M, N, K)) return launch_gemm_tc(A, w_split, bias, residual, act, M, N, K, C, ws, ws_bytes, s);
    return launch_gemm_f32(A, W, bias, residual, act, M, N, K, C, s);
}

struct BlockWs {
    float *u, *qkv, *pt, *o, *x1, *h2, *hid, *x2;
    void* gemm;
    size_t gemm_bytes, total;
    // tensor-core pipeline: bf16 hi/lo planes alias the fp32 buffers of the same stage (2 planes x 2 B == 4 B per element)
    __nv_bfloat16 *u_s, *o_s, *h2_s, *hid_s, *proxy_s;
};

static BlockWs carve_block(void* ws, int B, int n, int l, int c, int hidden) {
    BlockWs r;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = ws ? (void*)((char*)ws + off) : nullptr; off += align_up(bytes, 256); return p; };
    const size_t rows = (size_t)B * n;
    r.u = (float*)take(rows * c * 4);
    r.qkv = (float*)take(rows * 3 * c * 4 + (size_t)B * 8 * c * 4);      // (+ the per-scene padding of the V^T planes, see below)
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
    r.u_s = (__nv_bfloat16*)r.u; r.o_s = (__nv_bfloat16*)r.o; r.h2_s = (__nv_bfloat16*)r.h2; r.hid_s = (__nv_bfloat16*)r.hid;
    r.proxy_s = (__nv_bfloat16*)r.gemm;          // (B*l, c) split planes fit: gemm_bytes >= gemm_tc_ws_bytes(B*l, c, c)
    return r;
}

}  // namespace pt

using namespace pt;

extern "C" int pt_abi_version(void) { return 1; }
extern "C" const char* pt_last_error_string(void) { return g_err; }
extern "C" int64_t pt_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int pt_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
    return PT_OK;
}
extern "C" int pt_profile_num_tags(void) { return PROF_NTAGS; }
extern "C" const char* pt_profile_tag_name(int tag) { return tag >= 0 && tag < PROF_NTAGS ? g_prof_names[tag] : ""; }
extern "C" int pt_profile_read(int tag, double* total_ms, int64_t* launches) {
    PT_REQUIRE(tag >= 0 && tag < PROF_NTAGS && total_ms && launches, "pt_profile_read: bad argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0.0;
    int64_t n = 0;
    for (auto& r : g_prof) {
        if (r.tag != tag) continue;
        PT_CUDA_OK(cudaEventSynchronize(r.b));
        float ms = 0.f;
        PT_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        tot += ms;
        ++n;
    }
    *total_ms = tot;
    *launches = n;
    return PT_OK;
}

extern "C" int pt_profile_timeline(int* tags, double* start_ms, double* end_ms, int max_records, int* n_records) {
    PT_REQUIRE(tags && start_ms && end_ms && n_records && max_records >= 0, "pt_profile_timeline: bad argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (auto& r : g_prof) {
        if (n >= max_records) break;
        PT_CUDA_OK(cudaEventSynchronize(r.b));
        float t0 = 0.f, t1 = 0.f;
        PT_CUDA_OK(cudaEventElapsedTime(&t0, g_prof.front().a, r.a));
        PT_CUDA_OK(cudaEventElapsedTime(&t1, g_prof.front().a, r.b));
        tags[n] = r.tag; start_ms[n] = t0; end_ms[n] = t1;
        ++n;
    }
    *n_records = n;
    return PT_OK;
}

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
    const bool tc = p->qkv_w_split && p->pp_w_split && p->proj_w_split && p->fc1_w_split && p->fc2_w_split && c % 64 == 0 &&
                    hidden % 64 == 0 && proxy_attention_mma_supported(n, l, c, heads);
    if (tc) {
        // Tensor-core pipeline: every GEMM operand is produced directly as bf16 hi/lo planes by the kernel before it
        // (LayerNorm, attention, fc1 epilogue), so no separate splitting passes and no fp32 copies of u / o / h2 / hid.
        auto gemm = [&](const __nv_bfloat16* a, int M, int K, const void* w, int N, const float* bias, const float* res, int act,
                        float* C, __nv_bfloat16* Cs) {
            GemmTc gp;
            gp.M = M; gp.N = N; gp.K = K;
            gp.a_split = a; gp.a_rows = M; gp.a_cols = K; gp.lda = K;
            gp.w_split = w; gp.w_rows = N; gp.ldw = K;
            gp.bias = bias; gp.residual = res; gp.act = act;
            gp.C = C; gp.ldc = N;
            gp.c_split = Cs; gp.cs_plane = (long long)M * N; gp.ldcs = N;
            return launch_gemm_tc_ex(gp, s);
        };
        // u = LN1(x) + bias[m]                                        (:274, :212-217)
        if ((rc = launch_layernorm_split(x, p->ln1_w, p->ln1_b, p->pos_bias, n, rows, c, nullptr, w.u_s, (long long)rows * c, s))) return rc;
        static const bool no_tc_attn = getenv("PT_ATTN_MMA") != nullptr;          // debug: force the mma.sync attention kernel
        if (!no_tc_attn && proxy_attention_tc_supported(n, l, c, heads)) {
            // tcgen05 / TMEM attention (attn_tc.cu): every operand is a bf16 hi/lo plane pair written by a projection GEMM.
            // The fp32 qkv buffer (rows x 3c x 4 B) holds [Q|K] planes (2 x rows x 2c) followed by the V^T planes (2 x c x B*npad8): every
            // scene's columns start on a 16-byte boundary (n padded to a multiple of 8) whatever n is.
            const int npad8 = (n + 7) & ~7;
            const long long vt_cols = (long long)B * npad8;
            __nv_bfloat16* qk_s = (__nv_bfloat16*)w.qkv;
            __nv_bfloat16* vt_s = qk_s + (size_t)2 * rows * 2 * c;
            __nv_bfloat16* pt_s = (__nv_bfloat16*)w.pt;
            {   // [Q|K|V] = u Wqkv^T (:221): Q and K as row-major planes, V TRANSPOSED (V^T planes, K-major over the clusters:
                // the B operand layout of the value contraction) straight from the epilogue
                GemmTc gp;
                gp.M = rows; gp.N = 3 * c; gp.K = c;
                gp.a_split = w.u_s; gp.a_rows = rows; gp.a_cols = c; gp.lda = c;
                gp.w_split = p->qkv_w_split; gp.w_rows = 3 * c; gp.ldw = c;
                gp.bias = p->qkv_b;
                gp.c_split = qk_s; gp.cs_plane = (long long)rows * 2 * c; gp.ldcs = 2 * c;
                gp.ct_split = vt_s; gp.ct_col0 = 2 * c; gp.ct_ld = vt_cols; gp.ct_plane = (long long)c * vt_cols;
                gp.ct_seg = n; gp.ct_seg_pad = npad8;
                if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
            }
            // Pt = proxy Wp^T + bp                                     (:223)
            if ((rc = split_rows_bf16(proxy, (long long)B * l * c, w.proxy_s, w.proxy_s + (size_t)B * l * c, s))) return rc;
            {
                GemmTc gp;
                gp.M = B * l; gp.N = c; gp.K = c;
                gp.a_split = w.proxy_s; gp.a_rows = B * l; gp.a_cols = c; gp.lda = c;
                gp.w_split = p->pp_w_split; gp.w_rows = c; gp.ldw = c;
                gp.bias = p->pp_b;
                gp.c_split = pt_s; gp.cs_plane = (long long)B * l * c; gp.ldcs = c;
                if ((rc = launch_gemm_tc_ex(gp, s))) return rc;
            }
            // two-stage proxy attention                                (:225-252)
            if ((rc = launch_proxy_attention_tc(qk_s, (long long)rows * 2 * c, 2 * c, vt_s, (long long)c * vt_cols, vt_cols, npad8, pt_s,
                                                (long long)B * l * c, mask, B, n, l, c, heads, nullptr, w.o_s, (long long)rows * c, s)))
                return rc;
        } else {
        // [Q|K|V] = u Wqkv^T                                          (:221)
        if ((rc = gemm(w.u_s, rows, c, p->qkv_w_split, 3 * c, p->qkv_b, nullptr, 0, w.qkv, nullptr))) return rc;
        // Pt = proxy Wp^T + bp                                        (:223)
        if ((rc = split_rows_bf16(proxy, (long long)B * l * c, w.proxy_s, w.proxy_s + (size_t)B * l * c, s))) return rc;
        if ((rc = gemm(w.proxy_s, B * l, c, p->pp_w_split, c, p->pp_b, nullptr, 0, w.pt, nullptr))) return rc;
        // two-stage proxy attention                                   (:225-252)
        if ((rc = launch_proxy_attention_mma(w.qkv, w.pt, mask, B, n, l, c, heads, nullptr, w.o_s, (long long)rows * c, s))) return rc;
        }
        // x1 = x + (o Wo^T + bo)                                      (:255, :274)
        if ((rc = gemm(w.o_s, rows, c, p->proj_w_split, c, p->proj_b, x, 0, w.x1, nullptr))) return rc;
        // x2 = x1 + fc2(GELU(fc1(LN2(x1))))                           (:275)
        if ((rc = launch_layernorm_split(w.x1, p->ln2_w, p->ln2_b, nullptr, 1, rows, c, nullptr, w.h2_s, (long long)rows * c, s))) return rc;
        if ((rc = gemm(w.h2_s, rows, c, p->fc1_w_split, hidden, p->fc1_b, nullptr, 1, nullptr, w.hid_s))) return rc;
        if ((rc = gemm(w.hid_s, rows, hidden, p->fc2_w_split, c, p->fc2_b, w.x1, 0, w.x2, nullptr))) return rc;
        return launch_layernorm(w.x2, p->lno_w, p->lno_b, nullptr, 1, rows, c, out, s);
    }
    // fp32 CUDA-core pipeline (no split weights supplied, or a shape the tensor-core kernels do not cover)
    // u = LN1(x) + bias[m]                                           (:274, :212-217)
    if ((rc = launch_layernorm(x, p->ln1_w, p->ln1_b, p->pos_bias, n, rows, c, w.u, s))) return rc;
    // [Q|K|V] = u Wqkv^T                                             (:221)
    if ((rc = gemm_any(w.u, p->qkv_w, p->qkv_w_split, p->qkv_b, nullptr, 0, rows, 3 * c, c, w.qkv, w.gemm, w.gemm_bytes, s))) return rc;
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

extern "C" int pt_gemm_tc(const pt_gemm_tc_desc* d, pt_stream_t stream) {
    PT_REQUIRE(d != nullptr, "pt_gemm_tc: null descriptor");
    GemmTc p;
    p.M = d->M; p.N = d->N; p.K = d->K; p.batch = d->batch;
    p.a_split = d->a_split; p.a_rows = d->a_rows; p.a_cols = d->a_cols; p.lda = d->lda; p.a_koff_z = d->a_koff_z;
    p.w_split = d->w_split; p.w_rows = d->w_rows; p.ldw = d->ldw; p.w_row_z = d->w_row_z;
    p.bias = d->bias; p.bias_off_z = d->bias_off_z; p.residual = d->residual; p.act = d->act;
    p.C = d->C; p.ldc = d->ldc; p.c_off_z = d->c_off_z;
    p.c_split = d->c_split; p.cs_plane = d->cs_plane; p.ldcs = d->ldcs; p.cs_off_z = d->cs_off_z; p.bn = d->bn;
    return launch_gemm_tc_ex(p, (cudaStream_t)stream);
}
