/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product path (proxytransformation_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * CPU restatement of the two geometric ops the reference takes from pytorch3d
 * (un-vendored, unpinned dependency: requirements/run.txt:6 "pytorch3d"; call
 * sites embodiedscan/models/necks/preshape_norm_reverse_drop.py:10,56,65,393).
 * The algorithm restated is the published pytorch3d 0.7.x CPU one:
 *   - ball_query: pytorch3d/csrc/ball_query/ball_query_cpu.cpp — for every
 *     query centre scan p2 in ascending index, keep the first K with
 *     d2 < r*r (strict), d2 accumulated over dims 0,1,2; idx pre-filled -1.
 *   - sample_farthest_points: csrc/sample_farthest_points/sample_farthest_points_cpu.cpp
 *     and the in-tree python copy preshape_norm_reverse_drop.py:527-625 — start
 *     index 0, running min of squared distances from +inf, first arg-max.
 * Compile with -ffp-contract=off (no FMA) and without -march=native; see build.py.
 * Single-threaded on purpose: upstream's CPU loops are single-threaded.
 */
#include <stdint.h>
#include <float.h>
#include <math.h>

/* p1: (B,M,3) centres, p2: (B,N,3) points, idx out: (B,M,K) int64, knn out: (B,M,K,3) (pad = 0.0,
 * the python wrapper's masked_gather, preshape_norm_reverse_drop.py:627-672). scanned (B,M) may be NULL:
 * number of p2 entries examined (diagnostics for the benchmark only). */
void oracle_ball_query(const float* p1, const float* p2, int64_t B, int64_t M, int64_t N, int64_t K,
                       float radius, int64_t* idx, float* knn, int64_t* scanned) {
    const float r2 = radius * radius;
    for (int64_t b = 0; b < B; ++b) {
        const float* P = p2 + b * N * 3;
        for (int64_t i = 0; i < M; ++i) {
            const float* c = p1 + (b * M + i) * 3;
            int64_t* out = idx + (b * M + i) * K;
            float* kn = knn + (b * M + i) * K * 3;
            for (int64_t k = 0; k < K; ++k) { out[k] = -1; kn[3 * k] = kn[3 * k + 1] = kn[3 * k + 2] = 0.0f; }
            int64_t cnt = 0, j = 0;
            for (; j < N && cnt < K; ++j) {
                float d2 = 0.0f;
                for (int d = 0; d < 3; ++d) {
                    float diff = c[d] - P[j * 3 + d];
                    d2 += diff * diff;
                }
                if (d2 < r2) {
                    out[cnt] = j;
                    kn[3 * cnt] = P[j * 3]; kn[3 * cnt + 1] = P[j * 3 + 1]; kn[3 * cnt + 2] = P[j * 3 + 2];
                    ++cnt;
                }
            }
            if (scanned) scanned[b * M + i] = j;
        }
    }
}

/* pts: (B,P,3); out: (B,K) int64.  K <= P assumed by the caller (the reference always has it, :388-393). */
void oracle_fps(const float* pts, int64_t B, int64_t P, int64_t K, int64_t* out, float* scratch /* P floats */) {
    for (int64_t b = 0; b < B; ++b) {
        const float* U = pts + b * P * 3;
        int64_t* o = out + b * K;
        for (int64_t k = 0; k < K; ++k) o[k] = -1;
        for (int64_t p = 0; p < P; ++p) scratch[p] = INFINITY;
        int64_t sel = 0;
        o[0] = 0;
        int64_t kn = K < P ? K : P;
        for (int64_t k = 1; k < kn; ++k) {
            float best = -1.0f;
            int64_t besti = 0;
            for (int64_t p = 0; p < P; ++p) {
                float d2 = 0.0f;
                for (int d = 0; d < 3; ++d) {
                    float diff = U[sel * 3 + d] - U[p * 3 + d];
                    d2 += diff * diff;
                }
                float m = d2 < scratch[p] ? d2 : scratch[p];
                scratch[p] = m;
                if (m > best) { best = m; besti = p; }   /* strict > keeps the first maximum */
            }
            sel = besti;
            o[k] = sel;
        }
    }
}
