/*
 * TEST INFRASTRUCTURE — plain-C restatement of the hot path, used only as a
 * checker (tests/, smoke, bench cpu legs).  Never linked into the product.
 *
 * Follows, line for line in meaning (paths under
 * /root/reference/articulation3d/articulation3d/):
 *   a3do_project : utils/vis.py:86-102 (get_pcd), utils/opt_utils.py:420-435 /
 *                  :553-574 / :724-728 (pytorch3d transform chain, restated
 *                  [3P-unverified]), utils/vis.py:62-75 (project2D) and the
 *                  splat loop utils/opt_utils.py:438-457
 *   a3do_score   : utils/opt_utils.py:464-477 (inter / union / iou / argmax)
 *   a3do_pack    : the `> 0.5` of :471 and `.nonzero()` of :409
 * It is pinned against oracle/restated.py (tests/test_c_oracle.py), which is in
 * turn pinned against the reference's own outputs (tests/golden/).
 *
 * Build: see oracle/Makefile — -ffp-contract=off so no FMA is ever formed; every
 * product and sum below is separately rounded, left to right.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static int pitch_words(int W) { return (((W + 31) >> 5) + 3) & ~3; }

int a3do_pitch_words(int W) { return pitch_words(W); }

/* dense fp32 (n,H,W) -> bits; mode 0: v > thresh, mode 1: v != 0 */
void a3do_pack_f32(const float* src, int64_t n, int H, int W, float thresh, int mode, uint32_t* bits) {
    const int pitch = pitch_words(W);
    memset(bits, 0, (size_t)n * H * pitch * 4);
    for (int64_t m = 0; m < n; ++m)
        for (int r = 0; r < H; ++r) {
            const float* row = src + ((size_t)m * H + r) * W;
            uint32_t* out = bits + ((size_t)m * H + r) * pitch;
            for (int x = 0; x < W; ++x) {
                const int set = mode ? (row[x] != 0.0f) : (row[x] > thresh);
                if (set) out[x >> 5] |= 1u << (x & 31);
            }
        }
}

/* `.long()` on x86 (truncate; NaN/inf/out-of-range -> INT64_MIN) then the clamp of
 * opt_utils.py:445-450 */
static int clamp_index(float v, int n) {
    int64_t i;
    if (!(fabsf(v) < 9.223372036854775807e18f)) i = INT64_MIN;
    else i = (int64_t)v;
    if (i >= n) i = n - 1;
    if (i < 0) i = 0;
    return (int)i;
}

/* one source mask -> A candidate masks.  mode 0 SEQ, 1 COMPOSED, 2 TRANSLATE.
 * xform[A][12]: R (row-major, p' = p*R) then t. */
void a3do_project(const double* kinv, float f, float cx, float cy, int H, int W,
                  const uint32_t* src_bits, const float* normal, float offset, const float* pivot,
                  int mode, const float* xform, int A, uint32_t* proj_bits) {
    const int pitch = pitch_words(W);
    const size_t words = (size_t)H * pitch;
    memset(proj_bits, 0, (size_t)A * words * 4);
    const double n0 = normal[0], n1 = normal[1], n2 = normal[2], off = offset;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (!((src_bits[(size_t)y * pitch + (x >> 5)] >> (x & 31)) & 1u)) continue;
            /* get_pcd in float64 */
            const double xd = x, yd = y;
            const double rx = (kinv[0] * xd + kinv[1] * yd) + kinv[2] * 1.0;
            const double ry = (kinv[3] * xd + kinv[4] * yd) + kinv[5] * 1.0;
            const double rz = (kinv[6] * xd + kinv[7] * yd) + kinv[8] * 1.0;
            const double dot = (n0 * rx + n1 * ry) + n2 * rz;
            const double depth = off / dot;
            float px = (float)(depth * rx), py = (float)(depth * ry), pz = (float)(depth * rz);
            if (!(isfinite(px) && isfinite(py) && isfinite(pz))) px = py = pz = NAN;
            for (int a = 0; a < A; ++a) {
                const float* m = xform + 12 * (size_t)a;
                float qx = px, qy = py, qz = pz, sx, sy, sz;
                if (mode == 2) {
                    sx = qx + m[9]; sy = qy + m[10]; sz = qz + m[11];
                } else {
                    if (mode == 0) { qx = qx - pivot[0]; qy = qy - pivot[1]; qz = qz - pivot[2]; }
                    sx = (qx * m[0] + qy * m[3]) + qz * m[6];
                    sy = (qx * m[1] + qy * m[4]) + qz * m[7];
                    sz = (qx * m[2] + qy * m[5]) + qz * m[8];
                    if (mode == 0) { sx = sx + pivot[0]; sy = sy + pivot[1]; sz = sz + pivot[2]; }
                    else { sx = sx + m[9]; sy = sy + m[10]; sz = sz + m[11]; }
                }
                /* project2D: K @ p, divide by w */
                const float u = (f * sx + 0.0f * sy) + cx * sz;
                const float v = (0.0f * sx + f * sy) + cy * sz;
                const float w = (0.0f * sx + 0.0f * sy) + 1.0f * sz;
                const int col = clamp_index(u / w, W);
                const int row = clamp_index(v / w, H);
                proj_bits[(size_t)a * words + (size_t)row * pitch + (col >> 5)] |= 1u << (col & 31);
            }
        }
}

/* inter/union tables + first-max argmax (NaN counts as max) */
void a3do_score(int H, int W, const uint32_t* tgt_bits, int T, const uint32_t* proj_bits, int A,
                int32_t* inter, int32_t* uni, int32_t* best_cand, float* best_iou) {
    const size_t words = (size_t)H * pitch_words(W);
    for (int t = 0; t < T; ++t) {
        const uint32_t* tb = tgt_bits + (size_t)t * words;
        int best = -1;
        float best_v = 0.0f;
        for (int a = 0; a < A; ++a) {
            const uint32_t* pb = proj_bits + (size_t)a * words;
            int32_t ni = 0, nu = 0;
            for (size_t i = 0; i < words; ++i) {
                ni += __builtin_popcount(tb[i] & pb[i]);
                nu += __builtin_popcount(tb[i] | pb[i]);
            }
            inter[(size_t)t * A + a] = ni;
            uni[(size_t)t * A + a] = nu;
            const float iou = (float)ni / (float)nu;
            if (best < 0) { best = a; best_v = iou; }
            else if (!isnan(best_v) && (isnan(iou) || iou > best_v)) { best = a; best_v = iou; }
        }
        best_cand[t] = best;
        best_iou[t] = best_v;
    }
}
