/*
 * rgbd_oracle.c -- CPU restatement of RGBD-GAN's 3D-consistency hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (rgbd_gan_b200/) never links or calls it.
 *
 * Pinning: the reference ships no tests or golden vectors, and its Chainer/CuPy
 * dependency cannot be installed here.  This restatement is pinned against
 * tests/golden/*.npz, which are produced by running the reference's own unmodified
 * source files over a restatement of the Chainer-v7 ops they call
 * (tests/golden/make_golden.py + chainer_shim.py).  tests/test_oracle_golden.py
 * holds the comparison (indices/masks/new_zp bit-exact, loss/grads 1e-5).
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference root).  Arithmetic is IEEE fp32 with one rounding per written
 * operation: compile with -ffp-contract=off; the only fused operations are the
 * explicit fmaf() chains that reproduce what BLAS sgemm does for the K=3 / K=4
 * products on the reference's NumPy CPU path (verified against the golden
 * vectors: new_zp and voxel_coords are bit-equal).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* numpy `astype(int32)` on x86: truncation; out-of-range / NaN -> INT32_MIN. */
static inline int32_t trunc_i32(float x)
{
    if (!(fabsf(x) < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}

typedef struct {
    /* per-pixel quantities of one warp direction */
    float q0, q1, q2, zc, vcol, urow;
    int32_t u0, v0, v1;          /* masked indices (0 where !m)          */
    float a, bb, cc, d;          /* unmasked 1-D weights                  */
    float w1, w2, w3, w4;        /* masked 2-D weights                    */
    int m;
} orc_px;

/* warp / inv_warp (common/loss_functions.py:171-182) followed by the coordinate
 * part of bilinear (:199-225).  Mm = K R K^-1 (row major 3x3), cv = vector that is
 * SUBTRACTED (direction 1: (K R) t; direction 2: -(K t), so that t - cv == t + K t). */
static inline void orc_project(const float *Mm, const float *cv, float z, int i, int j,
                               int H, int W, orc_px *o)
{
    const float x = (float)j, y = (float)i;            /* p = (col,row,1)  :59-61  */
    const float P0 = z * x, P1 = z * y, P2 = z;        /* z * p            :174    */
    float q[3];
    for (int r = 0; r < 3; ++r) {                      /* F.matmul -> sgemm, K=3   */
        float t = Mm[3 * r + 0] * P0;
        t = fmaf(Mm[3 * r + 1], P1, t);
        t = fmaf(Mm[3 * r + 2], P2, t);
        q[r] = t - cv[r];
    }
    o->q0 = q[0]; o->q1 = q[1]; o->q2 = q[2];
    float zc = q[2];                                   /* F.clip(zp2,1e-4,1e4) :199 */
    if (zc < 1e-4f) zc = 1e-4f;
    if (zc > 10000.0f) zc = 10000.0f;
    o->zc = zc;
    o->vcol = q[0] / zc;                               /* :199, renamed at :202    */
    o->urow = q[1] / zc;                               /* :200                     */
    int32_t u0 = trunc_i32(o->urow), v0 = trunc_i32(o->vcol);       /* :203-206   */
    int32_t u1 = (int32_t)((uint32_t)u0 + 1u), v1 = (int32_t)((uint32_t)v0 + 1u);
    o->a = (float)u1 - o->urow;  o->bb = o->urow - (float)u0;       /* :209-212   */
    o->cc = (float)v1 - o->vcol; o->d = o->vcol - (float)v0;
    o->m = (o->urow >= 0.0f) && (o->urow < (float)(H - 1)) &&
           (o->vcol >= 0.0f) && (o->vcol < (float)(W - 1)) && (q[2] > 1e-4f); /* :215-216 */
    const float mf = o->m ? 1.0f : 0.0f;
    o->w1 = (o->a * o->cc) * mf;  o->w2 = (o->bb * o->cc) * mf;     /* :222-225   */
    o->w3 = (o->a * o->d) * mf;   o->w4 = (o->bb * o->d) * mf;
    o->u0 = o->m ? u0 : 0;  o->v0 = o->m ? v0 : 0;  o->v1 = o->m ? v1 : 0;  /* :218-221 (u1 := u0) */
}

/* the 4-term blend of :226-227 with both row taps on row u0 (quirk Q1, :219) */
static inline float orc_blend(const orc_px *p, float A, float Bv)
{
    return ((p->w1 * A + p->w2 * A) + p->w3 * Bv) + p->w4 * Bv;
}

/* One direction of LossFuncRotate.__call__ (:93-144).  src = image whose pixels are
 * projected (supplies depth and the targets), oth = image that is sampled.
 * sums[0] += sum over rgb channels of |diff| (or diff^2), sums[1] += depth channel.
 * Optional debug outputs (may be NULL): new_zp (HW,3), warped (HW,C), idx (HW,3) int32
 * = masked u0,v0,v1, mask (HW) = m, occ (HW) = not_occluded (1 when occlusion off). */
static void orc_direction_fwd(const float *src, const float *oth, const float *Mm, const float *cv,
                              int C, int H, int W, int norm, int occlusion, float max_depth,
                              float min_depth, double *sums, float *new_zp, float *warped,
                              int32_t *idx, uint8_t *mask, uint8_t *occ)
{
    const int HW = H * W;
    double s_rgb = 0.0, s_d = 0.0;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const int n = i * W + j;
            const float z = src[(size_t)(C - 1) * HW + n];
            orc_px p;
            orc_project(Mm, cv, z, i, j, H, W, &p);
            const size_t ta = (size_t)p.u0 * W + p.v0, tb = (size_t)p.u0 * W + p.v1;
            const float mf = p.m ? 1.0f : 0.0f;
            const float wd = orc_blend(&p, oth[(size_t)(C - 1) * HW + ta], oth[(size_t)(C - 1) * HW + tb]);
            float of = 1.0f;
            if (occlusion) of = (wd > p.q2) ? 1.0f : 0.0f;                     /* :114 */
            float sf = 1.0f;
            if (!isnan(max_depth)) sf *= (z < max_depth) ? 1.0f : 0.0f;        /* :122 */
            if (!isnan(min_depth)) sf *= (z > min_depth) ? 1.0f : 0.0f;        /* :130 */
            for (int ch = 0; ch < C; ++ch) {
                const float wv = (ch == C - 1) ? wd
                                               : orc_blend(&p, oth[(size_t)ch * HW + ta], oth[(size_t)ch * HW + tb]);
                float tg = ((ch == C - 1) ? p.q2 : src[(size_t)ch * HW + n]) * mf;  /* :107-108 */
                float wm = wv;
                if (occlusion) { wm = wm * of; tg = tg * of; }                 /* :116-119 */
                if (!isnan(max_depth) || !isnan(min_depth)) { wm = wm * sf; tg = tg * sf; }
                const float diff = wm - tg;
                const double t = (norm == 1) ? fabs((double)diff) : (double)diff * (double)diff;
                if (ch == C - 1) s_d += t; else s_rgb += t;
                if (warped) warped[(size_t)n * C + ch] = wv;    /* bilinear() output, pre-occlusion */
            }
            if (new_zp) { new_zp[3 * n] = p.q0; new_zp[3 * n + 1] = p.q1; new_zp[3 * n + 2] = p.q2; }
            if (idx) { idx[3 * n] = p.u0; idx[3 * n + 1] = p.v0; idx[3 * n + 2] = p.v1; }
            if (mask) mask[n] = (uint8_t)p.m;
            if (occ) occ[n] = (uint8_t)(of != 0.0f);
        }
    sums[0] += s_rgb;
    sums[1] += s_d;
}

/* LossFuncRotate.__call__ forward (common/loss_functions.py:63-146).
 * img, img_rot: (B,C,H,W).  M,c: direction img->img_rot (B,9),(B,3); Mi,ci: the inverse
 * direction.  n_pairs_global: number of pairs the means are taken over (== B unless the
 * batch is sharded).  loss_parts[4] = the four means {rgb, rgb_rot, depth, depth_rot}
 * restricted to these B pairs (:141-144; sum over shards, then
 * loss = (p0+p1) + (p2*lambda + p3*lambda)).
 * Debug outputs are for both directions, direction-major: new_zp (2B,HW,3) as :146. */
ORC_API int orc_consistency_fwd(const float *img, const float *img_rot, const float *M, const float *c,
                                const float *Mi, const float *ci, int B, int C, int H, int W, int norm,
                                int occlusion, float max_depth, float min_depth, long long n_pairs_global,
                                double *loss_parts, float *new_zp, float *warped, int32_t *idx,
                                uint8_t *mask, uint8_t *occ)
{
    const size_t HW = (size_t)H * W, img_sz = (size_t)C * HW;
    double *part = (double *)calloc((size_t)B * 4, sizeof(double));
    if (!part) return -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const size_t n0 = (size_t)b * HW, n1 = ((size_t)B + b) * HW;
        orc_direction_fwd(img + b * img_sz, img_rot + b * img_sz, M + 9 * b, c + 3 * b, C, H, W, norm,
                          occlusion, max_depth, min_depth, part + 4 * b,
                          new_zp ? new_zp + 3 * n0 : NULL, warped ? warped + n0 * C : NULL,
                          idx ? idx + 3 * n0 : NULL, mask ? mask + n0 : NULL, occ ? occ + n0 : NULL);
        orc_direction_fwd(img_rot + b * img_sz, img + b * img_sz, Mi + 9 * b, ci + 3 * b, C, H, W, norm,
                          occlusion, max_depth, min_depth, part + 4 * b + 2,
                          new_zp ? new_zp + 3 * n1 : NULL, warped ? warped + n1 * C : NULL,
                          idx ? idx + 3 * n1 : NULL, mask ? mask + n1 : NULL, occ ? occ + n1 : NULL);
    }
    double s[4] = {0, 0, 0, 0};
    for (int b = 0; b < B; ++b) { s[0] += part[4 * b]; s[2] += part[4 * b + 1]; s[1] += part[4 * b + 2]; s[3] += part[4 * b + 3]; }
    free(part);
    const double N = (double)n_pairs_global * (double)HW;
    loss_parts[0] = s[0] / (N * (C - 1));
    loss_parts[1] = s[1] / (N * (C - 1));
    loss_parts[2] = s[2] / N;
    loss_parts[3] = s[3] / N;
    return 0;
}

/* Backward of one direction (Chainer autograd of :93-144 written in closed form, see
 * SURVEY.md 8(a) row a11).  Adds into g_src (own-pixel terms) and g_oth (scatter).
 *   k_rgb, k_d : upstream coefficient per element, i.e. gy*fp32(1/size) for L1 and
 *                gy*fp32(2/size) for L2, with lambda folded into k_d (host computes
 *                them in Chainer's order, see orc_consistency_bwd).
 *   g_zp       : optional upstream gradient of new_zp (HW,3) (second output of __call__). */
static void orc_direction_bwd(const float *src, const float *oth, const float *Mm, const float *cv,
                              int C, int H, int W, int norm, int occlusion, float max_depth,
                              float min_depth, float k_rgb, float k_d, const float *g_zp,
                              float *g_src, float *g_oth)
{
    const int HW = H * W;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const int n = i * W + j;
            const float z = src[(size_t)(C - 1) * HW + n];
            orc_px p;
            orc_project(Mm, cv, z, i, j, H, W, &p);
            const size_t ta = (size_t)p.u0 * W + p.v0, tb = (size_t)p.u0 * W + p.v1;
            const float mf = p.m ? 1.0f : 0.0f;
            const float wd = orc_blend(&p, oth[(size_t)(C - 1) * HW + ta], oth[(size_t)(C - 1) * HW + tb]);
            float of = 1.0f, sf = 1.0f;
            if (occlusion) of = (wd > p.q2) ? 1.0f : 0.0f;
            if (!isnan(max_depth)) sf *= (z < max_depth) ? 1.0f : 0.0f;
            if (!isnan(min_depth)) sf *= (z > min_depth) ? 1.0f : 0.0f;
            const float os = of * sf;
            float GA = 0.0f, GB = 0.0f, e_depth = 0.0f;
            for (int ch = 0; ch < C; ++ch) {
                const float A = oth[(size_t)ch * HW + ta], Bv = oth[(size_t)ch * HW + tb];
                const float wv = (ch == C - 1) ? wd : orc_blend(&p, A, Bv);
                const float tg = (((ch == C - 1) ? p.q2 : src[(size_t)ch * HW + n]) * mf) * os;
                const float diff = wv * os - tg;
                const float k = (ch == C - 1) ? k_d : k_rgb;
                float e;                                   /* d loss / d (masked warped)  */
                if (norm == 1) e = k * (float)((diff > 0.0f) - (diff < 0.0f));
                else e = k * diff;
                e = e * os;                                /* through `warped * mask`     */
                /* gather backward: GetItem -> add.at on the sampled image (:226-227)    */
                g_oth[(size_t)ch * HW + ta] += e * p.w1;
                g_oth[(size_t)ch * HW + ta] += e * p.w2;
                g_oth[(size_t)ch * HW + tb] += e * p.w3;
                g_oth[(size_t)ch * HW + tb] += e * p.w4;
                GA += e * A;
                GB += e * Bv;
                if (ch == C - 1) e_depth = e; else g_src[(size_t)ch * HW + n] += -(e * mf);  /* own RGB target */
            }
            /* weights -> coordinates: only the column coordinate carries gradient (Q2) */
            const float g_cc = (GA * mf) * p.a + (GA * mf) * p.bb;
            const float g_d = (GB * mf) * p.a + (GB * mf) * p.bb;
            const float g_v = g_d - g_cc;
            float gq0 = g_v / p.zc;                          /* Div backward             */
            float g_zc = -gq0 * p.q0 / p.zc;
            float gq2 = -(e_depth * mf);                     /* target depth = q2*m      */
            if (p.q2 >= 1e-4f && p.q2 <= 10000.0f) gq2 += g_zc;   /* Clip backward      */
            float gq1 = 0.0f;
            if (g_zp) { gq0 += g_zp[3 * n]; gq1 += g_zp[3 * n + 1]; gq2 += g_zp[3 * n + 2]; }
            /* MatMul backward: gP = M^T gq ; z*p backward: gz = sum_k gP_k p_k          */
            const float gP0 = Mm[0] * gq0 + Mm[3] * gq1 + Mm[6] * gq2;
            const float gP1 = Mm[1] * gq0 + Mm[4] * gq1 + Mm[7] * gq2;
            const float gP2 = Mm[2] * gq0 + Mm[5] * gq1 + Mm[8] * gq2;
            g_src[(size_t)(C - 1) * HW + n] += (gP0 * (float)j + gP1 * (float)i) + gP2;
        }
}

/* Backward of LossFuncRotate.__call__ w.r.t. img and img_rot for upstream loss gradient gy.
 * g_img, g_img_rot are OVERWRITTEN.  lambda_geometric enters as in :143-144.
 * g_new_zp (2B,HW,3) may be NULL. */
ORC_API int orc_consistency_bwd(const float *img, const float *img_rot, const float *M, const float *c,
                                const float *Mi, const float *ci, int B, int C, int H, int W, int norm,
                                int occlusion, float max_depth, float min_depth, float lambda_geo,
                                long long n_pairs_global, float gy, const float *g_new_zp,
                                float *g_img, float *g_img_rot)
{
    const size_t HW = (size_t)H * W, img_sz = (size_t)C * HW;
    const double N = (double)n_pairs_global * (double)HW;
    /* Chainer: MulConstant backward gives lambda*gy; MeanAbsoluteError: gy*fp32(1/size);
     * MeanSquaredError: gy*diff*fp32(2/size). */
    const float two = (norm == 1) ? 1.0f : 2.0f;
    const float k_rgb = gy * (float)(two / (N * (C - 1)));
    const float k_d = (lambda_geo * gy) * (float)(two / N);
    memset(g_img, 0, sizeof(float) * img_sz * B);
    memset(g_img_rot, 0, sizeof(float) * img_sz * B);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        orc_direction_bwd(img + b * img_sz, img_rot + b * img_sz, M + 9 * b, c + 3 * b, C, H, W, norm,
                          occlusion, max_depth, min_depth, k_rgb, k_d,
                          g_new_zp ? g_new_zp + 3 * (size_t)b * HW : NULL,
                          g_img + b * img_sz, g_img_rot + b * img_sz);
        orc_direction_bwd(img_rot + b * img_sz, img + b * img_sz, Mi + 9 * b, ci + 3 * b, C, H, W, norm,
                          occlusion, max_depth, min_depth, k_rgb, k_d,
                          g_new_zp ? g_new_zp + 3 * ((size_t)B + b) * HW : NULL,
                          g_img_rot + b * img_sz, g_img + b * img_sz);
    }
    return 0;
}

/* ---- "next" row: the depth hinge the updaters add right after the loss,
 * loss_rotate += F.mean(F.relu(depth_min - x_fake[:, -1]) ** 2) * lambda_depth   (updater.py:357-359,
 * updater_deepvoxels.py:198-199), over the depth channel of all 2B images.
 * Returns the term in *hinge_out; when g_img / g_img_rot are given, ADDS its gradient for upstream gy
 * to their depth channel (Chainer: MulConstant, Mean, PowVarConst, ReLU, SubFromConstant backward). */
ORC_API int orc_depth_hinge(const float *img, const float *img_rot, int B, int C, int H, int W, float depth_min,
                            float lambda_depth, long long n_pairs_global, float gy, double *hinge_out,
                            float *g_img, float *g_img_rot)
{
    const size_t HW = (size_t)H * W, img_sz = (size_t)C * HW;
    const double n = 2.0 * (double)n_pairs_global * (double)HW;
    const float coef = (lambda_depth * gy) * (float)(1.0 / n);
    double sum = 0.0;
    for (int sel = 0; sel < 2; ++sel) {
        const float *im = sel ? img_rot : img;
        float *g = sel ? g_img_rot : g_img;
        for (int b = 0; b < B; ++b)
            for (size_t k = 0; k < HW; ++k) {
                const size_t at = b * img_sz + (size_t)(C - 1) * HW + k;
                float h = depth_min - im[at];
                if (!(h > 0.0f)) h = 0.0f;
                sum += (double)(h * h);
                if (g) g[at] += -((2.0f * h) * coef);
            }
    }
    *hinge_out = (double)lambda_depth * (sum / n);
    return 0;
}

/* ---- standalone surface: warp / inv_warp / bilinear (common/loss_functions.py:171-228) ---- */

/* new_zp[b,n,:] = M[b] (z[b,n] * p[:,n]) - cv[b]     (:171-175; inv_warp :178-182 with cv = -(K t)) */
ORC_API int orc_warp_fwd(const float *z, const float *M, const float *cv, int B, int H, int W, float *new_zp)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                orc_px p;
                const size_t n = (size_t)i * W + j;
                orc_project(M + 9 * b, cv + 3 * b, z[b * HW + n], i, j, H, W, &p);
                float *o = new_zp + 3 * (b * HW + n);
                o[0] = p.q0; o[1] = p.q1; o[2] = p.q2;
            }
    return 0;
}

/* g_z[b,n] = sum_k (M[b]^T g_zp[b,n,:])_k p_k[n] */
ORC_API int orc_warp_bwd(const float *g_zp, const float *M, int B, int H, int W, float *g_z)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *Mm = M + 9 * b;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const size_t n = (size_t)i * W + j;
                const float *g = g_zp + 3 * (b * HW + n);
                const float gP0 = Mm[0] * g[0] + Mm[3] * g[1] + Mm[6] * g[2];
                const float gP1 = Mm[1] * g[0] + Mm[4] * g[1] + Mm[7] * g[2];
                const float gP2 = Mm[2] * g[0] + Mm[5] * g[1] + Mm[8] * g[2];
                g_z[b * HW + n] = (gP0 * (float)j + gP1 * (float)i) + gP2;
            }
    }
    return 0;
}

static inline void orc_coords(const float *zp, int H, int W, orc_px *o)
{
    /* same as the tail of orc_project, starting from a given zp (:199-225) */
    o->q0 = zp[0]; o->q1 = zp[1]; o->q2 = zp[2];
    float zc = zp[2];
    if (zc < 1e-4f) zc = 1e-4f;
    if (zc > 10000.0f) zc = 10000.0f;
    o->zc = zc;
    o->vcol = zp[0] / zc;
    o->urow = zp[1] / zc;
    int32_t u0 = trunc_i32(o->urow), v0 = trunc_i32(o->vcol);
    int32_t u1 = (int32_t)((uint32_t)u0 + 1u), v1 = (int32_t)((uint32_t)v0 + 1u);
    o->a = (float)u1 - o->urow;  o->bb = o->urow - (float)u0;
    o->cc = (float)v1 - o->vcol; o->d = o->vcol - (float)v0;
    o->m = (o->urow >= 0.0f) && (o->urow < (float)(H - 1)) &&
           (o->vcol >= 0.0f) && (o->vcol < (float)(W - 1)) && (zp[2] > 1e-4f);
    const float mf = o->m ? 1.0f : 0.0f;
    o->w1 = (o->a * o->cc) * mf;  o->w2 = (o->bb * o->cc) * mf;
    o->w3 = (o->a * o->d) * mf;   o->w4 = (o->bb * o->d) * mf;
    o->u0 = o->m ? u0 : 0;  o->v0 = o->m ? v0 : 0;  o->v1 = o->m ? v1 : 0;
}

/* bilinear(img, zp) -> warped (N,C), not_getting_out (N,)   (:185-228) */
ORC_API int orc_bilinear_fwd(const float *img, const float *zp, int B, int C, int H, int W,
                             float *warped, uint8_t *mask)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (size_t n = 0; n < HW; ++n) {
            orc_px p;
            orc_coords(zp + 3 * (b * HW + n), H, W, &p);
            const float *im = img + (size_t)b * C * HW;
            const size_t ta = (size_t)p.u0 * W + p.v0, tb = (size_t)p.u0 * W + p.v1;
            for (int ch = 0; ch < C; ++ch)
                warped[(b * HW + n) * C + ch] = orc_blend(&p, im[ch * HW + ta], im[ch * HW + tb]);
            mask[b * HW + n] = (uint8_t)p.m;
        }
    return 0;
}

/* backward of bilinear: g_warped (N,C) -> g_img (B,C,H,W) [overwritten], g_zp (B,HW,3) [overwritten] */
ORC_API int orc_bilinear_bwd(const float *img, const float *zp, const float *g_warped, int B, int C,
                             int H, int W, float *g_img, float *g_zp)
{
    const size_t HW = (size_t)H * W;
    memset(g_img, 0, sizeof(float) * (size_t)B * C * HW);
    for (int b = 0; b < B; ++b)
        for (size_t n = 0; n < HW; ++n) {
            orc_px p;
            orc_coords(zp + 3 * (b * HW + n), H, W, &p);
            const float *im = img + (size_t)b * C * HW;
            float *gi = g_img + (size_t)b * C * HW;
            const size_t ta = (size_t)p.u0 * W + p.v0, tb = (size_t)p.u0 * W + p.v1;
            const float mf = p.m ? 1.0f : 0.0f;
            float GA = 0.0f, GB = 0.0f;
            for (int ch = 0; ch < C; ++ch) {
                const float e = g_warped[(b * HW + n) * C + ch];
                gi[ch * HW + ta] += e * p.w1;
                gi[ch * HW + ta] += e * p.w2;
                gi[ch * HW + tb] += e * p.w3;
                gi[ch * HW + tb] += e * p.w4;
                GA += e * im[ch * HW + ta];
                GB += e * im[ch * HW + tb];
            }
            const float g_cc = (GA * mf) * p.a + (GA * mf) * p.bb;
            const float g_d = (GB * mf) * p.a + (GB * mf) * p.bb;
            const float g_v = g_d - g_cc;
            const float gq0 = g_v / p.zc;
            const float g_zc = -gq0 * p.q0 / p.zc;
            float *o = g_zp + 3 * (b * HW + n);
            o[0] = gq0;
            o[1] = 0.0f;                                /* row-coordinate gradient cancels (Q2) */
            o[2] = (p.q2 >= 1e-4f && p.q2 <= 10000.0f) ? g_zc : 0.0f;
        }
    return 0;
}

/* ------------------ DeepVoxels projection (deepvoxel/projection.py:48-105) ------------------ */

typedef struct {
    int W, H, D;          /* projection_image_dims[0], [1], frustrum_depth */
    int G;                /* grid_dims (cubic)                              */
    float fx, fy, cx, cy; /* projection_intrinsic                          */
    float voxel_size, near_plane;
} orc_dv_params;

/* voxel coordinates of frustum element l; returns keep flag */
static inline int orc_dv_coords(const orc_dv_params *P, const float *T, int l, float *vc)
{
    const int WH = P->W * P->H;
    const int d = l / WH;                                          /* :64        */
    const int tmp = l - d * WH;                                    /* :65-66     */
    const float yrow = (float)((double)tmp / (double)P->W);        /* :67 true division (Q5) */
    const float xcol = (float)(tmp % P->W);                        /* :68        */
    float zc = (float)d * P->voxel_size;                           /* :73        */
    zc = zc + P->near_plane;                                       /* :74 (fp32, Q6) */
    float xc = (xcol - P->cx) / P->fx;                             /* :78        */
    float yc = (yrow - P->cy) / P->fy;                             /* :79        */
    xc = xc * zc; yc = yc * zc;                                    /* :80        */
    int keep = 1;
    for (int r = 0; r < 3; ++r) {                                  /* xp.dot -> sgemm, K=4 (:82) */
        float g = T[4 * r + 0] * xc;
        g = fmaf(T[4 * r + 1], yc, g);
        g = fmaf(T[4 * r + 2], zc, g);
        g = fmaf(T[4 * r + 3], 1.0f, g);
        float v = g / P->voxel_size;                               /* :87        */
        v = v + (float)P->G / 2.0f;                                /* :88        */
        vc[r] = v;
        keep = keep && (v >= 0.0f) && (v < (float)P->G);           /* :92-96     */
    }
    return keep;
}

/* compute_proj_idcs: returns M (number kept); lin_ind (M) ascending, voxel_coords (3,M) with
 * row stride `ld` (pass ld = W*H*D capacity).  M == 0 is the reference's `None`. */
ORC_API int orc_dv_compute_proj_idcs(const orc_dv_params *P, const float *cam2world, int32_t *lin_ind,
                                     float *voxel_coords, int ld)
{
    const int n = P->W * P->H * P->D;
    int M = 0;
    for (int l = 0; l < n; ++l) {
        float vc[3];
        if (orc_dv_coords(P, cam2world, l, vc)) {
            lin_ind[M] = l;
            voxel_coords[M] = vc[0]; voxel_coords[ld + M] = vc[1]; voxel_coords[2 * ld + M] = vc[2];
            ++M;
        }
    }
    return M;
}

/* the same with the optional grid2world argument (projection.py:53-54,83-84): grid_coords = world2grid . (cam2world .
 * coords), two sgemm products with K = 4 (all four rows of the first feed the second) */
static inline int orc_dv_coords_g2w(const orc_dv_params *P, const float *T, const float *Wg, int l, float *vc)
{
    const int WH = P->W * P->H;
    const int d = l / WH;
    const int tmp = l - d * WH;
    const float yrow = (float)((double)tmp / (double)P->W);
    const float xcol = (float)(tmp % P->W);
    float zc = (float)d * P->voxel_size;
    zc = zc + P->near_plane;
    float xc = (xcol - P->cx) / P->fx;
    float yc = (yrow - P->cy) / P->fy;
    xc = xc * zc; yc = yc * zc;
    float gc[4];
    for (int r = 0; r < 4; ++r) {
        float g = T[4 * r + 0] * xc;
        g = fmaf(T[4 * r + 1], yc, g);
        g = fmaf(T[4 * r + 2], zc, g);
        gc[r] = fmaf(T[4 * r + 3], 1.0f, g);
    }
    int keep = 1;
    for (int r = 0; r < 3; ++r) {
        float g = Wg[4 * r + 0] * gc[0];
        g = fmaf(Wg[4 * r + 1], gc[1], g);
        g = fmaf(Wg[4 * r + 2], gc[2], g);
        g = fmaf(Wg[4 * r + 3], gc[3], g);
        float v = g / P->voxel_size;
        v = v + (float)P->G / 2.0f;
        vc[r] = v;
        keep = keep && (v >= 0.0f) && (v < (float)P->G);
    }
    return keep;
}

ORC_API int orc_dv_compute_proj_idcs_g2w(const orc_dv_params *P, const float *cam2world, const float *world2grid,
                                         int32_t *lin_ind, float *voxel_coords, int ld)
{
    const int n = P->W * P->H * P->D;
    int M = 0;
    for (int l = 0; l < n; ++l) {
        float vc[3];
        if (orc_dv_coords_g2w(P, cam2world, world2grid, l, vc)) {
            lin_ind[M] = l;
            voxel_coords[M] = vc[0]; voxel_coords[ld + M] = vc[1]; voxel_coords[2 * ld + M] = vc[2];
            ++M;
        }
    }
    return M;
}

typedef struct { int x0, x1, y0, y1, z0, z1; float wx0, wx1, wy0, wy1, wz0, wz1; } orc_taps;

/* deepvoxel/deepvoxel.py:394-412: axis swap (Q8), truncation, clamp, fp64 fractions (Q7) */
static inline void orc_dv_taps(const float *vc, int G, orc_taps *t)
{
    const float X = vc[2], Y = vc[1], Z = vc[0];
    t->x0 = trunc_i32(X); t->y0 = trunc_i32(Y); t->z0 = trunc_i32(Z);
    t->x1 = t->x0 + 1 > G - 1 ? G - 1 : (t->x0 + 1 < 0 ? 0 : t->x0 + 1);
    t->y1 = t->y0 + 1 > G - 1 ? G - 1 : (t->y0 + 1 < 0 ? 0 : t->y0 + 1);
    t->z1 = t->z0 + 1 > G - 1 ? G - 1 : (t->z0 + 1 < 0 ? 0 : t->z0 + 1);
    const double fx = (double)X - (double)t->x0, fy = (double)Y - (double)t->y0, fz = (double)Z - (double)t->z0;
    t->wx1 = (float)fx; t->wx0 = (float)(1.0 - fx);
    t->wy1 = (float)fy; t->wy0 = (float)(1.0 - fy);
    t->wz1 = (float)fz; t->wz0 = (float)(1.0 - fz);
}

/* corner order of :416-423 */
#define ORC_CORNERS(t)                                                                         \
    const int cx_[8] = {t.x0, t.x1, t.x0, t.x0, t.x1, t.x0, t.x1, t.x1};                       \
    const int cy_[8] = {t.y0, t.y0, t.y1, t.y0, t.y0, t.y1, t.y1, t.y1};                       \
    const int cz_[8] = {t.z0, t.z0, t.z0, t.z1, t.z1, t.z1, t.z0, t.z1};                       \
    const float ax_[8] = {t.wx0, t.wx1, t.wx0, t.wx0, t.wx1, t.wx0, t.wx1, t.wx1};             \
    const float ay_[8] = {t.wy0, t.wy0, t.wy1, t.wy0, t.wy0, t.wy1, t.wy1, t.wy1};             \
    const float az_[8] = {t.wz0, t.wz0, t.wz0, t.wz1, t.wz1, t.wz1, t.wz0, t.wz1};

/* interpolate_trilinear forward for ONE sample: grid (F,G,G,G) -> out (F, D*H*W), zero where not listed */
ORC_API int orc_dv_trilinear_fwd(const float *grid, const int32_t *lin_ind, const float *voxel_coords,
                                 int ld, int M, int F, int G, int n_frustum, float *out)
{
    const size_t G3 = (size_t)G * G * G;
    memset(out, 0, sizeof(float) * (size_t)F * n_frustum);
    for (int m = 0; m < M; ++m) {
        const float vc[3] = {voxel_coords[m], voxel_coords[ld + m], voxel_coords[2 * ld + m]};
        orc_taps t;
        orc_dv_taps(vc, G, &t);
        ORC_CORNERS(t)
        for (int f = 0; f < F; ++f) {
            const float *g = grid + f * G3;
            float acc = 0.0f;
            for (int k = 0; k < 8; ++k) {
                const float v = g[((size_t)cx_[k] * G + cy_[k]) * G + cz_[k]];
                const float term = ((v * ax_[k]) * ay_[k]) * az_[k];
                acc = (k == 0) ? term : acc + term;
            }
            out[(size_t)f * n_frustum + lin_ind[m]] += acc;          /* F.scatter_add into zeros (:425) */
        }
    }
    return 0;
}

/* backward ("lift"): g_out (F, D*H*W) -> g_grid (F,G,G,G) [overwritten] */
ORC_API int orc_dv_trilinear_bwd(const float *g_out, const int32_t *lin_ind, const float *voxel_coords,
                                 int ld, int M, int F, int G, int n_frustum, float *g_grid)
{
    const size_t G3 = (size_t)G * G * G;
    memset(g_grid, 0, sizeof(float) * (size_t)F * G3);
    for (int m = 0; m < M; ++m) {
        const float vc[3] = {voxel_coords[m], voxel_coords[ld + m], voxel_coords[2 * ld + m]};
        orc_taps t;
        orc_dv_taps(vc, G, &t);
        ORC_CORNERS(t)
        for (int f = 0; f < F; ++f) {
            const float g = g_out[(size_t)f * n_frustum + lin_ind[m]];
            float *gg = g_grid + f * G3;
            for (int k = 0; k < 8; ++k)
                gg[((size_t)cx_[k] * G + cy_[k]) * G + cz_[k]] += ((g * az_[k]) * ay_[k]) * ax_[k];
        }
    }
    return 0;
}

/* fused convenience: compute_proj_idcs + interpolate_trilinear (+ backward) for a batch.
 * grid (B,F,G,G,G), cam2world (B,16), frustum (B,F,D,H,W).  Samples with no element in
 * bounds produce zeros (the reference returns None and the caller would fail). */
ORC_API int orc_dv_project_fwd(const orc_dv_params *P, const float *grid, const float *cam2world,
                               int B, int F, float *frustum)
{
    const int n = P->W * P->H * P->D;
    const size_t G3 = (size_t)P->G * P->G * P->G;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int32_t *li = (int32_t *)malloc(sizeof(int32_t) * n);
        float *vc = (float *)malloc(sizeof(float) * 3 * (size_t)n);
        if (!li || !vc) { rc = -1; free(li); free(vc); continue; }
        const int M = orc_dv_compute_proj_idcs(P, cam2world + 16 * b, li, vc, n);
        orc_dv_trilinear_fwd(grid + (size_t)b * F * G3, li, vc, n, M, F, P->G, n, frustum + (size_t)b * F * n);
        free(li); free(vc);
    }
    return rc;
}

ORC_API int orc_dv_project_bwd(const orc_dv_params *P, const float *g_frustum, const float *cam2world,
                               int B, int F, float *g_grid)
{
    const int n = P->W * P->H * P->D;
    const size_t G3 = (size_t)P->G * P->G * P->G;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int32_t *li = (int32_t *)malloc(sizeof(int32_t) * n);
        float *vc = (float *)malloc(sizeof(float) * 3 * (size_t)n);
        if (!li || !vc) { rc = -1; free(li); free(vc); continue; }
        const int M = orc_dv_compute_proj_idcs(P, cam2world + 16 * b, li, vc, n);
        orc_dv_trilinear_bwd(g_frustum + (size_t)b * F * n, li, vc, n, M, F, P->G, n, g_grid + (size_t)b * F * G3);
        free(li); free(vc);
    }
    return rc;
}
