// deepvoxels.cu -- DeepVoxels frustum <-> voxel-grid projection sampling:
//   ProjectionHelper.compute_proj_idcs  (deepvoxel/projection.py:48-105)
//   interpolate_trilinear + its autograd (deepvoxel/deepvoxel.py:388-428)
// Per-element recipe: SURVEY.md Appendix A2 (fp32-pinned scalar semantics, quirks Q5-Q8, Q10).
#include <stdlib.h>

#include "common.cuh"

namespace rgbd {

struct Cam { float t[12]; };   // rows 0..2 of cam2world (row 3 is not needed: voxel_coords = grid_coords[:3])

__device__ __forceinline__ Cam load_cam(const float *__restrict__ cam2world)
{
    Cam c;
#pragma unroll
    for (int k = 0; k < 12; ++k) c.t[k] = __ldg(cam2world + k);
    return c;
}

// voxel coordinates of frustum element l; returns the in-bounds flag (projection.py:64-96)
// (d, tmp) = (l // (W*H), l - d*W*H) given directly
__device__ __forceinline__ bool dv_coords_at(const rgbd_dv_params &P, const Cam &T, int d, int tmp, float vc[3])
{
    // :67 true division (fractional row, Q5).  For a power-of-two width the float64 quotient is exactly tmp * 2^-k
    // (tmp < 2^24), so one exact fp32 multiply replaces the double division
    const bool pow2 = (P.W & (P.W - 1)) == 0 && P.W * P.H <= (1 << 24);
    const float yrow = pow2 ? __fmul_rn((float)tmp, 1.0f / (float)P.W) : (float)((double)tmp / (double)P.W);
    const float xcol = (float)(pow2 ? (tmp & (P.W - 1)) : (tmp % P.W));   // :68
    float zc = __fmul_rn((float)d, P.voxel_size);                     // :73
    zc = __fadd_rn(zc, P.near_plane);                                 // :74 (fp32, Q6)
    float xc = __fdiv_rn(__fsub_rn(xcol, P.cx), P.fx);                // :78
    float yc = __fdiv_rn(__fsub_rn(yrow, P.cy), P.fy);                // :79
    xc = __fmul_rn(xc, zc);                                           // :80
    yc = __fmul_rn(yc, zc);
    bool keep = true;
    const float half = (float)P.G / 2.0f, Gf = (float)P.G;
#pragma unroll
    for (int r = 0; r < 3; ++r) {                                     // xp.dot (sgemm, K=4): fma chain  :82
        float g = __fmul_rn(T.t[4 * r], xc);
        g = __fmaf_rn(T.t[4 * r + 1], yc, g);
        g = __fmaf_rn(T.t[4 * r + 2], zc, g);
        g = __fmaf_rn(T.t[4 * r + 3], 1.0f, g);
        const float v = __fadd_rn(__fdiv_rn(g, P.voxel_size), half);  // :87-88
        vc[r] = v;
        keep = keep && (v >= 0.0f) && (v < Gf);                       // :92-96
    }
    return keep;
}

__device__ __forceinline__ bool dv_coords(const rgbd_dv_params &P, const Cam &T, int l, float vc[3])
{
    const int WH = P.W * P.H;
    const int d = l / WH;                                             // :64
    return dv_coords_at(P, T, d, l - d * WH, vc);                     // :65-66
}

// compute_proj_idcs with the optional grid2world argument (projection.py:53-54, :83-84): grid_coords = world2grid .
// (cam2world . coords) -- two sgemm products with K = 4, all four rows of the first feed the second
struct Cam2 { float t[16]; float g[12]; };

__device__ __forceinline__ Cam2 load_cam2(const float *__restrict__ cam2world, const float *__restrict__ world2grid)
{
    Cam2 c;
#pragma unroll
    for (int k = 0; k < 16; ++k) c.t[k] = __ldg(cam2world + k);
#pragma unroll
    for (int k = 0; k < 12; ++k) c.g[k] = __ldg(world2grid + k);
    return c;
}

__device__ __forceinline__ bool dv_coords_g2w(const rgbd_dv_params &P, const Cam2 &T, int l, float vc[3])
{
    const int WH = P.W * P.H;
    const int d = l / WH, tmp = l - d * WH;
    const float yrow = (float)((double)tmp / (double)P.W);
    const float xcol = (float)(tmp % P.W);
    float zc = __fmul_rn((float)d, P.voxel_size);
    zc = __fadd_rn(zc, P.near_plane);
    float xc = __fdiv_rn(__fsub_rn(xcol, P.cx), P.fx);
    float yc = __fdiv_rn(__fsub_rn(yrow, P.cy), P.fy);
    xc = __fmul_rn(xc, zc);
    yc = __fmul_rn(yc, zc);
    float gc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float g = __fmul_rn(T.t[4 * r], xc);
        g = __fmaf_rn(T.t[4 * r + 1], yc, g);
        g = __fmaf_rn(T.t[4 * r + 2], zc, g);
        gc[r] = __fmaf_rn(T.t[4 * r + 3], 1.0f, g);
    }
    bool keep = true;
    const float half = (float)P.G / 2.0f, Gf = (float)P.G;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float g = __fmul_rn(T.g[4 * r], gc[0]);
        g = __fmaf_rn(T.g[4 * r + 1], gc[1], g);
        g = __fmaf_rn(T.g[4 * r + 2], gc[2], g);
        g = __fmaf_rn(T.g[4 * r + 3], gc[3], g);
        const float v = __fadd_rn(__fdiv_rn(g, P.voxel_size), half);
        vc[r] = v;
        keep = keep && (v >= 0.0f) && (v < Gf);
    }
    return keep;
}

struct Taps {
    int off[8];         // element offsets into one (G,G,G) feature brick, corner order of deepvoxel.py:416-423
    float ax[8], ay[8], az[8];
};

// deepvoxel.py:394-412: axis swap (grid axis 2 <- vc[2], axis 4 <- vc[0]), truncation, clamp, fp64 fractions
__device__ __forceinline__ void dv_taps(const float vc[3], int G, Taps &t)
{
    const float X = vc[2], Y = vc[1], Z = vc[0];
    const int x0 = __float2int_rz(X), y0 = __float2int_rz(Y), z0 = __float2int_rz(Z);
    const int x1 = min(max(x0 + 1, 0), G - 1), y1 = min(max(y0 + 1, 0), G - 1), z1 = min(max(z0 + 1, 0), G - 1);
    const double fx = (double)X - (double)x0, fy = (double)Y - (double)y0, fz = (double)Z - (double)z0;
    const float wx1 = (float)fx, wx0 = (float)(1.0 - fx);
    const float wy1 = (float)fy, wy0 = (float)(1.0 - fy);
    const float wz1 = (float)fz, wz0 = (float)(1.0 - fz);
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1}, zs[2] = {z0, z1};
    const float wxs[2] = {wx0, wx1}, wys[2] = {wy0, wy1}, wzs[2] = {wz0, wz1};
    // corner order: (0,0,0)(1,0,0)(0,1,0)(0,0,1)(1,0,1)(0,1,1)(1,1,0)(1,1,1)
    const int ox[8] = {0, 1, 0, 0, 1, 0, 1, 1}, oy[8] = {0, 0, 1, 0, 0, 1, 1, 1}, oz[8] = {0, 0, 0, 1, 1, 1, 0, 1};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        t.off[k] = (xs[ox[k]] * G + ys[oy[k]]) * G + zs[oz[k]];
        t.ax[k] = wxs[ox[k]]; t.ay[k] = wys[oy[k]]; t.az[k] = wzs[oz[k]];
    }
}

__device__ __forceinline__ float dv_interp(const float *__restrict__ brick, const Taps &t)
{
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float term = __fmul_rn(__fmul_rn(__fmul_rn(__ldg(brick + t.off[k]), t.ax[k]), t.ay[k]), t.az[k]);
        acc = (k == 0) ? term : __fadd_rn(acc, term);
    }
    return acc;
}

// ---------------------------------------------------------------- fused batch path (no index lists)
__global__ void __launch_bounds__(kThreads)
k_dv_project_fwd(const rgbd_dv_params P, const float *__restrict__ grid, const float *__restrict__ cam2world,
                 int F, float *__restrict__ frustum)
{
    const int n = P.W * P.H * P.D;
    const int l = blockIdx.x * kThreads + threadIdx.x;
    const int b = blockIdx.y;
    if (l >= n) return;
    const Cam T = load_cam(cam2world + 16 * b);
    float vc[3];
    const bool keep = dv_coords(P, T, l, vc);
    const size_t G3 = (size_t)P.G * P.G * P.G;
    float *out = frustum + (size_t)b * F * n + l;
    if (!keep) {
        for (int f = 0; f < F; ++f) out[(size_t)f * n] = 0.0f;
        return;
    }
    Taps t;
    dv_taps(vc, P.G, t);
    const float *g = grid + (size_t)b * F * G3;
    for (int f = 0; f < F; ++f) out[(size_t)f * n] = dv_interp(g + (size_t)f * G3, t);
}

__global__ void __launch_bounds__(kThreads)
k_dv_project_bwd(const rgbd_dv_params P, const float *__restrict__ g_frustum, const float *__restrict__ cam2world,
                 int F, float *__restrict__ g_grid)
{
    const int n = P.W * P.H * P.D;
    const int l = blockIdx.x * kThreads + threadIdx.x;
    const int b = blockIdx.y;
    if (l >= n) return;
    const Cam T = load_cam(cam2world + 16 * b);
    float vc[3];
    if (!dv_coords(P, T, l, vc)) return;
    Taps t;
    dv_taps(vc, P.G, t);
    const size_t G3 = (size_t)P.G * P.G * P.G;
    const float *go = g_frustum + (size_t)b * F * n + l;
    float *gg = g_grid + (size_t)b * F * G3;
    for (int f = 0; f < F; ++f) {
        const float g = __ldg(go + (size_t)f * n);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            atomicAdd(gg + (size_t)f * G3 + t.off[k], ((g * t.az[k]) * t.ay[k]) * t.ax[k]);
    }
}

// ---------------------------------------------------------------- channels-last fast path (fused batch entry points)
// The planar (B,F,G,G,G) grid makes every tap a 4-byte gather from F different planes: 8*F scattered
// 4-byte loads (or REDs) per frustum element, ~3.5 L1 wavefronts each.  Staging the grid of a chunk of
// samples as (B,G^3,F) in L2 turns one tap of one element into ONE 128-byte line that a warp reads (or
// REDs) with lane = feature: 8 wavefronts per element for all 32 features.  The frustum side stays in the
// reference's planar (F, D*H*W) layout; a 32x32 shared-memory tile per warp transposes between the two.

// (B,F,G3) <-> (B,G3,F), 32x32 tiles through shared memory; block (32,8)
__global__ void __launch_bounds__(256)
k_dv_to_cl(const float *__restrict__ grid, float *__restrict__ cl, int F, int G3)
{
    __shared__ float t[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
    const int v0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int f = f0 + ty + j, v = v0 + tx;
        t[ty + j][tx] = (f < F && v < G3) ? __ldg(grid + ((size_t)b * F + f) * G3 + v) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int v = v0 + ty + j, f = f0 + tx;
        if (f < F && v < G3) cl[((size_t)b * G3 + v) * F + f] = t[tx][ty + j];
    }
}

__global__ void __launch_bounds__(256)
k_dv_from_cl(const float *__restrict__ cl, float *__restrict__ grid, int F, int G3)
{
    __shared__ float t[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
    const int v0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int v = v0 + ty + j, f = f0 + tx;
        t[ty + j][tx] = (f < F && v < G3) ? cl[((size_t)b * G3 + v) * F + f] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int f = f0 + ty + j, v = v0 + tx;
        if (f < F && v < G3) grid[((size_t)b * F + f) * G3 + v] = t[tx][ty + j];
    }
}

struct ElemTaps {            // what one lane computes for its own element and then broadcasts
    int off[8];              // voxel index * F of the 8 corners, corner order of deepvoxel.py:416-423
    float wx0, wx1, wy0, wy1, wz0, wz1;
};

__device__ __forceinline__ void dv_elem_taps(const float vc[3], int G, int F, ElemTaps &t)
{
    const float X = vc[2], Y = vc[1], Z = vc[0];                 // axis swap (deepvoxel.py:394-396)
    const int x0 = __float2int_rz(X), y0 = __float2int_rz(Y), z0 = __float2int_rz(Z);
    const int x1 = min(max(x0 + 1, 0), G - 1), y1 = min(max(y0 + 1, 0), G - 1), z1 = min(max(z0 + 1, 0), G - 1);
    const double fx = (double)X - (double)x0, fy = (double)Y - (double)y0, fz = (double)Z - (double)z0;
    t.wx1 = (float)fx; t.wx0 = (float)(1.0 - fx);
    t.wy1 = (float)fy; t.wy0 = (float)(1.0 - fy);
    t.wz1 = (float)fz; t.wz0 = (float)(1.0 - fz);
    // (0,0,0)(1,0,0)(0,1,0)(0,0,1)(1,0,1)(0,1,1)(1,1,0)(1,1,1)
    t.off[0] = (x0 * G + y0) * G + z0; t.off[1] = (x1 * G + y0) * G + z0;
    t.off[2] = (x0 * G + y1) * G + z0; t.off[3] = (x0 * G + y0) * G + z1;
    t.off[4] = (x1 * G + y0) * G + z1; t.off[5] = (x0 * G + y1) * G + z1;
    t.off[6] = (x1 * G + y1) * G + z0; t.off[7] = (x1 * G + y1) * G + z1;
#pragma unroll
    for (int k = 0; k < 8; ++k) t.off[k] *= F;       // element offset of the corner's feature line in the chunk (< 2^31)
}

constexpr int kDvWarps = 8;
constexpr int kDvTileStride = 36;        // floats per tile row: 16-byte aligned rows for the float4 accesses

// One warp = 32 consecutive frustum elements.  While gathering, a lane is (element slot eg = lane/8,
// feature quad fq = lane%8): four kept elements are processed per iteration, each 8-lane group reading the
// 128-byte feature line of one corner with 16-byte loads.  While storing, lane = element (coalesced along l).
// EXACT: every term is ((v*wx)*wy)*wz summed left to right, as deepvoxel.py:416-423 evaluates it (bit-equal to the
// reference's fp32 elementwise chain; 31 FP instructions per output).  !EXACT: the three factors of a corner are
// folded once per element and a feature costs 8 FMAs -- same indices and masks, values within ~1e-7 relative
// (the parity bar for floating-point outputs is 1e-5); this is the default, RGBD_B200_DV_EXACT=1 selects EXACT.
template <bool EXACT>
__global__ void __launch_bounds__(32 * kDvWarps)
k_dv_project_fwd_cl(const rgbd_dv_params P, const float *__restrict__ cl, const float *__restrict__ cam2world,
                    int F, float *__restrict__ frustum)
{
    __shared__ __align__(16) float tile[kDvWarps][32][kDvTileStride];
    __shared__ unsigned char elist[kDvWarps][32];                      // lanes of the kept elements, in order
    const int n = P.W * P.H * P.D;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int base = (blockIdx.x * kDvWarps + wid) * 32;
    if (base >= n) return;
    const int l = base + lane;
    const Cam T = load_cam(cam2world + 16 * b);
    float vc[3] = {0.f, 0.f, 0.f};
    const bool keep = (l < n) && dv_coords(P, T, l, vc);
    ElemTaps mine = {};
    if (keep) dv_elem_taps(vc, P.G, F, mine);
    const unsigned FULL = 0xffffffffu;
    const unsigned kept = __ballot_sync(FULL, keep);
    const int nk = __popc(kept);
    if (keep) elist[wid][__popc(kept & ((1u << lane) - 1u))] = (unsigned char)lane;
    __syncwarp();
    const int eg = lane >> 3, fq = lane & 7;
    const size_t G3 = (size_t)P.G * P.G * P.G;
    float (*tl)[kDvTileStride] = tile[wid];
    float wf[8];                                                       // !EXACT: folded corner weights of MY element
    wf[0] = (mine.wx0 * mine.wy0) * mine.wz0; wf[1] = (mine.wx1 * mine.wy0) * mine.wz0;
    wf[2] = (mine.wx0 * mine.wy1) * mine.wz0; wf[3] = (mine.wx0 * mine.wy0) * mine.wz1;
    wf[4] = (mine.wx1 * mine.wy0) * mine.wz1; wf[5] = (mine.wx0 * mine.wy1) * mine.wz1;
    wf[6] = (mine.wx1 * mine.wy1) * mine.wz0; wf[7] = (mine.wx1 * mine.wy1) * mine.wz1;
    for (int f0 = 0; f0 < F; f0 += 32) {
        const float *__restrict__ src = cl + (size_t)b * G3 * F;
        const int f = f0 + 4 * fq;
#pragma unroll 4
        for (int e = 0; e < 32; ++e) tl[e][lane] = 0.0f;               // elements outside the grid stay 0
        __syncwarp();
        for (int it = 0; it < nk; it += 4) {                           // warp-uniform trip count
            const int idx = it + eg;
            const bool active = idx < nk;
            const int e = active ? (int)elist[wid][idx] : 0;           // lane that owns the idx-th kept element
            int off[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) off[k] = __shfl_sync(FULL, mine.off[k], e);
            if (!EXACT) {
                float w[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = __shfl_sync(FULL, wf[k], e);
                if (active && f < F) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (off[k] + f)));
                        acc.x = fmaf(v.x, w[k], acc.x); acc.y = fmaf(v.y, w[k], acc.y);
                        acc.z = fmaf(v.z, w[k], acc.z); acc.w = fmaf(v.w, w[k], acc.w);
                    }
                    *reinterpret_cast<float4 *>(&tl[e][4 * fq]) = acc;
                }
                continue;
            }
            const float wx0 = __shfl_sync(FULL, mine.wx0, e), wx1 = __shfl_sync(FULL, mine.wx1, e);
            const float wy0 = __shfl_sync(FULL, mine.wy0, e), wy1 = __shfl_sync(FULL, mine.wy1, e);
            const float wz0 = __shfl_sync(FULL, mine.wz0, e), wz1 = __shfl_sync(FULL, mine.wz1, e);
            if (active && f < F) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4 *>(src + (off[k] + f)));
                // each term ((v*wx)*wy)*wz, summed left to right in the corner order of deepvoxel.py:416-423
#define RGBD_TERM(K_, C_, WX_, WY_, WZ_) __fmul_rn(__fmul_rn(__fmul_rn(v[K_].C_, WX_), WY_), WZ_)
#define RGBD_SUM(C_)                                                                                   \
    __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(RGBD_TERM(0, C_, wx0, wy0, wz0), \
        RGBD_TERM(1, C_, wx1, wy0, wz0)), RGBD_TERM(2, C_, wx0, wy1, wz0)), RGBD_TERM(3, C_, wx0, wy0, wz1)), \
        RGBD_TERM(4, C_, wx1, wy0, wz1)), RGBD_TERM(5, C_, wx0, wy1, wz1)), RGBD_TERM(6, C_, wx1, wy1, wz0)), \
        RGBD_TERM(7, C_, wx1, wy1, wz1))
                const float4 acc = make_float4(RGBD_SUM(x), RGBD_SUM(y), RGBD_SUM(z), RGBD_SUM(w));
#undef RGBD_SUM
#undef RGBD_TERM
                *reinterpret_cast<float4 *>(&tl[e][4 * fq]) = acc;     // row = element, columns = features
            }
        }
        __syncwarp();
        if (l < n) {
            const int fmax = min(32, F - f0);
            float *out = frustum + ((size_t)b * F + f0) * n + l;
            for (int ff = 0; ff < fmax; ++ff) out[(size_t)ff * n] = tl[lane][ff];   // coalesced along l
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * kDvWarps)
k_dv_project_bwd_cl(const rgbd_dv_params P, const float *__restrict__ g_frustum, const float *__restrict__ cam2world,
                    int F, float *__restrict__ gcl)
{
    __shared__ __align__(16) float tile[kDvWarps][32][kDvTileStride];
    __shared__ unsigned char elist[kDvWarps][32];
    const int n = P.W * P.H * P.D;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int base = (blockIdx.x * kDvWarps + wid) * 32;
    if (base >= n) return;
    const int l = base + lane;
    const Cam T = load_cam(cam2world + 16 * b);
    float vc[3] = {0.f, 0.f, 0.f};
    const bool keep = (l < n) && dv_coords(P, T, l, vc);
    ElemTaps mine = {};
    if (keep) dv_elem_taps(vc, P.G, F, mine);
    const unsigned FULL = 0xffffffffu;
    const unsigned kept = __ballot_sync(FULL, keep);
    if (kept == 0u) return;
    const int nk = __popc(kept);
    if (keep) elist[wid][__popc(kept & ((1u << lane) - 1u))] = (unsigned char)lane;
    // the lift only needs 1e-5: fold the three factors of every corner once per element (owner lane)
    float w[8];
    w[0] = (mine.wz0 * mine.wy0) * mine.wx0; w[1] = (mine.wz0 * mine.wy0) * mine.wx1;
    w[2] = (mine.wz0 * mine.wy1) * mine.wx0; w[3] = (mine.wz1 * mine.wy0) * mine.wx0;
    w[4] = (mine.wz1 * mine.wy0) * mine.wx1; w[5] = (mine.wz1 * mine.wy1) * mine.wx0;
    w[6] = (mine.wz0 * mine.wy1) * mine.wx1; w[7] = (mine.wz1 * mine.wy1) * mine.wx1;
    const int eg = lane >> 3, fq = lane & 7;
    // each 8-lane group walks a CONTIGUOUS quarter of the kept elements: neighbouring elements of an image row
    // usually fall into the same voxel cell, so their contributions to a corner are summed in registers and
    // leave as ONE 128-byte RED per run (the L2 atomic traffic, 8 lines per element, was the limiter)
    const int q = (nk + 3) >> 2;
    const int lo = eg * q, hi = min(nk, lo + q);
    const size_t G3 = (size_t)P.G * P.G * P.G;
    float (*tl)[kDvTileStride] = tile[wid];
    for (int f0 = 0; f0 < F; f0 += 32) {
        const int fmax = min(32, F - f0);
        if (l < n) {
            const float *go = g_frustum + ((size_t)b * F + f0) * n + l;
            for (int ff = 0; ff < fmax; ++ff) tl[lane][ff] = __ldg(go + (size_t)ff * n);   // coalesced along l
        }
        __syncwarp();
        float *__restrict__ dst = gcl + (size_t)b * G3 * F;
        const int f = f0 + 4 * fq;
        const bool fok = f < F;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int prev = -1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int it = 0; it < q; ++it) {                           // warp-uniform trip count
                const int idx = lo + it;
                const bool active = idx < hi;
                const int e = active ? (int)elist[wid][idx] : 0;
                const int off = __shfl_sync(FULL, mine.off[k], e);
                const float wk = __shfl_sync(FULL, w[k], e);
                if (active && fok) {
                    const float4 g = *reinterpret_cast<const float4 *>(&tl[e][4 * fq]);
                    if (off != prev) {
                        if (prev >= 0) atomicAdd(reinterpret_cast<float4 *>(dst + (prev + f)), acc);
                        prev = off;
                        acc = make_float4(g.x * wk, g.y * wk, g.z * wk, g.w * wk);
                    } else {
                        acc.x += g.x * wk; acc.y += g.y * wk; acc.z += g.z * wk; acc.w += g.w * wk;
                    }
                }
            }
            if (prev >= 0) atomicAdd(reinterpret_cast<float4 *>(dst + (prev + f)), acc);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- explicit index-list path
__global__ void __launch_bounds__(kThreads)
k_dv_trilinear_fwd(const float *__restrict__ grid, const int32_t *__restrict__ lin_ind,
                   const float *__restrict__ voxel_coords, int ld, int M, int F, int G, int n,
                   float *__restrict__ frustum)
{
    const int m = blockIdx.x * kThreads + threadIdx.x;
    if (m >= M) return;
    const float vc[3] = {voxel_coords[m], voxel_coords[ld + m], voxel_coords[2 * (size_t)ld + m]};
    Taps t;
    dv_taps(vc, G, t);
    const size_t G3 = (size_t)G * G * G;
    float *out = frustum + lin_ind[m];
    for (int f = 0; f < F; ++f) out[(size_t)f * n] = dv_interp(grid + (size_t)f * G3, t);
}

__global__ void __launch_bounds__(kThreads)
k_dv_trilinear_bwd(const float *__restrict__ g_frustum, const int32_t *__restrict__ lin_ind,
                   const float *__restrict__ voxel_coords, int ld, int M, int F, int G, int n,
                   float *__restrict__ g_grid)
{
    const int m = blockIdx.x * kThreads + threadIdx.x;
    if (m >= M) return;
    const float vc[3] = {voxel_coords[m], voxel_coords[ld + m], voxel_coords[2 * (size_t)ld + m]};
    Taps t;
    dv_taps(vc, G, t);
    const size_t G3 = (size_t)G * G * G;
    const float *go = g_frustum + lin_ind[m];
    for (int f = 0; f < F; ++f) {
        const float g = __ldg(go + (size_t)f * n);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            atomicAdd(g_grid + (size_t)f * G3 + t.off[k], ((g * t.az[k]) * t.ay[k]) * t.ax[k]);
    }
}

// ---------------------------------------------------------------- compute_proj_idcs (ordered compaction)
__global__ void __launch_bounds__(kThreads)
k_dv_count(const rgbd_dv_params P, const float *__restrict__ cam2world, const float *__restrict__ world2grid,
           int *__restrict__ block_counts)
{
    const int n = P.W * P.H * P.D;
    const int l = blockIdx.x * kThreads + threadIdx.x;
    float vc[3];
    bool keep;
    if (world2grid) keep = (l < n) && dv_coords_g2w(P, load_cam2(cam2world, world2grid), l, vc);
    else keep = (l < n) && dv_coords(P, load_cam(cam2world), l, vc);
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// exclusive scan of the block counts by one block (nblocks is a few thousand at most)
__global__ void __launch_bounds__(1024)
k_dv_scan(const int *__restrict__ block_counts, int nblocks, int *__restrict__ block_offsets, int *__restrict__ total)
{
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < nblocks ? block_counts[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int s = 1; s < 1024; s <<= 1) {                         // Hillis-Steele inclusive scan
            const int add = threadIdx.x >= s ? sh[threadIdx.x - s] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (idx < nblocks) block_offsets[idx] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kThreads)
k_dv_compact(const rgbd_dv_params P, const float *__restrict__ cam2world, const float *__restrict__ world2grid,
             const int *__restrict__ block_offsets, int32_t *__restrict__ lin_ind, float *__restrict__ voxel_coords, int ld)
{
    __shared__ int warp_base[kThreads / 32];
    const int n = P.W * P.H * P.D;
    const int l = blockIdx.x * kThreads + threadIdx.x;
    float vc[3];
    bool keep;
    if (world2grid) keep = (l < n) && dv_coords_g2w(P, load_cam2(cam2world, world2grid), l, vc);
    else keep = (l < n) && dv_coords(P, load_cam(cam2world), l, vc);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_base[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < kThreads / 32; ++w) { const int v = warp_base[w]; warp_base[w] = acc; acc += v; }
    }
    __syncthreads();
    if (keep) {
        const int pos = block_offsets[blockIdx.x] + warp_base[wid] + __popc(bal & ((1u << lane) - 1u));
        lin_ind[pos] = l;
        voxel_coords[pos] = vc[0];
        voxel_coords[(size_t)ld + pos] = vc[1];
        voxel_coords[2 * (size_t)ld + pos] = vc[2];
    }
}

// ---------------------------------------------------------------- fused render tail ("next" row)
// DeepVoxels.forward with `occlusion_type: accumulative` (deepvoxels_shapenet_car.yml:34), per sample:
//     vol = interpolate_trilinear(...)                                         deepvoxel.py:880-884
//     occ = sigmoid(conv1x1(leaky_relu(conv1x1(concat(depth_coords, vol)))) - threshold)   :560-567,:575-576 (per voxel MLP, F+1 -> 4 -> 1)
//     c = cumsum_d(occ); w_d = clip(c_d,0,1) - clip(c_{d-1},0,1)               :579-582
//     depth = sum_d depth_coords_d * w_d;  novel = sum_d w_d * vol_d;  fg = sum_d w_d         :583,:888,:892
//     depth = (depth + 0.5) * D * voxel_size + near_plane                       :903-904
// The reference materialises vol (B,F,D,H,W: 28-58 MB per sample) and a dozen temporaries of the same size; here a
// ray is walked front to back by an 8-lane group (lane = feature quad) and nothing but the (F+2) output planes
// is written.  A ray stops at the first depth with c > 1: every later weight is exactly 0 and the clip passes
// no gradient there.  Coordinates and in-grid mask are the bit-exact recipe of compute_proj_idcs (dv_coords);
// values are tolerance-checked (1e-5) against golden vectors produced by the reference's own forward().
constexpr int kRenderMaxD = 128;
constexpr int kRenderNf = 4;             // occnet_nf (deepvoxel.py:830)

struct RenderW {                          // the occlusion MLP as seen by one lane (feature quad fq)
    float w1q[kRenderNf][4];              // W1[j][1 + 4*fq + c]
    float w1d[kRenderNf], b1[kRenderNf], w2[kRenderNf];   // W1[j][0], b1[j], W2[0][j]
    float b2;
};

__device__ __forceinline__ RenderW render_load_w(const float *__restrict__ W1, const float *__restrict__ b1,
                                                 const float *__restrict__ W2, const float *__restrict__ b2, int F, int fq)
{
    RenderW r;
#pragma unroll
    for (int j = 0; j < kRenderNf; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) r.w1q[j][c] = (4 * fq + c < F) ? __ldg(W1 + j * (F + 1) + 1 + 4 * fq + c) : 0.0f;
        r.w1d[j] = __ldg(W1 + j * (F + 1)); r.b1[j] = __ldg(b1 + j); r.w2[j] = __ldg(W2 + j);
    }
    r.b2 = __ldg(b2);
    return r;
}

// sum over the 8 lanes of a feature-quad group (all 8 lanes get the result)
__device__ __forceinline__ float group8_sum(float v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// occlusion MLP of one voxel: a_j (pre-activation) and occ; x = inv_c1 * (depth_coord | features)   pggan.py:38
__device__ __forceinline__ float render_mlp(const RenderW &w, const rgbd_dv_render_params &R, float4 feat, float dc,
                                            float a[kRenderNf])
{
    const float x0 = R.inv_c1 * feat.x, x1 = R.inv_c1 * feat.y, x2 = R.inv_c1 * feat.z, x3 = R.inv_c1 * feat.w;
    const float xd = R.inv_c1 * dc;
    float s = w.b2;
#pragma unroll
    for (int j = 0; j < kRenderNf; ++j) {
        float p = fmaf(w.w1q[j][3], x3, fmaf(w.w1q[j][2], x2, fmaf(w.w1q[j][1], x1, w.w1q[j][0] * x0)));
        p = group8_sum(p);
        const float aj = (p + w.w1d[j] * xd) + w.b1[j];
        a[j] = aj;
        const float h = aj < 0.0f ? 0.2f * aj : aj;                      // F.leaky_relu (slope 0.2)
        s = fmaf(w.w2[j], R.inv_c2 * h, s);
    }
    s -= R.threshold;                                                    // :565
    return tanhf(s * 0.5f) * 0.5f + 0.5f;                                // F.sigmoid as Chainer evaluates it
}

__device__ __forceinline__ float render_depth_coord(int d, int D)
{
    // np.arange(-D // 2, D // 2)[d] / D  (deepvoxel.py:568-569; -D // 2 is a floor division)
    const int lo = -((D + 1) / 2);
    return (float)((double)(lo + d) / (double)D);
}

// forward-only variant: the four hidden pre-activations are reduced TRANSPOSED over the 8 lanes (2 + 1 + 1 shuffles
// leave lane (b2 b1 b0) with the complete sum of unit j = 2*b2 + b1), the output layer then needs 2 more shuffles:
// 6 instead of 12, and every lane evaluates one hidden unit instead of four
struct RenderWm { float w1d, b1, w2; };       // the constants of "my" hidden unit
__device__ __forceinline__ RenderWm render_my_unit(const RenderW &w, int fq)
{
    const int j = ((fq >> 2) & 1) * 2 + ((fq >> 1) & 1);
    RenderWm m;
    m.w1d = j == 0 ? w.w1d[0] : (j == 1 ? w.w1d[1] : (j == 2 ? w.w1d[2] : w.w1d[3]));
    m.b1 = j == 0 ? w.b1[0] : (j == 1 ? w.b1[1] : (j == 2 ? w.b1[2] : w.b1[3]));
    m.w2 = j == 0 ? w.w2[0] : (j == 1 ? w.w2[1] : (j == 2 ? w.w2[2] : w.w2[3]));
    return m;
}
__device__ __forceinline__ float render_mlp_fwd(const RenderW &w, const RenderWm &m, const rgbd_dv_render_params &R,
                                                float4 feat, float dc, int fq)
{
    const unsigned FULL = 0xffffffffu;
    const float x0 = R.inv_c1 * feat.x, x1 = R.inv_c1 * feat.y, x2 = R.inv_c1 * feat.z, x3 = R.inv_c1 * feat.w;
    float p[kRenderNf];
#pragma unroll
    for (int j = 0; j < kRenderNf; ++j)
        p[j] = fmaf(w.w1q[j][3], x3, fmaf(w.w1q[j][2], x2, fmaf(w.w1q[j][1], x1, w.w1q[j][0] * x0)));
    const bool b2 = fq & 4, b1 = fq & 2;
    const float r0 = __shfl_xor_sync(FULL, b2 ? p[0] : p[2], 4), r1 = __shfl_xor_sync(FULL, b2 ? p[1] : p[3], 4);
    const float q0 = (b2 ? p[2] : p[0]) + r0, q1 = (b2 ? p[3] : p[1]) + r1;
    float t = (b1 ? q1 : q0) + __shfl_xor_sync(FULL, b1 ? q0 : q1, 2);
    t += __shfl_xor_sync(FULL, t, 1);
    const float aj = (t + m.w1d * (R.inv_c1 * dc)) + m.b1;
    const float h = aj < 0.0f ? 0.2f * aj : aj;
    float s = m.w2 * (R.inv_c2 * h);
    s += __shfl_xor_sync(FULL, s, 2);
    s += __shfl_xor_sync(FULL, s, 4);
    s = (s + w.b2) - R.threshold;
    return tanhf(s * 0.5f) * 0.5f + 0.5f;
}

struct RayTaps { int off[8]; float w[8]; int keep; float dc; };

// taps of frustum element (d, row, col) for the lane that owns depth slot `d`; folded corner weights
__device__ __forceinline__ void render_taps(const rgbd_dv_params &P, const Cam &T, int d, int pix, bool valid, int F,
                                            RayTaps &t)
{
    float vc[3] = {0.f, 0.f, 0.f};
    const bool keep = valid && d < P.D && dv_coords_at(P, T, d, pix, vc);
    t.keep = keep ? 1 : 0;
    t.dc = render_depth_coord(d, P.D);          // (a double division: once per chunk and lane, then shuffled)
    if (keep) {
        ElemTaps e;
        dv_elem_taps(vc, P.G, F, e);
#pragma unroll
        for (int k = 0; k < 8; ++k) t.off[k] = e.off[k];
        t.w[0] = (e.wx0 * e.wy0) * e.wz0; t.w[1] = (e.wx1 * e.wy0) * e.wz0; t.w[2] = (e.wx0 * e.wy1) * e.wz0;
        t.w[3] = (e.wx0 * e.wy0) * e.wz1; t.w[4] = (e.wx1 * e.wy0) * e.wz1; t.w[5] = (e.wx0 * e.wy1) * e.wz1;
        t.w[6] = (e.wx1 * e.wy1) * e.wz0; t.w[7] = (e.wx1 * e.wy1) * e.wz1;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) { t.off[k] = 0; t.w[k] = 0.0f; }
    }
}

#ifndef RGBD_RENDER_MINBLK
#define RGBD_RENDER_MINBLK 3
#endif
// one warp = 4 rays x 8 feature quads; block = 8 warps = 32 consecutive pixels
__global__ void __launch_bounds__(32 * kDvWarps, RGBD_RENDER_MINBLK)
k_dv_render_fwd(const rgbd_dv_params P, const rgbd_dv_render_params R, const float *__restrict__ cl,
                const float *__restrict__ cam2world, const float *__restrict__ W1, const float *__restrict__ b1,
                const float *__restrict__ W2, const float *__restrict__ b2, int F, float *__restrict__ novel,
                float *__restrict__ depth, float *__restrict__ fg, float *__restrict__ saved)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int eg = lane >> 3, fq = lane & 7, gbase = lane & ~7;
    const int b = blockIdx.y;
    const int HW = P.W * P.H;
    const int pix = (blockIdx.x * kDvWarps + wid) * 4 + eg;
    const bool valid = pix < HW;
    const Cam T = load_cam(cam2world + 16 * b);
    const RenderW w = render_load_w(W1, b1, W2, b2, F, fq);
    const RenderWm wm = render_my_unit(w, fq);
    const size_t G3 = (size_t)P.G * P.G * P.G;
    const float *__restrict__ src = cl + (size_t)b * G3 * F;
    const int f = 4 * fq;
    float c = 0.0f, clip_prev = 0.0f, dm = 0.0f, fgs = 0.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool done = !valid;
    int nd = 0;
    // state for the backward pass (optional): per ray [nd | c_0 .. c_{D-1}], 4 B per depth step instead of the
    // reference's F*4 B view volume; lets the backward skip its front-to-back recomputation pass
    float *sv = saved ? saved + ((size_t)b * HW + (valid ? pix : 0)) * (P.D + 1) : nullptr;
    // (gathering the 8 steps of a chunk together before the MLPs and the scan was measured: no gain, the kernel
    //  is bound by shuffle / issue throughput at 128 registers, not by the latency of the gathers)
    for (int d0 = 0; d0 < P.D; d0 += 8) {
        if (__all_sync(FULL, done)) break;                              // every ray of the warp is saturated
        RayTaps mine;
        render_taps(P, T, d0 + fq, pix, valid && !done, F, mine);       // lane fq prepares depth d0 + fq of its ray
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const int d = d0 + q;
            if (d >= P.D) break;                                        // warp-uniform
            const int sl = gbase | q;
            const int keep = __shfl_sync(FULL, mine.keep, sl);
            int off[8];
            float wk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { off[k] = __shfl_sync(FULL, mine.off[k], sl); wk[k] = __shfl_sync(FULL, mine.w[k], sl); }
            float4 feat = make_float4(0.f, 0.f, 0.f, 0.f);
            if (keep && !done && f < F) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (off[k] + f)));
                    feat.x = fmaf(v.x, wk[k], feat.x); feat.y = fmaf(v.y, wk[k], feat.y);
                    feat.z = fmaf(v.z, wk[k], feat.z); feat.w = fmaf(v.w, wk[k], feat.w);
                }
            }
            const float dc = __shfl_sync(FULL, mine.dc, sl);
            const float occ = render_mlp_fwd(w, wm, R, feat, dc, fq);   // (shuffles: executed by every lane)
            if (!done) {
                c += occ;                                               // F.cumsum
                const float clip = fminf(fmaxf(c, 0.0f), 1.0f);         // F.clip(., 0, 1)
                const float wd = clip - clip_prev;                      // cumsum[1:] - cumsum[:-1]
                clip_prev = clip;
                acc.x = fmaf(wd, feat.x, acc.x); acc.y = fmaf(wd, feat.y, acc.y);
                acc.z = fmaf(wd, feat.z, acc.z); acc.w = fmaf(wd, feat.w, acc.w);
                dm = fmaf(dc, wd, dm);
                fgs += wd;
                nd = d + 1;
                if (sv && fq == 0) sv[1 + d] = c;
                if (c > 1.0f) done = true;                              // later weights are exactly 0
            }
        }
    }
    if (sv && valid && fq == 0) sv[0] = __int_as_float(nd);
    if (valid) {
        float *o = novel + ((size_t)b * F + f) * HW + pix;
        if (f + 0 < F) o[0] = acc.x;
        if (f + 1 < F) o[(size_t)HW] = acc.y;
        if (f + 2 < F) o[2 * (size_t)HW] = acc.z;
        if (f + 3 < F) o[3 * (size_t)HW] = acc.w;
        if (fq == 0) {
            // ((depth + 0.5) * int(ceil(sqrt(3) G))) * voxel_size + near_plane, one fp32 rounding per step  :903-904
            depth[(size_t)b * HW + pix] = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(dm, 0.5f), (float)R.depth_steps), P.voxel_size), P.near_plane);
            if (fg) fg[(size_t)b * HW + pix] = fgs;
        }
    }
}

// Backward of the fused render tail by recomputation.  Pass 1 walks the ray front to back and keeps, per depth,
// the running sum c_d and dL/dw_d in shared memory (2 x 4 B per depth step and ray); pass 2 walks back to front
// with the suffix sum of dL/dc (backward of F.cumsum), re-gathers the features, back-propagates the occlusion
// MLP (weight gradients accumulate in registers, reduced once per block) and scatters dL/dfeature into the
// channels-last grid gradient with 16-byte REDs (the lift).  Depths behind the first c_d > 1 carry no gradient.
#ifndef RGBD_RENDER_BWD_MINBLK
#define RGBD_RENDER_BWD_MINBLK 2
#endif
__global__ void __launch_bounds__(32 * kDvWarps, RGBD_RENDER_BWD_MINBLK)
k_dv_render_bwd(const rgbd_dv_params P, const rgbd_dv_render_params R, const float *__restrict__ cl,
                const float *__restrict__ cam2world, const float *__restrict__ W1, const float *__restrict__ b1,
                const float *__restrict__ W2, const float *__restrict__ b2, int F, const float *__restrict__ g_novel,
                const float *__restrict__ g_depth, const float *__restrict__ g_fg, float *__restrict__ gcl,
                float *__restrict__ partials, int nvals, const float *__restrict__ saved)
{
    __shared__ float s_c[kDvWarps][4][kRenderMaxD];
    __shared__ float s_red[kDvWarps][kRenderNf * 33 + 2 * kRenderNf + 1];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int eg = lane >> 3, fq = lane & 7, gbase = lane & ~7;
    const int b = blockIdx.y;
    const int HW = P.W * P.H;
    const int pix = (blockIdx.x * kDvWarps + wid) * 4 + eg;
    const bool valid = pix < HW;
    const Cam T = load_cam(cam2world + 16 * b);
    const RenderW w = render_load_w(W1, b1, W2, b2, F, fq);
    const size_t G3 = (size_t)P.G * P.G * P.G;
    const float *__restrict__ src = cl + (size_t)b * G3 * F;
    float *__restrict__ dst = gcl + (size_t)b * G3 * F;
    const int f = 4 * fq;
    float4 gcol = make_float4(0.f, 0.f, 0.f, 0.f);
    float gdm = 0.0f, gfg = 0.0f;
    if (valid) {
        const float *g = g_novel + ((size_t)b * F + f) * HW + pix;
        if (f + 0 < F) gcol.x = __ldg(g);
        if (f + 1 < F) gcol.y = __ldg(g + (size_t)HW);
        if (f + 2 < F) gcol.z = __ldg(g + 2 * (size_t)HW);
        if (f + 3 < F) gcol.w = __ldg(g + 3 * (size_t)HW);
        gdm = (__ldg(g_depth + (size_t)b * HW + pix) * P.voxel_size) * (float)R.depth_steps;   // backward of :903-904
        if (g_fg) gfg = __ldg(g_fg + (size_t)b * HW + pix);
    }
    float *sc = s_c[wid][eg];
    int nd = 0;                                                          // depth steps that matter for this ray

    // ---- pass 1: front to back -- only when the forward did not leave its running sums (`saved`)
    if (saved) {
        const float *sv = saved + ((size_t)b * HW + (valid ? pix : 0)) * (P.D + 1);
        nd = valid ? __float_as_int(__ldg(sv)) : 0;
        for (int d = fq; d < nd; d += 8) sc[d] = __ldg(sv + 1 + d);
    } else {
        float c = 0.0f;
        bool done = !valid;
        for (int d0 = 0; d0 < P.D; d0 += 8) {
            if (__all_sync(FULL, done)) break;
            RayTaps mine;
            render_taps(P, T, d0 + fq, pix, valid && !done, F, mine);
#pragma unroll 1
            for (int q = 0; q < 8; ++q) {
                const int d = d0 + q;
                if (d >= P.D) break;
                const int sl = gbase | q;
                const int keep = __shfl_sync(FULL, mine.keep, sl);
                int off[8];
                float wk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { off[k] = __shfl_sync(FULL, mine.off[k], sl); wk[k] = __shfl_sync(FULL, mine.w[k], sl); }
                float4 feat = make_float4(0.f, 0.f, 0.f, 0.f);
                if (keep && !done && f < F) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (off[k] + f)));
                        feat.x = fmaf(v.x, wk[k], feat.x); feat.y = fmaf(v.y, wk[k], feat.y);
                        feat.z = fmaf(v.z, wk[k], feat.z); feat.w = fmaf(v.w, wk[k], feat.w);
                    }
                }
                const float dc = __shfl_sync(FULL, mine.dc, sl);
                float a[kRenderNf];
                const float occ = render_mlp(w, R, feat, dc, a);
                if (!done) {
                    c += occ;
                    if (fq == 0) sc[d] = c;
                    nd = d + 1;
                    if (c > 1.0f) done = true;
                }
            }
        }
    }
    __syncwarp();

    // ---- pass 2: back to front
    float gW1q[kRenderNf][4], gW1d[kRenderNf], gb1[kRenderNf], gW2[kRenderNf], gb2 = 0.0f;
#pragma unroll
    for (int j = 0; j < kRenderNf; ++j) {
        gW1d[j] = gb1[j] = gW2[j] = 0.0f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) gW1q[j][cc] = 0.0f;
    }
    int ndmax = nd;
    ndmax = max(ndmax, __shfl_xor_sync(FULL, ndmax, 8));
    ndmax = max(ndmax, __shfl_xor_sync(FULL, ndmax, 16));
    float S = 0.0f;                                                      // suffix sum of dL/dc (F.cumsum backward)
    float gw_next = 0.0f;                                                // dL/dw_{d+1} (0 behind the last step)
    for (int d0 = ((ndmax - 1) >> 3) << 3; d0 >= 0; d0 -= 8) {
        RayTaps mine;
        render_taps(P, T, d0 + fq, pix, valid && (d0 + fq) < nd, F, mine);
#pragma unroll 1
        for (int q = 7; q >= 0; --q) {
            const int d = d0 + q;
            if (d >= ndmax) continue;                                    // warp-uniform
            const bool active = valid && d < nd;
            const int sl = gbase | q;
            const int keep = __shfl_sync(FULL, mine.keep, sl);
            int off[8];
            float wk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { off[k] = __shfl_sync(FULL, mine.off[k], sl); wk[k] = __shfl_sync(FULL, mine.w[k], sl); }
            float4 ft = make_float4(0.f, 0.f, 0.f, 0.f);
            if (keep && active && f < F) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (off[k] + f)));
                    ft.x = fmaf(v.x, wk[k], ft.x); ft.y = fmaf(v.y, wk[k], ft.y);
                    ft.z = fmaf(v.z, wk[k], ft.z); ft.w = fmaf(v.w, wk[k], ft.w);
                }
            }
            const float dc = __shfl_sync(FULL, mine.dc, sl);
            float a[kRenderNf];
            const float occ = render_mlp(w, R, ft, dc, a);
            // dL/dw_d = <g_novel, feat_d> + g_depth_map * depth_coord_d + g_fg      (:583,:888,:892)
            const float dot = group8_sum(fmaf(gcol.w, ft.w, fmaf(gcol.z, ft.z, fmaf(gcol.y, ft.y, gcol.x * ft.x))));
            if (!active) continue;
            const float gw = fmaf(gdm, dc, dot) + gfg;
            const float cd = sc[d], cp = d > 0 ? sc[d - 1] : 0.0f;
            const float wd = fminf(fmaxf(cd, 0.0f), 1.0f) - fminf(fmaxf(cp, 0.0f), 1.0f);
            if (cd >= 0.0f && cd <= 1.0f) S += gw - gw_next;            // Clip backward (inclusive bounds), diff backward
            gw_next = gw;
            const float gs = S * occ * (1.0f - occ);                     // Sigmoid backward
            gb2 += gs;
            float gx0 = 0.0f, gx1 = 0.0f, gx2 = 0.0f, gx3 = 0.0f;
            const float x0 = R.inv_c1 * ft.x, x1 = R.inv_c1 * ft.y, x2 = R.inv_c1 * ft.z, x3 = R.inv_c1 * ft.w;
            const float xd = R.inv_c1 * dc;
#pragma unroll
            for (int j = 0; j < kRenderNf; ++j) {
                const float h = a[j] < 0.0f ? 0.2f * a[j] : a[j];
                gW2[j] = fmaf(gs, R.inv_c2 * h, gW2[j]);
                const float gh = (gs * w.w2[j]) * R.inv_c2;
                const float ga = a[j] < 0.0f ? 0.2f * gh : gh;           // LeakyReLU backward
                gb1[j] += ga;
                gW1d[j] = fmaf(ga, xd, gW1d[j]);
                gW1q[j][0] = fmaf(ga, x0, gW1q[j][0]); gW1q[j][1] = fmaf(ga, x1, gW1q[j][1]);
                gW1q[j][2] = fmaf(ga, x2, gW1q[j][2]); gW1q[j][3] = fmaf(ga, x3, gW1q[j][3]);
                gx0 = fmaf(ga, w.w1q[j][0], gx0); gx1 = fmaf(ga, w.w1q[j][1], gx1);
                gx2 = fmaf(ga, w.w1q[j][2], gx2); gx3 = fmaf(ga, w.w1q[j][3], gx3);
            }
            if (keep && f < F) {
                // dL/dfeat = w_d * g_novel + inv_c1 * W1^T ga, scattered to the 8 corners (interpolate_trilinear backward)
                const float4 gf = make_float4(fmaf(wd, gcol.x, R.inv_c1 * gx0), fmaf(wd, gcol.y, R.inv_c1 * gx1),
                                              fmaf(wd, gcol.z, R.inv_c1 * gx2), fmaf(wd, gcol.w, R.inv_c1 * gx3));
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    atomicAdd(reinterpret_cast<float4 *>(dst + (off[k] + f)),
                              make_float4(gf.x * wk[k], gf.y * wk[k], gf.z * wk[k], gf.w * wk[k]));
            }
        }
    }

    // ---- weight gradients: sum over the 4 rays of the warp, then over the warps of the block
    const int nW1 = kRenderNf * (F + 1);
#pragma unroll
    for (int j = 0; j < kRenderNf; ++j) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            float v = gW1q[j][cc];
            v += __shfl_xor_sync(FULL, v, 8); v += __shfl_xor_sync(FULL, v, 16);
            if (eg == 0 && f + cc < F) s_red[wid][j * (F + 1) + 1 + f + cc] = v;
        }
        float v0 = gW1d[j], v1 = gb1[j], v2 = gW2[j];
        v0 += __shfl_xor_sync(FULL, v0, 8); v0 += __shfl_xor_sync(FULL, v0, 16);
        v1 += __shfl_xor_sync(FULL, v1, 8); v1 += __shfl_xor_sync(FULL, v1, 16);
        v2 += __shfl_xor_sync(FULL, v2, 8); v2 += __shfl_xor_sync(FULL, v2, 16);
        if (lane == 0) { s_red[wid][j * (F + 1)] = v0; s_red[wid][nW1 + j] = v1; s_red[wid][nW1 + kRenderNf + j] = v2; }
    }
    {
        float v = gb2;
        v += __shfl_xor_sync(FULL, v, 8); v += __shfl_xor_sync(FULL, v, 16);
        if (lane == 0) s_red[wid][nW1 + 2 * kRenderNf] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < nvals) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < kDvWarps; ++k) v += s_red[k][threadIdx.x];
        partials[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * nvals + threadIdx.x] = v;
    }
}

// fixed-order sum of the per-block weight-gradient partials: one block per value
__global__ void __launch_bounds__(256)
k_dv_render_wgrad(const float *__restrict__ partials, int nblocks, int nvals, int F, float *__restrict__ g_W1,
                  float *__restrict__ g_b1, float *__restrict__ g_W2, float *__restrict__ g_b2)
{
    __shared__ double sh[256];
    const int v = blockIdx.x;
    double acc = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += 256) acc += (double)partials[(size_t)k * nvals + v];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) sh[threadIdx.x] += sh[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int nW1 = kRenderNf * (F + 1);
        const float r = (float)sh[0];
        if (v < nW1) g_W1[v] = r;
        else if (v < nW1 + kRenderNf) g_b1[v - nW1] = r;
        else if (v < nW1 + 2 * kRenderNf) g_W2[v - nW1 - kRenderNf] = r;
        else g_b2[0] = r;
    }
}

static bool dv_ok(const rgbd_dv_params *p)
{
    return p && p->W > 0 && p->H > 0 && p->D > 0 && p->G > 1 && p->voxel_size > 0.0f && p->fx != 0.0f && p->fy != 0.0f;
}

}  // namespace rgbd

using namespace rgbd;

extern "C" {

RGBD_API size_t rgbd_dv_workspace_bytes(const rgbd_dv_params *p)
{
    if (!dv_ok(p)) return 0;
    const size_t nblocks = ((size_t)p->W * p->H * p->D + kThreads - 1) / kThreads;
    return (2 * nblocks + 64) * sizeof(int);
}

static int dv_compute_proj_idcs_impl(const rgbd_dv_params *p, const float *cam2world, const float *world2grid,
                                     int32_t *lin_ind, float *voxel_coords, int *M_host, void *workspace,
                                     size_t workspace_bytes, void *stream);

RGBD_API int rgbd_dv_compute_proj_idcs(const rgbd_dv_params *p, const float *cam2world, int32_t *lin_ind,
                              float *voxel_coords, int *M_host, void *workspace, size_t workspace_bytes,
                              void *stream)
{
    return dv_compute_proj_idcs_impl(p, cam2world, nullptr, lin_ind, voxel_coords, M_host, workspace, workspace_bytes, stream);
}

RGBD_API int rgbd_dv_compute_proj_idcs_g2w(const rgbd_dv_params *p, const float *cam2world, const float *world2grid,
                                           int32_t *lin_ind, float *voxel_coords, int *M_host, void *workspace,
                                           size_t workspace_bytes, void *stream)
{
    if (!world2grid) { set_error("rgbd_dv_compute_proj_idcs_g2w: null world2grid"); return RGBD_E_ARG; }
    return dv_compute_proj_idcs_impl(p, cam2world, world2grid, lin_ind, voxel_coords, M_host, workspace, workspace_bytes, stream);
}

static int dv_compute_proj_idcs_impl(const rgbd_dv_params *p, const float *cam2world, const float *world2grid,
                                     int32_t *lin_ind, float *voxel_coords, int *M_host, void *workspace,
                                     size_t workspace_bytes, void *stream)
{
    if (!dv_ok(p) || !cam2world || !lin_ind || !voxel_coords || !M_host) {
        set_error("rgbd_dv_compute_proj_idcs: null pointer or bad params");
        return RGBD_E_ARG;
    }
    if (!workspace || workspace_bytes < rgbd_dv_workspace_bytes(p)) {
        set_error("rgbd_dv_compute_proj_idcs: workspace too small");
        return RGBD_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->W * p->H * p->D;
    const int nblocks = (n + kThreads - 1) / kThreads;
    int *counts = (int *)workspace, *offsets = counts + nblocks, *total = offsets + nblocks;
    k_dv_count<<<nblocks, kThreads, 0, st>>>(*p, cam2world, world2grid, counts);
    k_dv_scan<<<1, 1024, 0, st>>>(counts, nblocks, offsets, total);
    k_dv_compact<<<nblocks, kThreads, 0, st>>>(*p, cam2world, world2grid, offsets, lin_ind, voxel_coords, n);
    count_launch(3);
    int rc = check_launch("rgbd_dv_compute_proj_idcs");
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(M_host, total, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("rgbd_dv_compute_proj_idcs: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

RGBD_API int rgbd_dv_trilinear_fwd(const float *grid, const int32_t *lin_ind, const float *voxel_coords, int ld, int M,
                          int F, const rgbd_dv_params *p, float *frustum, void *stream)
{
    if (!dv_ok(p) || !grid || !frustum || F <= 0 || M < 0 || (M > 0 && (!lin_ind || !voxel_coords))) {
        set_error("rgbd_dv_trilinear_fwd: null pointer or bad params");
        return RGBD_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->W * p->H * p->D;
    cudaError_t e = cudaMemsetAsync(frustum, 0, sizeof(float) * (size_t)F * n, st);   // xp.zeros (:415)
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    if (M > 0)
        k_dv_trilinear_fwd<<<(M + kThreads - 1) / kThreads, kThreads, 0, st>>>(grid, lin_ind, voxel_coords, ld, M, F,
                                                                               p->G, n, frustum);
    count_launch();
    return check_launch("rgbd_dv_trilinear_fwd");
}

RGBD_API int rgbd_dv_trilinear_bwd(const float *g_frustum, const int32_t *lin_ind, const float *voxel_coords, int ld, int M,
                          int F, const rgbd_dv_params *p, float *g_grid, void *stream)
{
    if (!dv_ok(p) || !g_frustum || !g_grid || F <= 0 || M < 0 || (M > 0 && (!lin_ind || !voxel_coords))) {
        set_error("rgbd_dv_trilinear_bwd: null pointer or bad params");
        return RGBD_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->W * p->H * p->D;
    const size_t G3 = (size_t)p->G * p->G * p->G;
    cudaError_t e = cudaMemsetAsync(g_grid, 0, sizeof(float) * (size_t)F * G3, st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    if (M > 0)
        k_dv_trilinear_bwd<<<(M + kThreads - 1) / kThreads, kThreads, 0, st>>>(g_frustum, lin_ind, voxel_coords, ld, M,
                                                                               F, p->G, n, g_grid);
    count_launch();
    return check_launch("rgbd_dv_trilinear_bwd");
}

static size_t dv_chunk_samples(const rgbd_dv_params *p, int B, int F)
{
    const size_t per = (size_t)p->G * p->G * p->G * F * sizeof(float);
    size_t n = ((size_t)48 << 20) / per;                 // keep the channels-last copy L2-resident
    if (n < 1) n = 1;
    if (n > (size_t)B) n = B;
    return n;
}

RGBD_API size_t rgbd_dv_project_workspace_bytes(const rgbd_dv_params *p, int B, int F)
{
    if (!dv_ok(p) || B <= 0 || F <= 0) return 0;
    return dv_chunk_samples(p, B, F) * (size_t)p->G * p->G * p->G * F * sizeof(float);
}

RGBD_API int rgbd_dv_project_fwd(const rgbd_dv_params *p, const float *grid, const float *cam2world, int B, int F,
                        float *frustum, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!dv_ok(p) || !grid || !cam2world || !frustum || B <= 0 || F <= 0) {
        set_error("rgbd_dv_project_fwd: null pointer or bad params");
        return RGBD_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->W * p->H * p->D;
    if (!workspace || (F & 3)) {                          // planar fallback: no staging memory, or F not a multiple of 4
        dim3 grid_dim((n + kThreads - 1) / kThreads, B);
        k_dv_project_fwd<<<grid_dim, kThreads, 0, st>>>(*p, grid, cam2world, F, frustum);
        count_launch();
        return check_launch("rgbd_dv_project_fwd");
    }
    if (workspace_bytes < rgbd_dv_project_workspace_bytes(p, B, F)) {
        set_error("rgbd_dv_project_fwd: workspace too small");
        return RGBD_E_WORKSPACE;
    }
    const int G3 = p->G * p->G * p->G;
    const int Bs = (int)dv_chunk_samples(p, B, F);
    float *cl = (float *)workspace;
    const char *ex = getenv("RGBD_B200_DV_EXACT");
    const bool exact = ex && ex[0] == '1';
    for (int b0 = 0; b0 < B; b0 += Bs) {
        const int nb = (B - b0 < Bs) ? (B - b0) : Bs;
        k_dv_to_cl<<<dim3((G3 + 31) / 32, (F + 31) / 32, nb), dim3(32, 8), 0, st>>>(grid + (size_t)b0 * F * G3, cl, F, G3);
        const dim3 gd((n + 32 * kDvWarps - 1) / (32 * kDvWarps), nb);
        if (exact) k_dv_project_fwd_cl<true><<<gd, 32 * kDvWarps, 0, st>>>(*p, cl, cam2world + 16 * (size_t)b0, F, frustum + (size_t)b0 * F * n);
        else k_dv_project_fwd_cl<false><<<gd, 32 * kDvWarps, 0, st>>>(*p, cl, cam2world + 16 * (size_t)b0, F, frustum + (size_t)b0 * F * n);
        count_launch(2);
    }
    return check_launch("rgbd_dv_project_fwd");
}

RGBD_API int rgbd_dv_project_bwd(const rgbd_dv_params *p, const float *g_frustum, const float *cam2world, int B, int F,
                        float *g_grid, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!dv_ok(p) || !g_frustum || !cam2world || !g_grid || B <= 0 || F <= 0) {
        set_error("rgbd_dv_project_bwd: null pointer or bad params");
        return RGBD_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->W * p->H * p->D;
    const size_t G3 = (size_t)p->G * p->G * p->G;
    if (!workspace || (F & 3)) {
        cudaError_t e = cudaMemsetAsync(g_grid, 0, sizeof(float) * (size_t)B * F * G3, st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        dim3 grid_dim((n + kThreads - 1) / kThreads, B);
        k_dv_project_bwd<<<grid_dim, kThreads, 0, st>>>(*p, g_frustum, cam2world, F, g_grid);
        count_launch();
        return check_launch("rgbd_dv_project_bwd");
    }
    if (workspace_bytes < rgbd_dv_project_workspace_bytes(p, B, F)) {
        set_error("rgbd_dv_project_bwd: workspace too small");
        return RGBD_E_WORKSPACE;
    }
    const int Bs = (int)dv_chunk_samples(p, B, F);
    float *gcl = (float *)workspace;
    for (int b0 = 0; b0 < B; b0 += Bs) {
        const int nb = (B - b0 < Bs) ? (B - b0) : Bs;
        cudaError_t e = cudaMemsetAsync(gcl, 0, sizeof(float) * (size_t)nb * F * G3, st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        k_dv_project_bwd_cl<<<dim3((n + 32 * kDvWarps - 1) / (32 * kDvWarps), nb), 32 * kDvWarps, 0, st>>>(
            *p, g_frustum + (size_t)b0 * F * n, cam2world + 16 * (size_t)b0, F, gcl);
        k_dv_from_cl<<<dim3(((int)G3 + 31) / 32, (F + 31) / 32, nb), dim3(32, 8), 0, st>>>(gcl, g_grid + (size_t)b0 * F * G3, F, (int)G3);
        count_launch(2);
    }
    return check_launch("rgbd_dv_project_bwd");
}

static bool render_ok(const rgbd_dv_params *p, const rgbd_dv_render_params *r, int B, int F)
{
    if (!dv_ok(p) || !r || B <= 0 || F <= 0) { set_error("rgbd_dv_render: null pointer or bad params"); return false; }
    if (r->nf != kRenderNf || (F & 3) || F > 32 || p->D > kRenderMaxD) {
        set_error("rgbd_dv_render: built for occnet_nf == %d, F %% 4 == 0, F <= 32, D <= %d (got nf %d, F %d, D %d)",
                  kRenderNf, kRenderMaxD, r->nf, F, p->D);
        return false;
    }
    return true;
}

struct RenderWs { size_t cl, gcl, partials, total; int Bs, nblk, nvals; };

// samples per launch of the render kernels: a sample has only H*W rays (4096 = 1024 warps), so one sample cannot fill
// the GPU; when the L2-sized chunk of the projection kernels is a single sample (G = 64: 33.5 MB per staged grid)
// several samples are launched together even though their staged grids then exceed L2 (measured, r01_tuning.md)
static int render_chunk_samples(const rgbd_dv_params *p, int B, int F)
{
    int n = (int)dv_chunk_samples(p, B, F);
    const char *e = getenv("RGBD_B200_DV_RENDER_CHUNK");
    const int want = (e && atoi(e) > 0) ? atoi(e) : 8;
    if (n < want) n = want;
    if (n > B) n = B;
    return n;
}

static RenderWs render_ws(const rgbd_dv_params *p, int B, int F)
{
    RenderWs w;
    const size_t per = (size_t)p->G * p->G * p->G * F * sizeof(float);
    w.Bs = render_chunk_samples(p, B, F);
    w.nblk = (p->W * p->H + 4 * kDvWarps - 1) / (4 * kDvWarps);
    w.nvals = kRenderNf * (F + 1) + 2 * kRenderNf + 1;
    w.cl = 0;
    w.gcl = (w.Bs * per + 255) / 256 * 256;
    w.partials = 2 * w.gcl;
    w.total = w.partials + ((size_t)B * w.nblk * w.nvals * sizeof(float) + 255) / 256 * 256;
    return w;
}

RGBD_API size_t rgbd_dv_render_saved_bytes(const rgbd_dv_params *p, int B)
{
    if (!dv_ok(p) || B <= 0) return 0;
    return (size_t)B * p->W * p->H * (p->D + 1) * sizeof(float);
}

RGBD_API size_t rgbd_dv_render_workspace_bytes(const rgbd_dv_params *p, int B, int F)
{
    if (!dv_ok(p) || B <= 0 || F <= 0) return 0;
    return render_ws(p, B, F).total;
}

RGBD_API int rgbd_dv_render_fwd(const rgbd_dv_params *p, const rgbd_dv_render_params *r, const float *grid,
                       const float *cam2world, const float *W1, const float *b1, const float *W2, const float *b2,
                       int B, int F, float *novel, float *depth, float *fg, float *saved, void *workspace,
                       size_t workspace_bytes, void *stream)
{
    if (!render_ok(p, r, B, F)) return RGBD_E_UNSUPPORTED;
    if (!grid || !cam2world || !W1 || !b1 || !W2 || !b2 || !novel || !depth) {
        set_error("rgbd_dv_render_fwd: null pointer");
        return RGBD_E_ARG;
    }
    const RenderWs L = render_ws(p, B, F);
    if (!workspace || workspace_bytes < L.total) { set_error("rgbd_dv_render_fwd: workspace too small"); return RGBD_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int G3 = p->G * p->G * p->G, HW = p->W * p->H;
    float *cl = (float *)((char *)workspace + L.cl);
    for (int b0 = 0; b0 < B; b0 += L.Bs) {
        const int nb = (B - b0 < L.Bs) ? (B - b0) : L.Bs;
        k_dv_to_cl<<<dim3((G3 + 31) / 32, (F + 31) / 32, nb), dim3(32, 8), 0, st>>>(grid + (size_t)b0 * F * G3, cl, F, G3);
        k_dv_render_fwd<<<dim3(L.nblk, nb), 32 * kDvWarps, 0, st>>>(*p, *r, cl, cam2world + 16 * (size_t)b0, W1, b1, W2, b2, F,
                                                                  novel + (size_t)b0 * F * HW, depth + (size_t)b0 * HW,
                                                                  fg ? fg + (size_t)b0 * HW : nullptr,
                                                                  saved ? saved + (size_t)b0 * HW * (p->D + 1) : nullptr);
        count_launch(2);
    }
    return check_launch("rgbd_dv_render_fwd");
}

RGBD_API int rgbd_dv_render_bwd(const rgbd_dv_params *p, const rgbd_dv_render_params *r, const float *grid,
                       const float *cam2world, const float *W1, const float *b1, const float *W2, const float *b2,
                       int B, int F, const float *saved, const float *g_novel, const float *g_depth, const float *g_fg,
                       float *g_grid, float *g_W1, float *g_b1, float *g_W2, float *g_b2, void *workspace,
                       size_t workspace_bytes, void *stream)
{
    if (!render_ok(p, r, B, F)) return RGBD_E_UNSUPPORTED;
    if (!grid || !cam2world || !W1 || !b1 || !W2 || !b2 || !g_novel || !g_depth || !g_grid || !g_W1 || !g_b1 || !g_W2 || !g_b2) {
        set_error("rgbd_dv_render_bwd: null pointer");
        return RGBD_E_ARG;
    }
    const RenderWs L = render_ws(p, B, F);
    if (!workspace || workspace_bytes < L.total) { set_error("rgbd_dv_render_bwd: workspace too small"); return RGBD_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int G3 = p->G * p->G * p->G, HW = p->W * p->H;
    float *cl = (float *)((char *)workspace + L.cl), *gcl = (float *)((char *)workspace + L.gcl);
    float *partials = (float *)((char *)workspace + L.partials);
    for (int b0 = 0; b0 < B; b0 += L.Bs) {
        const int nb = (B - b0 < L.Bs) ? (B - b0) : L.Bs;
        k_dv_to_cl<<<dim3((G3 + 31) / 32, (F + 31) / 32, nb), dim3(32, 8), 0, st>>>(grid + (size_t)b0 * F * G3, cl, F, G3);
        cudaError_t e = cudaMemsetAsync(gcl, 0, sizeof(float) * (size_t)nb * F * G3, st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        k_dv_render_bwd<<<dim3(L.nblk, nb), 32 * kDvWarps, 0, st>>>(
            *p, *r, cl, cam2world + 16 * (size_t)b0, W1, b1, W2, b2, F, g_novel + (size_t)b0 * F * HW,
            g_depth + (size_t)b0 * HW, g_fg ? g_fg + (size_t)b0 * HW : nullptr, gcl,
            partials + (size_t)b0 * L.nblk * L.nvals, L.nvals, saved ? saved + (size_t)b0 * HW * (p->D + 1) : nullptr);
        k_dv_from_cl<<<dim3((G3 + 31) / 32, (F + 31) / 32, nb), dim3(32, 8), 0, st>>>(gcl, g_grid + (size_t)b0 * F * G3, F, G3);
        count_launch(3);
    }
    k_dv_render_wgrad<<<L.nvals, 256, 0, st>>>(partials, B * L.nblk, L.nvals, F, g_W1, g_b1, g_W2, g_b2);
    count_launch();
    return check_launch("rgbd_dv_render_bwd");
}

}  // extern "C"
