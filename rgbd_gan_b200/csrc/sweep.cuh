// sweep.cuh -- the C == 4 consistency loss as ONE persistent, TMA-fed row sweep (included by consistency.cu).
//
// Reference: LossFuncRotate.__call__ (common/loss_functions.py:63-146) with warp / inv_warp (:171-182) and
// bilinear (:185-228) and its scatter_add backward, both directions of every pair, forward + backward in one pass.
//
// Why: the three-kernel chain (stage-in / main / stage-out) moves every pixel through L2 eleven times; this
// kernel moves it five times and touches HBM exactly once per input and once per output plane.
//
//   * All rows of all pairs form one linear sequence of BLOCKS of kSwR rows.  CTA i (one per SM, persistent) owns a
//     contiguous range of blocks [LBs, LBe) -- perfectly balanced for every batch size -- and sweeps it top to bottom,
//     both warp directions in lock step.
//   * One thread (of the retire warps) streams the rows of BOTH images of the current pair straight from the caller's NCHW planes
//     into a shared-memory ring with cp.async.bulk (TMA, SASS UBLKCP) + mbarrier full / empty pairs: each row is
//     fetched once and serves as "own pixels" for one direction and as gather window for the other.  The window is
//     the own block +- kSwReach blocks (own row -16 .. +23 rows at kSwR = 8); the 2-tap gathers are LDS.
//   * The gradient scatter (GetItem backward, :226-227) goes with 16-byte vector REDs into a small CTA-private NHWC
//     ring in global memory that stays L2-resident (kSwRingRows rows per image); own-pixel terms go there too.
//   * Retire warps follow the sweep: once a block of rows can no longer be hit they read it back from L2, scale it,
//     transpose it and write the caller's NCHW gradient planes with plain coalesced stores (each output row is
//     written exactly once by its owner; nothing needs zeroing) and zero the ring rows behind the sweep for re-use.
//     compute -> retire ordering: CTA-scope release / acquire through an mbarrier (the ring is CTA-private), signalled
//     in the middle of the next step.
//   * Taps outside the window (pathological poses / depths) and scatter targets outside the CTA's own rows
//     (chunk boundaries) are exact but slow: gathers read the planes through L2, scatter contributions are appended
//     to a per-CTA list that k_sweep_fixup adds to the finished planes afterwards.
//
// Remaining nondeterminism: the order of the fp32 REDs into one ring entry (<= ~10 addends) and of the fixup
// atomics; bounded at 1e-6 of the gradient max-norm by tests/test_gpu_consistency.py.
#pragma once

namespace rgbd {

constexpr int kSwR = 8;                                  // rows per block
constexpr int kSwNB = 6;                                 // shared-memory ring: blocks per image
constexpr int kSwNR = kSwR * kSwNB;                      // ... rows per image
constexpr int kSwReach = 2;                              // gather / scatter window: own block +- 2 blocks
constexpr int kSwWin = (2 * kSwReach + 1) * kSwR;        // window rows
constexpr int kSwComputeWarps = 2 * kSwR;                // one warp per (row of the block, direction)
constexpr int kSwRetireWarps = 4;                        // 16 + 4 = 20 warps: 5 per scheduler, 96 registers each
constexpr int kSwThreads = (kSwComputeWarps + kSwRetireWarps) * 32;
constexpr int kSwRingRows = 64;                          // global gradient ring: rows per image per CTA (power of two)
constexpr int kSwBarBytes = 512;                        // barriers + reduction scratch behind the planes

struct __align__(16) SweepRec {                          // one out-of-ring scatter contribution (both taps of a pixel)
    int img;                                             // sel * B + b of the TARGET image
    int pix;                                             // u0 * W + v0
    int pad0, pad1;
    float4 ta, tb;                                       // addends for pixel pix and pix + 1 (NHWC order)
};

struct SweepArgs {
    const float *img, *img_rot;                          // (B,4,H,W)
    float *g_img, *g_img_rot;                            // (B,4,H,W), GRAD
    const float *M, *c, *Mi, *ci;
    float4 *ring;                                        // [ncta][2][kSwRingRows][W]
    float4 *gz;                                          // debug mode (RING == false): global [2][B][HW] accumulator
    SweepRec *ovf;                                       // [ncta][ovf_cap]
    int *ovf_count;                                      // [ncta]
    float2 *partials;                                    // [2][ncta]
    float *hinge_partials;                               // [2][ncta] or null
    int B, H, W, HW, bpp, total_blocks, ovf_cap;
    float Hm1f;                                          // (float)(H - 1)
    int norm, occ;
    float k_rgb, k_d, max_depth, min_depth;              // +inf / -inf = no depth-range mask
    float hinge_min, hinge_coef;                         // NaN = off; coef = gradient factor of relu(hinge_min - z)
    float scale;
    const float *scale_dev;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier, L2 evict-first (the planes are read once)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void red_add_v4(float4 *p, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 ld_cg_v4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg_zero_v4(float4 *p)
{
    asm volatile("st.global.cg.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(0.0f) : "memory");
}
// The gradient ring is CTA-private: REDs (compute warps), write-back loads and re-zeroing stores (retire warps) all
// come from ONE CTA, so the release / acquire pairs between them only need CTA scope (MEMBAR.ALL.CTA, a few cycles)
// instead of a GPU-scope fence that waits for every outstanding RED to be acknowledged by L2 (measured: 14 % of the
// compute warps' time in profiles/r02_sweep_v2_256.csv).  The loads on the other side are L2 loads (ld.global.cg).
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// two IEEE-correct divisions by the same denominator for TWO pixels with one shared (rare) slow path: the refined
// reciprocal + residual step of div2_rn() is bit-identical to __fdiv_rn for numerators in [2^-60, 2^60] and
// denominators in [1e-4, 1e4] (tests/test_gpu_sweep.py::test_div2_matches_ieee_division); everything else --
// including exact zeros -- takes __fdiv_rn itself behind ONE branch per pair of pixels.
__device__ __forceinline__ void div_fast(float a0, float a1, float b, float &q0, float &q1, float &rinv, bool &bad)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float t = __fmaf_rn(-b, r, 1.0f);
    r = __fmaf_rn(r, t, r);
    float x0 = __fmul_rn(a0, r), x1 = __fmul_rn(a1, r);
    q0 = __fmaf_rn(r, __fmaf_rn(-b, x0, a0), x0);
    q1 = __fmaf_rn(r, __fmaf_rn(-b, x1, a1), x1);
    rinv = r;
    const float f0 = fabsf(a0), f1 = fabsf(a1);
    bad = !((fminf(f0, f1) > 8.6736174e-19f) && (fmaxf(f0, f1) < 1.1529215e18f));      // 2^-60, 2^60 (NaN -> bad)
}

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// warp-uniform values of one step of one compute warp (one row of a block, one direction)
struct SweepStep {
    float4 pA, pB, pC;        // rows of K R K^-1: (pA.x pA.y pA.z) (pA.w pB.x pB.y) (pB.z pB.w pC.x); subtracted vector (pC.y pC.z pC.w)
    float y;                  // own row as float
    uint32_t own_sa;          // shared address of (own row, column lane) in plane 0 of the own image
    uint32_t soth_sa;         // shared address of the other image's ring
    int ws, wlo;              // ring row and image row of the window's first row
    int rlo;                  // rows [rlo, rlo + rn) of the other image scatter into the ring (window AND owned)
    unsigned rn;
    int Grow0;                // linear row of this pair's row 0
    unsigned gown_off, goth_off;   // float4 offsets into the gradient accumulator: own row (+ lane), other image
    int p, dir;
};

// The per-pixel work of one iteration: the two pixels (columns c0 + lane, c0 + 32 + lane) of this lane.  `mid()` runs
// between the gathers and the first RED (the row-sweep kernel publishes / waits there).
template <int W, bool L1, bool LOSS, bool GRAD, bool RING, bool HINGE, typename Mid>
__device__ __forceinline__ void sweep_pixels(const SweepArgs &a, const SweepStep &st, const int c0, const int lane,
                                             float4 *const gbase, int *scount, const int cta, const float xlane,
                                             float &s_rgb, float &s_d, float &s_h, Mid mid)
{
    constexpr uint32_t PSB = kSwNR * W * sizeof(float);
    constexpr float Wm1 = (float)(W - 1);
    const int HW = a.HW;
    const float Hm1 = a.Hm1f, y = st.y;
    const float4 pA = st.pA, pB = st.pB, pC = st.pC;
    const uint32_t own_sa = st.own_sa, soth_sa = st.soth_sa;
    const int ws = st.ws, wlo = st.wlo, rlo = st.rlo, Grow0 = st.Grow0, p = st.p, dir = st.dir;
    const unsigned rn = st.rn, gown_off = st.gown_off, goth_off = st.goth_off;
        // ---- phase 1+2: own pixels, warp / inv_warp (:171-182), bilinear coordinates (:199-216)
        float4 own[2];
        float q2v[2], vcolv[2], urowv[2], rinvv[2], q0v[2], q1v[2], zcv[2];
        bool badv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t sa = own_sa + (uint32_t)(c0 + 32 * k) * 4u;
            own[k] = make_float4(lds_f32(sa), lds_f32(sa + PSB), lds_f32(sa + 2 * PSB), lds_f32(sa + 3 * PSB));
            const float z = own[k].w, x = xlane + (float)(c0 + 32 * k);
            const float P0 = __fmul_rn(z, x), P1 = __fmul_rn(z, y);              // z * p, K=3 fma chains, minus c
            q0v[k] = __fsub_rn(__fmaf_rn(pA.z, z, __fmaf_rn(pA.y, P1, __fmul_rn(pA.x, P0))), pC.y);
            q1v[k] = __fsub_rn(__fmaf_rn(pB.y, z, __fmaf_rn(pB.x, P1, __fmul_rn(pA.w, P0))), pC.z);
            q2v[k] = __fsub_rn(__fmaf_rn(pC.x, z, __fmaf_rn(pB.w, P1, __fmul_rn(pB.z, P0))), pC.w);
            zcv[k] = fminf(fmaxf(q2v[k], 1e-4f), 10000.0f);
            div_fast(q0v[k], q1v[k], zcv[k], vcolv[k], urowv[k], rinvv[k], badv[k]);
        }
        if (badv[0] || badv[1]) {                                 // rare: tiny / huge / zero numerators
#pragma unroll
            for (int k = 0; k < 2; ++k) { vcolv[k] = __fdiv_rn(q0v[k], zcv[k]); urowv[k] = __fdiv_rn(q1v[k], zcv[k]); }
        }
        float wav[2], wbv[2], wcv[2], wdv[2];
        int u0v[2], v0v[2];
        bool mv[2], okv[2];
        uint32_t tapv[2];
        bool any_far = false;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float urow = urowv[k], vcol = vcolv[k];
            const bool m = (urow >= 0.0f) && (urow < Hm1) && (vcol >= 0.0f) && (vcol < Wm1) && (q2v[k] > 1e-4f);
            // in bounds: 0 <= u0 <= H-2, 0 <= v0 <= W-2 (masked pixels never use their indices)
            const int u0 = __float2int_rz(urow), v0 = __float2int_rz(vcol);
            const float u0f = (float)u0, v0f = (float)v0;
            wav[k] = __fsub_rn(u0f + 1.0f, urow); wbv[k] = __fsub_rn(urow, u0f);   // (u1-u), (u-u0)
            wcv[k] = __fsub_rn(v0f + 1.0f, vcol); wdv[k] = __fsub_rn(vcol, v0f);   // (v1-v), (v-v0)
            const int rel = u0 - wlo;
            const bool inw = (unsigned)rel < (unsigned)kSwWin;
            int sr = ws + rel;
            sr = sr >= kSwNR ? sr - kSwNR : sr;
            mv[k] = m; okv[k] = m && inw; u0v[k] = u0; v0v[k] = v0;
            // taps that are masked or outside the window read the own pixel instead (always resident; the values
            // are never used: masked pixels are excluded below, far taps are re-read through L2)
            tapv[k] = okv[k] ? soth_sa + (uint32_t)(sr * W + v0) * 4u : own_sa + (uint32_t)(c0 + 32 * k) * 4u;
            any_far = any_far || (m && !inw);
        }
        // ---- phase 3: the 2-tap gathers; both row taps read row u0 (:219)
        float4 Av[2], Bv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t tp = tapv[k];
            Av[k] = make_float4(lds_f32(tp), lds_f32(tp + PSB), lds_f32(tp + 2 * PSB), lds_f32(tp + 3 * PSB));
            Bv[k] = make_float4(lds_f32(tp + 4), lds_f32(tp + PSB + 4), lds_f32(tp + 2 * PSB + 4), lds_f32(tp + 3 * PSB + 4));
        }
        if (__any_sync(0xffffffffu, any_far)) {                   // outside the window: through L2, exact but slow
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (mv[k] && !okv[k]) {
                    const float *tp = (dir ? a.img : a.img_rot) + (size_t)p * 4 * HW + (size_t)u0v[k] * W + v0v[k];
                    Av[k] = make_float4(__ldg(tp), __ldg(tp + HW), __ldg(tp + 2 * (size_t)HW), __ldg(tp + 3 * (size_t)HW));
                    Bv[k] = make_float4(__ldg(tp + 1), __ldg(tp + HW + 1), __ldg(tp + 2 * (size_t)HW + 1),
                                        __ldg(tp + 3 * (size_t)HW + 1));
                }
        }
    mid();
        // ---- phase 4: blend (:226-227), residuals (:107-110), occlusion (:114), loss, gradients
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 A = Av[k], B4 = Bv[k], ow = own[k];
            const float q2 = q2v[k];
            const float w1 = __fmul_rn(wav[k], wcv[k]), w2 = __fmul_rn(wbv[k], wcv[k]),
                        w3 = __fmul_rn(wav[k], wdv[k]), w4 = __fmul_rn(wbv[k], wdv[k]);
#define RGBD_BLEND(ch) __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, A.ch), __fmul_rn(w2, A.ch)), __fmul_rn(w3, B4.ch)), __fmul_rn(w4, B4.ch))
            // sampled depth in the reference's exact order (it decides the occlusion mask); masked pixels carry
            // arbitrary finite values and are excluded through mv[] below
            const float wdp = RGBD_BLEND(w);
            const bool o = a.occ ? (wdp > q2) : true;                         // :114 strict >
            const bool sd = (ow.w < a.max_depth) && (ow.w > a.min_depth);     // :121-135
            float gzo = 0.0f;                                                 // own-pixel depth gradient
            bool own_red = false;
            if (HINGE) {                                                      // updater.py:357-359
                const float h = fmaxf(a.hinge_min - ow.w, 0.0f);
                if (LOSS) s_h += h * h;
                if (GRAD && h > 0.0f) { gzo = a.hinge_coef * h; own_red = true; }
            }
            float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f;
            if (mv[k] && o && sd) {
                // colour residuals only feed sums and signs: two-term blend (wA*A + wB*B), and the exact four-term
                // order only where a residual is so small that its sign could depend on the rounding
                const float wA = w1 + w2, wB = w3 + w4;
                float d0 = __fmaf_rn(wB, B4.x, wA * A.x) - ow.x, d1 = __fmaf_rn(wB, B4.y, wA * A.y) - ow.y,
                      d2 = __fmaf_rn(wB, B4.z, wA * A.z) - ow.z;
                const float d3 = __fsub_rn(wdp, q2);
                if (fminf(fminf(fabsf(d0), fabsf(d1)), fabsf(d2)) < 1e-5f) {
                    d0 = __fsub_rn(RGBD_BLEND(x), ow.x); d1 = __fsub_rn(RGBD_BLEND(y), ow.y); d2 = __fsub_rn(RGBD_BLEND(z), ow.z);
                }
                if (LOSS) {
                    if (L1) { s_rgb += (fabsf(d0) + fabsf(d1)) + fabsf(d2); s_d += fabsf(d3); }
                    else { s_rgb += (d0 * d0 + d1 * d1) + d2 * d2; s_d += d3 * d3; }
                }
                if (GRAD) {
                    constexpr int NORM = L1 ? RGBD_NORM_L1 : RGBD_NORM_L2;
                    e0 = sign_coeff(NORM, a.k_rgb, d0); e1 = sign_coeff(NORM, a.k_rgb, d1);
                    e2 = sign_coeff(NORM, a.k_rgb, d2);
                    const float e3 = sign_coeff(NORM, a.k_d, d3);
                    const int u0 = u0v[k], v0 = v0v[k];
                    if (!RING || (unsigned)(u0 - rlo) < rn) {                 // scatter-add (GetItem backward)
                        float4 *gt = gbase + (goth_off + (RING ? ((Grow0 + u0) & (kSwRingRows - 1)) * W + v0 : u0 * W + v0));
                        red_add_v4(gt, e0 * wA, e1 * wA, e2 * wA, e3 * wA);
                        red_add_v4(gt + 1, e0 * wB, e1 * wB, e2 * wB, e3 * wB);
                    } else {                                                  // not this CTA's row (or out of window)
                        const int slot = atomicAdd(scount, 1);
                        SweepRec r;
                        r.img = (1 - dir) * a.B + p; r.pix = u0 * W + v0; r.pad0 = r.pad1 = 0;
                        r.ta = make_float4(e0 * wA, e1 * wA, e2 * wA, e3 * wA);
                        r.tb = make_float4(e0 * wB, e1 * wB, e2 * wB, e3 * wB);
                        a.ovf[(size_t)cta * a.ovf_cap + slot] = r;
                    }
                    const float GA = ((e0 * A.x + e1 * A.y) + e2 * A.z) + e3 * A.w;
                    const float GB = ((e0 * B4.x + e1 * B4.y) + e2 * B4.z) + e3 * B4.w;
                    // weights -> column coordinate only (row gradient cancels); Div / Clip backward
                    const float g_v = (GB - GA) * (wav[k] + wbv[k]);
                    const float gq0 = g_v * rinvv[k];
                    float gq2 = -e3;
                    if (q2 >= 1e-4f && q2 <= 10000.0f) gq2 -= gq0 * vcolv[k];
                    // MatMul backward (M^T gq, gq1 = 0) and z*p backward
                    const float gP0 = pA.x * gq0 + pB.z * gq2, gP1 = pA.y * gq0 + pB.w * gq2, gP2 = pA.z * gq0 + pC.x * gq2;
                    const float x = xlane + (float)(c0 + 32 * k);
                    gzo += (gP0 * x + gP1 * y) + gP2;
                    own_red = true;
                }
            }
            if (GRAD && own_red) red_add_v4(gbase + (gown_off + c0 + 32 * k), -e0, -e1, -e2, gzo);
#undef RGBD_BLEND
        }
}

// shared-memory layout: planes [2 images][4 channels][kSwNR rows][W] floats, then the barriers
//   full[kSwNB] : TMA completion;  done[4] : compute -> retire ("step t is finished": its window rows were read, its REDs
//   issued -- frees the oldest window block for the next TMA load and the oldest ring block for write-back);
//   zero[1 + 4] : retire -> compute ("the ring starts zeroed", "iteration t has written back and re-zeroed its block")
template <int W, bool L1, bool LOSS, bool GRAD, bool RING, bool HINGE>
__global__ void __launch_bounds__(kSwThreads, 1) k_consistency_sweep(const SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char sw_smem[];
    float *sring = reinterpret_cast<float *>(sw_smem);
    const int HW = a.HW;
    constexpr int plane_stride = kSwNR * W;                  // floats
    constexpr int img_stride = 4 * plane_stride;
    constexpr uint32_t PSB = plane_stride * sizeof(float);   // plane stride in bytes
    const uint32_t bars = smem_u32(sw_smem + (size_t)2 * img_stride * sizeof(float));
    const uint32_t bar_full = bars, bar_done = bars + 16 * kSwNB, bar_zero = bar_done + 32;
    float *sred = reinterpret_cast<float *>(sw_smem + (size_t)2 * img_stride * sizeof(float) + 16 * kSwNB + 80);   // [3][16]
    int *scount = reinterpret_cast<int *>(sred + 3 * kSwComputeWarps);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int LBs = (int)((long long)cta * a.total_blocks / ncta);
    const int LBe = (int)((long long)(cta + 1) * a.total_blocks / ncta);
    const int Lfirst = LBs - kSwReach > 0 ? LBs - kSwReach : 0;
    const int Llast = LBe + kSwReach - 1 < a.total_blocks - 1 ? LBe + kSwReach - 1 : a.total_blocks - 1;
    const int nsteps = LBe - LBs;

    if (threadIdx.x == 0) {
        for (int k = 0; k < kSwNB; ++k) mbar_init(bar_full + 8 * k, 1);
        for (int k = 0; k < 4; ++k) mbar_init(bar_done + 8 * k, kSwComputeWarps);
        for (int k = 0; k < 5; ++k) mbar_init(bar_zero + 8 * k, kSwRetireWarps);      // [0]: initial zero, [1..4]: retire iterations
        *scount = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();                       // inputs, ring and lists of the previous kernels on the stream are complete

    if (warp < kSwComputeWarps) {
        // ------------------------------------------------------------------ compute: warp = (row of block, direction)
        const int ri = warp & (kSwR - 1), dir = warp / kSwR;
        const uint32_t sown_sa = smem_u32(sring + (dir ? img_stride : 0));
        const uint32_t soth_sa = smem_u32(sring + (dir ? 0 : img_stride));
        // gradient accumulator: the CTA's ring [2][kSwRingRows][W] (debug variant: the global [2][B][HW] buffer)
        float4 *const gbase = RING ? a.ring + (size_t)cta * 2 * kSwRingRows * W : a.gz;
        const int Gs = LBs * kSwR, Gn = nsteps * kSwR;        // owned linear rows [Gs, Gs + Gn)
        const float xlane = (float)lane;
        float s_rgb = 0.0f, s_d = 0.0f, s_h = 0.0f;
        float4 pA = make_float4(0.f, 0.f, 0.f, 0.f), pB = pA, pC = pA;
        int cur_pair = -1, q_waited = -1, wait_slot = 0, wait_par = 0;
        int own_slot = (LBs - Lfirst) % kSwNB;                               // ring slot of the own block
        int win_slot = (LBs - kSwReach - Lfirst + kSwNB) % kSwNB;            // ring slot of the window's first block

        int p = LBs / a.bpp, j = LBs - p * a.bpp - 1;          // pair and block-in-pair of the step (advanced below)
#pragma unroll 1
        for (int t = 0; t < nsteps; ++t) {
            const int LB = LBs + t;
            if (++j == a.bpp) { j = 0; ++p; }
            if (p != cur_pair) {
                // rows of K R K^-1: (pA.x pA.y pA.z) (pA.w pB.x pB.y) (pB.z pB.w pC.x); subtracted vector (pC.y pC.z pC.w)
                const float *Ms = (dir ? a.Mi : a.M) + 9 * (size_t)p, *cs = (dir ? a.ci : a.c) + 3 * (size_t)p;
                pA = make_float4(__ldg(Ms), __ldg(Ms + 1), __ldg(Ms + 2), __ldg(Ms + 3));
                pB = make_float4(__ldg(Ms + 4), __ldg(Ms + 5), __ldg(Ms + 6), __ldg(Ms + 7));
                pC = make_float4(__ldg(Ms + 8), __ldg(cs), __ldg(cs + 1), __ldg(cs + 2));
                cur_pair = p;
            }
            {   // the window's blocks have landed
                const int top = LB + kSwReach < Llast ? LB + kSwReach : Llast;
                const int q_need = top - Lfirst;
                while (q_waited < q_need) {
                    ++q_waited;
                    mbar_wait(bar_full + 8 * wait_slot, (uint32_t)wait_par);
                    if (++wait_slot == kSwNB) { wait_slot = 0; wait_par ^= 1; }
                }
            }
            // ---- warp-uniform values of this step
            const int i = j * kSwR + ri;                                  // own row
            const float y = (float)i;
            const int own_srow = own_slot * kSwR + ri;
            const int ws = win_slot * kSwR;                               // its ring row
            const int wlo = (j - kSwReach) * kSwR;                        // first row of the window
            const int Grow0 = (LB - j) * kSwR;                            // linear row of this pair's row 0
            // rows of the other image whose scatter goes to the ring: inside the window AND owned by this CTA
            int rlo = Gs - Grow0, rhi = Gs + Gn - Grow0;
            rlo = rlo > wlo ? rlo : wlo;
            rhi = rhi < wlo + kSwWin ? rhi : wlo + kSwWin;
            const unsigned rn = rhi > rlo ? (unsigned)(rhi - rlo) : 0u;
            const uint32_t own_sa = sown_sa + (uint32_t)(own_srow * W + lane) * 4u;
            const unsigned gown_off = (RING ? (unsigned)(dir * kSwRingRows + ((Grow0 + i) & (kSwRingRows - 1))) * W
                                            : (unsigned)(dir * a.B + p) * HW + (unsigned)i * W) + lane;
            const unsigned goth_off = RING ? (unsigned)((1 - dir) * kSwRingRows) * W : (unsigned)((1 - dir) * a.B + p) * HW;

            SweepStep st;
            st.pA = pA; st.pB = pB; st.pC = pC; st.y = y; st.own_sa = own_sa; st.soth_sa = soth_sa; st.ws = ws; st.wlo = wlo;
            st.rlo = rlo; st.rn = rn; st.Grow0 = Grow0; st.gown_off = gown_off; st.goth_off = goth_off; st.p = p; st.dir = dir;
#pragma unroll
            for (int c0 = 0; c0 < W; c0 += 64) {                          // (unrolled: +1.6 % measured)
                sweep_pixels<W, L1, LOSS, GRAD, RING, HINGE>(a, st, c0, lane, gbase, scount, cta, xlane, s_rgb, s_d, s_h, [&]() {
                    if (GRAD && RING && c0 == 0) {
                        // the ring rows this step can hit must be zero: the first kSwRingRows rows were zeroed at the start, a
                        // later block re-uses the rows of block LB + kSwReach - kSwRingRows / kSwR, which retire iteration
                        // t - kLag wrote back and zeroed (far behind: this wait does not block)
                        constexpr int kLag = kSwRingRows / kSwR - 2 * kSwReach;
                        if (t == 0) mbar_wait(bar_zero, 0);
                        if (t >= kLag) mbar_wait(bar_zero + 8 + 8 * ((t - kLag) & 3), (uint32_t)(((t - kLag) >> 2) & 1));
                    }
                });
            }
            // step finished: the oldest block of the window is dead, the REDs of this step are issued (CTA-scope release)
            if (GRAD && RING) fence_cta();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_done + 8 * (t & 3));
            own_slot = own_slot + 1 == kSwNB ? 0 : own_slot + 1;
            win_slot = win_slot + 1 == kSwNB ? 0 : win_slot + 1;
        }
        // ---- per-CTA partial sums (fixed order: warp shuffle tree, then rows 0..7 of each direction)
        s_rgb = warp_sum(s_rgb); s_d = warp_sum(s_d); s_h = warp_sum(s_h);
        if (lane == 0) { sred[warp] = s_rgb; sred[kSwComputeWarps + warp] = s_d; sred[2 * kSwComputeWarps + warp] = s_h; }
        asm volatile("bar.sync 2, %0;" ::"n"(kSwComputeWarps * 32) : "memory");
        if (ri == 0 && lane == 0) {
            float r = 0.0f, d = 0.0f, h = 0.0f;
            for (int k = 0; k < kSwR; ++k) {
                r += sred[dir * kSwR + k]; d += sred[kSwComputeWarps + dir * kSwR + k]; h += sred[2 * kSwComputeWarps + dir * kSwR + k];
            }
            if (LOSS) a.partials[(size_t)dir * ncta + cta] = make_float2(r, d);
            if (LOSS && a.hinge_partials) a.hinge_partials[(size_t)dir * ncta + cta] = h;
            if (GRAD && RING && dir == 0) a.ovf_count[cta] = *scount;
        }
        if (GRAD && RING) {
            // tail: the last kSwReach blocks can only be written back now.  The 4 retire warps would do them one after the
            // other (an L2 round trip each, ~5 us with 16 idle warps); here every compute warp takes one (image, row) item
            // per block.  All compute warps have passed the barrier above, i.e. every RED into the ring has been issued
            // (CTA-scope fence at the end of each step); the ring needs no re-zeroing (the next launch zeroes it).
            fence_cta();
            float scale = a.scale;
            if (a.scale_dev) scale *= __ldg(a.scale_dev);
            const float4 *ringc = a.ring + (size_t)cta * 2 * kSwRingRows * W;
            const int X0 = LBe - kSwReach > LBs ? LBe - kSwReach : LBs;
            const int sel = warp / kSwR, r = warp % kSwR;                  // 16 warps = 2 images x 8 rows
            float4 v[kSwReach][W / 32];
#pragma unroll
            for (int n = 0; n < kSwReach; ++n)
                if (X0 + n < LBe) {
                    const float4 *src = ringc + ((size_t)sel * kSwRingRows + (((X0 + n) * kSwR + r) & (kSwRingRows - 1))) * W + lane;
#pragma unroll
                    for (int k = 0; k < W / 32; ++k) v[n][k] = ld_cg_v4(src + 32 * k);
                }
#pragma unroll
            for (int n = 0; n < kSwReach; ++n)
                if (X0 + n < LBe) {
                    const int X = X0 + n, pp = X / a.bpp, jj = X - pp * a.bpp;
                    float *dst = (sel ? a.g_img_rot : a.g_img) + (size_t)pp * 4 * HW + (size_t)(jj * kSwR + r) * W + lane;
#pragma unroll
                    for (int k = 0; k < W / 32; ++k) {
                        __stcs(dst + 32 * k, v[n][k].x * scale); __stcs(dst + HW + 32 * k, v[n][k].y * scale);
                        __stcs(dst + 2 * (size_t)HW + 32 * k, v[n][k].z * scale); __stcs(dst + 3 * (size_t)HW + 32 * k, v[n][k].w * scale);
                    }
                }
        }
    } else {
        // ------------------------------------------------------------------ retire warps (warp 0, lane 0 also drives the TMA)
        const int rw = warp - kSwComputeWarps;
        const bool retire = GRAD && RING;
        uint64_t policy = 0;
        if (rw == 0 && lane == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        // stream block L (kSwR rows of both images, 4 planes each) into its ring slot
        auto load_block = [&](int L) {
            constexpr uint32_t row_bytes = (uint32_t)(kSwR * W * sizeof(float));      // one plane of one block
            const int slot = (L - Lfirst) % kSwNB;
            const int p = L / a.bpp, j = L - p * a.bpp;
            mbar_expect_tx(bar_full + 8 * slot, 8 * row_bytes);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float *src = (s ? a.img_rot : a.img) + (size_t)p * 4 * HW + (size_t)j * kSwR * W;
                const uint32_t dst = smem_u32(sring + (size_t)s * img_stride + (size_t)slot * kSwR * W);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    bulk_g2s(dst + ch * PSB, src + (size_t)ch * HW, row_bytes, bar_full + 8 * slot, policy);
            }
        };
        if (rw == 0 && lane == 0) {                                       // fill the ring: every block no finished step frees
            const int last0 = LBs + kSwNB - kSwReach - 1 < Llast ? LBs + kSwNB - kSwReach - 1 : Llast;
            for (int L = Lfirst; L <= last0; ++L) load_block(L);
        }
        if (!retire && rw != 0) return;
        float scale = a.scale;
        if (retire && a.scale_dev) scale *= __ldg(a.scale_dev);
        float4 *ring = a.ring + (size_t)cta * 2 * kSwRingRows * W;
        constexpr int kItems = 2 * kSwR / kSwRetireWarps;                 // (image, row) items of a block per warp
        auto zero_block = [&](int X) {
#pragma unroll
            for (int n = 0; n < kItems; ++n) {
                const int item = rw * kItems + n, sel = item / kSwR, r = item % kSwR;
                float4 *dst = ring + ((size_t)sel * kSwRingRows + ((X * kSwR + r) & (kSwRingRows - 1))) * W;
#pragma unroll
                for (int col = lane; col < W; col += 32) st_cg_zero_v4(dst + col);
            }
        };
        // write back one finished block: ALL loads of the warp's (image, row) items in flight at once (one L2 round trip
        // per block), then zero the ring rows (same thread, same address: ordered) and store the four planes
        // (coalesced 128-byte rows per warp)
        auto retire_block = [&](int X) {
            const int p = X / a.bpp, j = X - p * a.bpp;
            float4 v[kItems][W / 32];
#pragma unroll
            for (int n = 0; n < kItems; ++n) {
                const int item = rw * kItems + n, sel = item / kSwR, r = item % kSwR;
                const float4 *src = ring + ((size_t)sel * kSwRingRows + ((X * kSwR + r) & (kSwRingRows - 1))) * W + lane;
#pragma unroll
                for (int k = 0; k < W / 32; ++k) v[n][k] = ld_cg_v4(src + 32 * k);
            }
#pragma unroll
            for (int n = 0; n < kItems; ++n) {
                const int item = rw * kItems + n, sel = item / kSwR, r = item % kSwR;
                float4 *src = ring + ((size_t)sel * kSwRingRows + ((X * kSwR + r) & (kSwRingRows - 1))) * W + lane;
                float *dst = (sel ? a.g_img_rot : a.g_img) + (size_t)p * 4 * HW + (size_t)(j * kSwR + r) * W + lane;
#pragma unroll
                for (int k = 0; k < W / 32; ++k) {
                    st_cg_zero_v4(src + 32 * k);
                    // written once, read by the caller's next kernels: streaming stores keep L2 for the ring
                    __stcs(dst + 32 * k, v[n][k].x * scale); __stcs(dst + HW + 32 * k, v[n][k].y * scale);
                    __stcs(dst + 2 * (size_t)HW + 32 * k, v[n][k].z * scale); __stcs(dst + 3 * (size_t)HW + 32 * k, v[n][k].w * scale);
                }
            }
        };
        if (retire) {   // the ring rows of the first blocks (later blocks re-use rows that were zeroed behind the sweep)
            const int nz = nsteps < kSwRingRows / kSwR ? nsteps : kSwRingRows / kSwR;
            for (int X = LBs; X < LBs + nz; ++X) zero_block(X);
            fence_cta();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_zero);
        }
#pragma unroll 1
        for (int t = 0; t < nsteps; ++t) {
            mbar_wait_relaxed(bar_done + 8 * (t & 3), (uint32_t)((t >> 2) & 1));   // every compute warp has finished step t
            if (rw == 0 && lane == 0) {                                    // block LB - kSwReach is dead: its slot takes LB + 4
                const int L = LBs + t + kSwNB - kSwReach;
                if (L <= Llast) load_block(L);
            }
            if (retire) {
                const int Xr = LBs + t - kSwReach;                         // no later step can hit this block's ring rows
                if (Xr >= LBs) retire_block(Xr);
                fence_cta();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_zero + 8 + 8 * (t & 3));
            }
        }
        // (the last kSwReach blocks are written back by the compute warps, see the end of their branch)
    }
}

// Adds the out-of-ring scatter contributions to the finished gradient planes (scalar atomics; a few percent of the
// taps at chunk boundaries, everything outside the window for pathological inputs); the last block finishes the loss.
constexpr int kSwFixSplit = 4;                           // fix-up blocks per list
__global__ void __launch_bounds__(kThreads)
k_sweep_fixup(const SweepArgs a, int nlists, const FinalizeArgs fin)
{
    pdl_launch_dependents();
    pdl_wait();
    if ((int)blockIdx.x == nlists * kSwFixSplit) {
        if (fin.partials) loss_finalize_block(fin);
        return;
    }
    float scale = a.scale;
    if (a.scale_dev) scale *= __ldg(a.scale_dev);
    const int list = blockIdx.x / kSwFixSplit, part = blockIdx.x % kSwFixSplit;
    const int n = a.ovf_count[list];
    const SweepRec *recs = a.ovf + (size_t)list * a.ovf_cap;
    for (int k = part * kThreads + threadIdx.x; k < n; k += kSwFixSplit * kThreads) {
        const SweepRec r = recs[k];
        const int sel = r.img >= a.B ? 1 : 0, b = r.img - sel * a.B;
        float *g = (sel ? a.g_img_rot : a.g_img) + (size_t)b * 4 * a.HW + r.pix;
        atomicAdd(g, r.ta.x * scale); atomicAdd(g + 1, r.tb.x * scale);
        atomicAdd(g + a.HW, r.ta.y * scale); atomicAdd(g + a.HW + 1, r.tb.y * scale);
        atomicAdd(g + 2 * (size_t)a.HW, r.ta.z * scale); atomicAdd(g + 2 * (size_t)a.HW + 1, r.tb.z * scale);
        atomicAdd(g + 3 * (size_t)a.HW, r.ta.w * scale); atomicAdd(g + 3 * (size_t)a.HW + 1, r.tb.w * scale);
    }
}

// ---- test aid: the shared-reciprocal divisions against IEEE division on the device (tests/test_gpu_sweep.py)
__device__ __forceinline__ unsigned sw_hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// numerators: sign * 2^e * mantissa with e uniform in [e_lo, e_hi]; denominators: log-uniform in [1e-4, 1e4] (the range
// of clip(q2, 1e-4, 1e4), :199) plus the exact bounds.  counts[0] = mismatches of div_fast + its fallback rule (sweep
// kernel), counts[1] = mismatches of div2_rn (three-kernel chain), counts[2] = samples that took a fallback.
__global__ void __launch_bounds__(kThreads)
k_debug_div2(unsigned long long n, unsigned seed, int e_lo, int e_hi, unsigned long long *counts)
{
    unsigned long long bad_fast = 0, bad_chain = 0, slow = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * kThreads) {
        const unsigned h0 = sw_hash((unsigned)k * 3u + seed), h1 = sw_hash((unsigned)k * 3u + 1u + seed),
                       h2 = sw_hash((unsigned)k * 3u + 2u + seed + (unsigned)(k >> 32));
        auto num = [&](unsigned h) {
            const int e = e_lo + (int)((h >> 24) % (unsigned)(e_hi - e_lo + 1));
            const float m = __uint_as_float(0x3f800000u | (h & 0x7fffffu));            // [1, 2)
            const float v = ldexpf(m, e);
            return (h & 0x800000u) ? -v : v;
        };
        const float a0 = (k % 1021 == 0) ? 0.0f : num(h0), a1 = num(h1);
        float b = __expf(-9.2103404f + 18.420681f * ((h2 >> 8) * (1.0f / 16777216.0f)));   // log-uniform [1e-4, 1e4]
        b = fminf(fmaxf(b, 1e-4f), 10000.0f);
        if (k % 4099 == 0) b = (k & 1) ? 1e-4f : 10000.0f;
        const float r0 = __fdiv_rn(a0, b), r1 = __fdiv_rn(a1, b);
        float q0, q1, rinv;
        bool bad;
        div_fast(a0, a1, b, q0, q1, rinv, bad);
        if (bad) { q0 = __fdiv_rn(a0, b); q1 = __fdiv_rn(a1, b); ++slow; }
        if (__float_as_uint(q0) != __float_as_uint(r0) || __float_as_uint(q1) != __float_as_uint(r1)) ++bad_fast;
        float c0, c1, cr;
        div2_rn(a0, a1, b, c0, c1, cr);
        if (__float_as_uint(c0) != __float_as_uint(r0) || __float_as_uint(c1) != __float_as_uint(r1)) ++bad_chain;
    }
    if (bad_fast) atomicAdd(counts, bad_fast);
    if (bad_chain) atomicAdd(counts + 1, bad_chain);
    if (slow) atomicAdd(counts + 2, slow);
}

}  // namespace rgbd
