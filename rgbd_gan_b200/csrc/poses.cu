// poses.cu -- the pose pipeline on the device: CameraParamPrior.sample -> get_camera_matries ->
// R, t -> K R K^-1, (K R) t, K R^T K^-1, -(K t) as ONE launch with one thread per pair, so that a training step needs
// no host round trip (the reference: np.random + NumPy on the host, xp.array() upload, then ~10 tiny cuBLAS / elementwise
// launches inside LossFuncRotate.__call__).
//
// Reference: train_rgbd.py:192-217 (CameraParamPrior), updater.py:26-60 (update_camera_matrices, get_camera_matries),
// common/loss_functions.py:85-91 (R, inv_R, t), :174 / :181 (the constant factors of warp / inv_warp).
//
// Rounding orders.  These 24 floats per pair feed truncated pixel indices, so every product below is evaluated in the
// order the reference's CPU path (NumPy over OpenBLAS, the order the golden vectors were produced with) evaluates it:
//   * 3x3 . 3x3 and 4x4 . 4x4 products: fma(a_k, b_k, ... fma(a_1, b_1, rn(a_0 b_0)))          
//   * R1^T (t2 - t1) (transposed, strided operand): the same chain
//   * (B,3,3) @ (B,3,1) with contiguous operands -- (K R) t and K t -- come out of a different BLAS kernel:
//     rows 0-1 rn(rn(rn(a0 b0) + rn(a1 b1)) + rn(a2 b2)), row 2 fma(a2, b2, fma(a0, b0, rn(a1 b1)))
// tests/test_gpu_poses.py checks all of them bit for bit against every golden case (cam, and through warp: new_zp).
// The one operation that cannot be pinned is NumPy's fp32 cos / sin (SIMD routines, < 1.5 ulp): the kernel either takes
// the caller's cos / sin (bit-exact chain) or evaluates them in double and rounds once (<= 1 ulp from NumPy's).
#include <math.h>

#include "common.cuh"

namespace rgbd {

struct PoseArgs {
    int B;
    // ---- stage A: CameraParamPrior.sample
    int do_sample;
    const double *draws;           // (B,15): uniform(-1,1) x6 | uniform(0,0.5) x6 | choice(2) x3, or null: Philox below
    unsigned long long seed, step;
    double range[6];               // camera_param_range = x,y,z rotate | x,y,z translate
    int uniform;
    // ---- stage B: get_camera_matries
    int do_cam;
    int rows_only;                 // rgbd_pose_camera_matrices: B independent rows, no pairing
    const float *thetas_in;        // (2B,6) when !do_sample
    const float *cos_sin;          // optional (2B,6): cos of the 3 angles | sin of the 3 angles
    int order[3];
    float *thetas_out;             // (2B,6) or null
    float *cam_out;                // (2B,4,4) or null
    // ---- stage C: pose algebra
    const float *theta, *theta_rot;   // (B,4,4) each when !do_cam
    float K[9], iK[9];
    float *M, *c, *Mi, *ci;        // (B,9) (B,3) (B,9) (B,3); M == null: stage C off
};

// Philox4x32-10 (Salmon et al. 2011): counter-based, so a step needs no generator state in device memory
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1)
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in [0,1) from two words, the construction of NumPy's random_sample
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// out = A . B, n x n, the BLAS fma chain
template <int N>
__device__ __forceinline__ void matmul_chain(const float *A, const float *Bm, float *out)
{
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float s = __fmul_rn(A[i * N], Bm[j]);
#pragma unroll
            for (int l = 1; l < N; ++l) s = __fmaf_rn(A[i * N + l], Bm[l * N + j], s);
            out[i * N + j] = s;
        }
}
// A (3x3) . v (3) in the order of the BLAS kernel behind contiguous (B,3,3) @ (B,3,1) products
__device__ __forceinline__ void matvec_blas(const float *A, const float *v, float *out)
{
#pragma unroll
    for (int i = 0; i < 2; ++i)
        out[i] = __fadd_rn(__fadd_rn(__fmul_rn(A[3 * i], v[0]), __fmul_rn(A[3 * i + 1], v[1])), __fmul_rn(A[3 * i + 2], v[2]));
    out[2] = __fmaf_rn(A[8], v[2], __fmaf_rn(A[6], v[0], __fmul_rn(A[7], v[1])));
}

// train_rgbd.py:199-217, float64 like NumPy, one operation per rounding (no contraction)
__device__ void sample_pair(const PoseArgs &a, int b, float th[2][6])
{
    double u[6], e[6], sg[3];
    if (a.draws) {
        const double *d = a.draws + 15 * (size_t)b;
        for (int k = 0; k < 6; ++k) { u[k] = d[k]; e[k] = d[6 + k]; }
        for (int k = 0; k < 3; ++k) sg[k] = __dsub_rn(__dmul_rn(d[12 + k], 2.0), 1.0);      // choice(2) * 2 - 1
    } else {
        uint32_t w[32];
        for (int q = 0; q < 8; ++q)
            philox4x32((uint32_t)b, (uint32_t)q, (uint32_t)a.step, (uint32_t)(a.step >> 32), (uint32_t)a.seed,
                       (uint32_t)(a.seed >> 32), w + 4 * q);
        for (int k = 0; k < 6; ++k) {
            u[k] = __dadd_rn(-1.0, __dmul_rn(2.0, u53(w[2 * k], w[2 * k + 1])));              // uniform(-1, 1)
            e[k] = __dmul_rn(0.5, u53(w[12 + 2 * k], w[13 + 2 * k]));                        // uniform(0, 0.5)
        }
        for (int k = 0; k < 3; ++k) sg[k] = (w[24 + k] & 1u) ? 1.0 : -1.0;
    }
    for (int k = 0; k < 3; ++k) {
        const double rr = a.range[k];
        double lim = __ddiv_rn(1.0, __dadd_rn(rr, 1e-8));                                     // np.clip(1 / (range + 1e-8), 0, 1)
        lim = fmin(fmax(lim, 0.0), 1.0);
        double f = sg[k];
        if (!a.uniform)                                                                       // sign * (r == 3.1415) + |sign| * (r != 3.1415)
            f = __dadd_rn(__dmul_rn(sg[k], rr == 3.1415 ? 1.0 : 0.0), __dmul_rn(fabs(sg[k]), rr != 3.1415 ? 1.0 : 0.0));
        e[k] = __dmul_rn(__dmul_rn(e[k], f), lim);
    }
    for (int k = 0; k < 6; ++k) {
        const double sgn = u[k] > 0.0 ? 1.0 : (u[k] < 0.0 ? -1.0 : 0.0);                      // np.sign
        double t2 = __dadd_rn(__dmul_rn(-e[k], sgn), u[k]);
        if (a.uniform) {                                                                      // reflect into [-1, 1]
            const double in = (-1.0 <= t2 ? 1.0 : 0.0), in2 = (t2 <= 1.0 ? 1.0 : 0.0);
            const double lo = t2 < -1.0 ? 1.0 : 0.0, hi = t2 > 1.0 ? 1.0 : 0.0;
            t2 = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(t2, in), in2), __dmul_rn(__dsub_rn(-2.0, t2), lo)),
                           __dmul_rn(__dsub_rn(2.0, t2), hi));
        }
        th[0][k] = (float)__dmul_rn(u[k], a.range[k]);                                        // * camera_param_range, astype(float32)
        th[1][k] = (float)__dmul_rn(t2, a.range[k]);
    }
}

// updater.py:45-60 for one row of thetas
__device__ void camera_matrix(const PoseArgs &a, const float *th, const float *cs, float *mat)
{
#pragma unroll
    for (int k = 0; k < 16; ++k) mat[k] = 0.0f;
    mat[0] = 1.0f; mat[5] = 1.0f; mat[10] = -1.0f; mat[15] = 1.0f; mat[11] = 1.0f;
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
        const int i = a.order[s], a1 = (i + 1) % 3, a2 = (i + 2) % 3;
        float co, si;
        if (cs) { co = cs[i]; si = cs[3 + i]; }
        else { co = (float)cos((double)th[i]); si = (float)sin((double)th[i]); }
        float rot[16], nxt[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) rot[k] = (k % 5 == 0) ? 1.0f : 0.0f;
        rot[a1 * 4 + a1] = co; rot[a1 * 4 + a2] = -si; rot[a2 * 4 + a1] = si; rot[a2 * 4 + a2] = co;
        matmul_chain<4>(rot, mat, nxt);
#pragma unroll
        for (int k = 0; k < 16; ++k) mat[k] = nxt[k];
    }
    mat[3] = __fadd_rn(mat[3], th[3]); mat[7] = __fadd_rn(mat[7], th[4]); mat[11] = __fadd_rn(mat[11], th[5]);
}

__global__ void __launch_bounds__(128) k_pose_pipeline(const PoseArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    float th[2][6], cam[2][16];
    if (a.rows_only) {                                              // get_camera_matries on its own: thread = row
        for (int k = 0; k < 6; ++k) th[0][k] = a.thetas_in[6 * (size_t)b + k];
        camera_matrix(a, th[0], a.cos_sin ? a.cos_sin + 6 * (size_t)b : nullptr, cam[0]);
        for (int k = 0; k < 16; ++k) a.cam_out[16 * (size_t)b + k] = cam[0][k];
        return;
    }
    if (a.do_sample) sample_pair(a, b, th);
    else if (a.do_cam)
        for (int k = 0; k < 6; ++k) { th[0][k] = a.thetas_in[6 * (size_t)b + k]; th[1][k] = a.thetas_in[6 * (size_t)(a.B + b) + k]; }
    if (a.do_sample && a.thetas_out)
        for (int k = 0; k < 6; ++k) { a.thetas_out[6 * (size_t)b + k] = th[0][k]; a.thetas_out[6 * (size_t)(a.B + b) + k] = th[1][k]; }
    if (a.do_cam) {
        camera_matrix(a, th[0], a.cos_sin ? a.cos_sin + 6 * (size_t)b : nullptr, cam[0]);
        camera_matrix(a, th[1], a.cos_sin ? a.cos_sin + 6 * (size_t)(a.B + b) : nullptr, cam[1]);
        if (a.cam_out)
            for (int k = 0; k < 16; ++k) { a.cam_out[16 * (size_t)b + k] = cam[0][k]; a.cam_out[16 * (size_t)(a.B + b) + k] = cam[1][k]; }
    } else if (a.M) {
        for (int k = 0; k < 16; ++k) { cam[0][k] = a.theta[16 * (size_t)b + k]; cam[1][k] = a.theta_rot[16 * (size_t)b + k]; }
    }
    if (!a.M) return;
    // ---- common/loss_functions.py:85-91
    float R1[9], R2t[9], R1t[9], R[9], Rt[9], dt[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R1[i * 3 + j] = cam[0][i * 4 + j]; R1t[j * 3 + i] = cam[0][i * 4 + j]; R2t[j * 3 + i] = cam[1][i * 4 + j];
        }
    matmul_chain<3>(R2t, R1, R);                                   // R = R2^T R1
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Rt[j * 3 + i] = R[i * 3 + j];   // inv_R = R^T
        dt[i] = __fsub_rn(cam[1][i * 4 + 3], cam[0][i * 4 + 3]);    // t2 - t1
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)                                     // t = R1^T (t2 - t1)
        t[i] = __fmaf_rn(R1t[i * 3 + 2], dt[2], __fmaf_rn(R1t[i * 3 + 1], dt[1], __fmul_rn(R1t[i * 3], dt[0])));
    // ---- :174 / :181
    float KR[9], KRi[9], M[9], Mi[9], c[3], kt[3];
    matmul_chain<3>(a.K, R, KR);
    matmul_chain<3>(KR, a.iK, M);
    matvec_blas(KR, t, c);
    matmul_chain<3>(a.K, Rt, KRi);
    matmul_chain<3>(KRi, a.iK, Mi);
    matvec_blas(a.K, t, kt);
#pragma unroll
    for (int k = 0; k < 9; ++k) { a.M[9 * (size_t)b + k] = M[k]; a.Mi[9 * (size_t)b + k] = Mi[k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.c[3 * (size_t)b + k] = c[k]; a.ci[3 * (size_t)b + k] = -kt[k]; }
}

static int launch_pose(const PoseArgs &a, void *stream, const char *what)
{
    k_pose_pipeline<<<(a.B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch(what);
}

static bool fill_algebra(PoseArgs &a, const float *K, const float *inv_K, float *M, float *c, float *Mi, float *ci, const char *what)
{
    if (!K || !inv_K || !M || !c || !Mi || !ci) { set_error("%s: null K / inv_K / output pointer", what); return false; }
    for (int k = 0; k < 9; ++k) { a.K[k] = K[k]; a.iK[k] = inv_K[k]; }
    a.M = M; a.c = c; a.Mi = Mi; a.ci = ci;
    return true;
}

static bool fill_order(PoseArgs &a, const int *order, const char *what)
{
    static const int dflt[3] = {0, 1, 2};
    if (!order) order = dflt;
    for (int k = 0; k < 3; ++k) {
        if (order[k] < 0 || order[k] > 2) { set_error("%s: order entries must be 0, 1 or 2", what); return false; }
        a.order[k] = order[k];
    }
    return true;
}

static bool fill_prior(PoseArgs &a, const rgbd_pose_prior *prior, const double *draws, unsigned long long seed,
                       unsigned long long step, const char *what)
{
    if (!prior) { set_error("%s: null prior", what); return false; }
    for (int k = 0; k < 6; ++k) a.range[k] = prior->camera_param_range[k];
    a.uniform = prior->uniform_distribution ? 1 : 0;
    a.do_sample = 1; a.draws = draws; a.seed = seed; a.step = step;
    return true;
}

}  // namespace rgbd

extern "C" {

RGBD_API int rgbd_pose_sample(const rgbd_pose_prior *prior, int B, const double *draws, unsigned long long seed,
                              unsigned long long step, float *thetas, void *stream)
{
    using namespace rgbd;
    PoseArgs a = {};
    if (B <= 0 || !thetas) { set_error("rgbd_pose_sample: B <= 0 or null output"); return RGBD_E_ARG; }
    if (!fill_prior(a, prior, draws, seed, step, "rgbd_pose_sample")) return RGBD_E_ARG;
    a.B = B; a.thetas_out = thetas;
    return launch_pose(a, stream, "rgbd_pose_sample");
}

RGBD_API int rgbd_pose_camera_matrices(const float *thetas, const float *cos_sin, int n_rows, const int *order,
                                       float *cam2world, void *stream)
{
    using namespace rgbd;
    PoseArgs a = {};
    if (n_rows <= 0 || !thetas || !cam2world) { set_error("rgbd_pose_camera_matrices: n_rows <= 0 or null pointer"); return RGBD_E_ARG; }
    if (!fill_order(a, order, "rgbd_pose_camera_matrices")) return RGBD_E_ARG;
    a.B = n_rows; a.do_cam = 1; a.rows_only = 1; a.thetas_in = thetas; a.cos_sin = cos_sin; a.cam_out = cam2world;
    return launch_pose(a, stream, "rgbd_pose_camera_matrices");
}

RGBD_API int rgbd_pose_algebra(const float *theta, const float *theta_rot, int B, const float *K, const float *inv_K,
                               float *M, float *c, float *Mi, float *ci, void *stream)
{
    using namespace rgbd;
    PoseArgs a = {};
    if (B <= 0 || !theta || !theta_rot) { set_error("rgbd_pose_algebra: B <= 0 or null pointer"); return RGBD_E_ARG; }
    if (!fill_algebra(a, K, inv_K, M, c, Mi, ci, "rgbd_pose_algebra")) return RGBD_E_ARG;
    a.B = B; a.theta = theta; a.theta_rot = theta_rot;
    return launch_pose(a, stream, "rgbd_pose_algebra");
}

RGBD_API int rgbd_pose_pipeline(const rgbd_pose_prior *prior, int B, const double *draws, unsigned long long seed,
                                unsigned long long step, const int *order, const float *K, const float *inv_K, float *thetas,
                                float *cam2world, float *M, float *c, float *Mi, float *ci, void *stream)
{
    using namespace rgbd;
    PoseArgs a = {};
    if (B <= 0) { set_error("rgbd_pose_pipeline: B <= 0"); return RGBD_E_ARG; }
    if (!fill_prior(a, prior, draws, seed, step, "rgbd_pose_pipeline")) return RGBD_E_ARG;
    if (!fill_order(a, order, "rgbd_pose_pipeline")) return RGBD_E_ARG;
    if (!fill_algebra(a, K, inv_K, M, c, Mi, ci, "rgbd_pose_pipeline")) return RGBD_E_ARG;
    a.B = B; a.do_cam = 1; a.thetas_out = thetas; a.cam_out = cam2world;
    return launch_pose(a, stream, "rgbd_pose_pipeline");
}

}
