/*
 * rgbdgan_b200.h -- C-ABI of librgbdgan_b200.so: the B200 (sm_100a) implementation of
 * RGBD-GAN's 3D-consistency hot path.
 *
 * The reference (nogu-atsu/RGBD-GAN) is pure Python: it has no FFI/plugin layer.  Its
 * boundary for this path is the Python call surface of common/loss_functions.py
 * (LossFuncRotate, warp, inv_warp, bilinear) and deepvoxel/projection.py +
 * deepvoxel/deepvoxel.py::interpolate_trilinear.  These entry points are what a
 * Chainer FunctionNode (or any other host binding: ctypes, CuPy, torch) calls in place
 * of the chain of Chainer/CuPy kernels the reference launches; each one cites the
 * reference lines it replaces.  INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - every pointer is a CALLER-OWNED DEVICE pointer (cupy `arr.data.ptr`, torch
 *     `tensor.data_ptr()`), fp32, C-contiguous, 16-byte aligned; the library never
 *     allocates device memory and keeps no global mutable state
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call
 *     is asynchronous on that stream and performs no host synchronisation, except
 *     rgbd_dv_compute_proj_idcs, whose result size is data dependent (as in the reference)
 *   - return value: 0 on success; > 0 a cudaError_t from a launch; < 0 an argument error
 *     (RGBD_E_*); rgbd_last_error() gives a thread-local message.  No exceptions, no prints
 *   - images are (B,C,H,W) NCHW with the depth in channel C-1; gradients have the
 *     same layout and are OVERWRITTEN (Chainer accumulates outside the node)
 *   - `workspace` is scratch for the duration of the call on `stream`, sized by the
 *     matching *_workspace_bytes(); contents are undefined afterwards
 *   - B is the number of PAIRS held by this process; `n_pairs_global` is the number of
 *     pairs the reference's means run over (== B unless the batch is sharded over GPUs),
 *     so per-shard results just add up (loss parts) or need no communication (gradients)
 */
#ifndef RGBDGAN_B200_H
#define RGBDGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGBD_B200_VERSION 100

#if defined(__GNUC__)
#define RGBD_API __attribute__((visibility("default")))
#else
#define RGBD_API
#endif

#define RGBD_E_ARG (-1)        /* null pointer / non-positive size / bad enum        */
#define RGBD_E_ALIGN (-2)      /* pointer not 16-byte aligned                        */
#define RGBD_E_WORKSPACE (-3)  /* workspace missing or too small                     */
#define RGBD_E_UNSUPPORTED (-4)

#define RGBD_NORM_L1 1 /* F.mean_absolute_error (loss_functions.py:137-138) */
#define RGBD_NORM_L2 2 /* F.mean_squared_error  (loss_functions.py:139-140) */

RGBD_API int rgbd_version(void);
RGBD_API const char *rgbd_last_error(void);

/* Measurement aids (bench.py).  rgbd_launch_count: kernels this library has launched in this
 * process so far.  rgbd_profile_hook: the next rgbd_consistency_* call on this host thread
 * records ev_start / ev_stop (cudaEvent_t, caller-owned) on its stream immediately before and
 * after its main kernel (the projection + gather + scatter kernel of the first chunk), then
 * clears the hook.                                                                          */
RGBD_API unsigned long long rgbd_launch_count(void);
RGBD_API void rgbd_profile_hook(void *ev_start, void *ev_stop);

/* Options of LossFuncRotate.__call__ (common/loss_functions.py:63-64) + __init__ (:32-37). */
typedef struct {
    int norm;             /* RGBD_NORM_L1 | RGBD_NORM_L2                                  */
    int occlusion_aware;  /* :112-119                                                     */
    float max_depth;      /* :121-127, NaN = None                                         */
    float min_depth;      /* :129-135, NaN = None                                         */
    float lambda_geometric; /* :143-144 (only used by backward entry points)              */
    long long n_pairs_global; /* denominator pairs; 0 means "= B"                         */
    void *peer_comm;      /* NULL, or a handle from rgbd_peer_comm_create: the loss parts are
                             then all-reduced over the ranks inside the finalize kernel        */
    int defer_loss;       /* with peer_comm: 1 = do not make `stream` wait for the exchange; the
                             gradients are ordered on `stream` as usual, loss_parts become valid
                             after rgbd_peer_comm_wait(comm, stream) (or the next loss call on this
                             comm).  The exchange then overlaps the stage-out of this call and the
                             stage-in of the next one.  Not usable under CUDA-graph stream capture.
                             2 = publish only: the kernel that finishes the loss pushes this rank's parts into
                             every peer's mailbox (posted NVLink stores, no wait, no side stream, no events:
                             capturable) and leaves THIS SHARD's values in loss_parts; rgbd_peer_comm_wait(comm,
                             stream) then sums the latest call's parts of all ranks in rank order into that
                             call's loss_parts (which must still be alive).  The GPUs of a box are not coupled
                             step by step: a rank may run up to 7 calls ahead of the slowest one.  Every rank
                             must call rgbd_peer_comm_wait at the same point of its call sequence.          */
    int reserved;
    float hinge_depth_min; /* "next" row fused around the loss (updater.py:357-359, C == 4 only):        */
    float hinge_lambda;    /* loss_rotate += mean(relu(depth_min - depth)^2) * lambda_depth over both      */
                           /* images; NaN depth_min or lambda <= 0 = off.  loss_parts[5] = that term,       */
                           /* loss_parts[6] = loss_parts[4] + loss_parts[5]; its gradient (times the         */
                           /* upstream gradient) is added to the depth channel of g_img / g_img_rot.          */
} rgbd_loss_opts;

/* ---- LossFuncRotate.__call__ : common/loss_functions.py:63-146 (+ warp :171-175,
 *      inv_warp :178-182, bilinear :185-228 fused) ---------------------------------------
 * Pose inputs are the constant factors the reference computes with raw xp.matmul
 * (:89-91,:174,:181); the host wrapper computes them with the same NumPy sequence:
 *   M  (B,9) = K R K^-1        c  (B,3) = (K R) t     direction img -> img_rot (c is subtracted)
 *   Mi (B,9) = K R^T K^-1      ci (B,3) = -(K t)      direction img_rot -> img (ci is subtracted)
 */

RGBD_API size_t rgbd_consistency_workspace_bytes(int B, int C, int H, int W);
/* 1 if the plain entry points (no new_zp / masks output, no upstream new_zp gradient) run the persistent row-sweep
 * kernel for this shape on the current device (C == 4, W in {64, 128}, H % 8 == 0 and enough rows per SM; env
 * RGBD_B200_SWEEP = 0 | 1 | 2 overrides the automatic choice), 0 if they run the three-kernel chain.  Both paths
 * implement the same reference lines (common/loss_functions.py:63-146,171-228); this is reporting only. */
RGBD_API int rgbd_consistency_uses_sweep(int B, int C, int H, int W);

/* Health check of the single-launch pipeline kernel (C == 4 path): the first bytes of the workspace
 * are its control block (ticket and per-pair dependency counters; all-zero between calls, initialised by the
 * kernel itself on first use and restored by its last block).  Every dependency wait is bounded (2 s); a wait
 * that expires -- only possible if the caller overwrites the head of the workspace while a call is in flight --
 * raises a sticky flag instead of hanging the GPU.  Synchronises `stream`; *status_host = 0 ok, 1 timed out. */
RGBD_API int rgbd_consistency_status(const void *workspace, void *stream, int *status_host);

/* Test aid: evaluates the kernels' shared-reciprocal division (two quotients by one denominator in [1e-4, 1e4], the
 * range of F.clip(zp2, 1e-4, 10000), loss_functions.py:199) on n pseudo-random (a0, a1, b) and compares it bit for bit
 * with IEEE division (__fdiv_rn): counts_dev[0] = mismatches of the row-sweep kernel's variant incl. its fallback rule,
 * [1] = mismatches of the three-kernel chain's variant, [2] = samples that took the fallback.  Numerator exponents are
 * uniform in [e_lo, e_hi] (e.g. -70..70 covers both sides of the 2^-60 / 2^60 fallback thresholds); exact zeros and the
 * exact denominator bounds are mixed in. */
RGBD_API int rgbd_debug_div2(unsigned long long n, unsigned seed, int e_lo, int e_hi, unsigned long long *counts_dev,
                             void *stream);

/* Test aid (no GPU needed): the ticket order of one launch of the pipeline kernel for a chunk of Bc pairs of H x W
 * images: tickets[3t..3t+2] = (role, pair, tile) with role 0 none, 1 stage-in, 2 main, 3 stage-out, 4 loss finalize;
 * *total_out = number of tickets.  tests/test_host_logic.py checks that every tile appears once and that every
 * ticket depends only on tickets with smaller numbers (the property that makes the pipeline deadlock-free).      */
RGBD_API int rgbd_debug_mega_schedule(int Bc, int H, int W, int grad, int fold, int lag_main, int lag_so,
                                      int *tickets, int max_tickets, int *total_out);

/* Forward only.  loss_parts (device, 8 floats): [0..3] = the four means of :141-144 restricted
 * to these B pairs, in the order {rgb, rgb_rot, depth, depth_rot}; [4] = the loss combined as
 * :141-144 does, (p0+p1) + (p2*lambda + p3*lambda), valid as is when the batch is not sharded
 * (sharded: all-reduce [0..3] and recombine); [5] = depth-hinge term and [6] = [4] + [5] (see
 * rgbd_loss_opts.hinge_*; [5] = 0 when off); [7] = 0, reserved.
 * new_zp  (nullable): (2B,HW,3), the second return value of __call__ (:146).
 * masks   (nullable): (2,2B,HW) uint8 debug planes: [0] = not_getting_out (:215-216),
 *                     [1] = not_occluded (:114-115; 1 when occlusion_aware == 0).          */
RGBD_API int rgbd_consistency_fwd(const float *img, const float *img_rot, const float *M, const float *c,
                         const float *Mi, const float *ci, int B, int C, int H, int W,
                         const rgbd_loss_opts *opts, float *loss_parts, float *new_zp,
                         uint8_t *masks, void *workspace, size_t workspace_bytes, void *stream);

/* Backward by recomputation, for an arbitrary upstream gradient of the loss (what Chainer's
 * autograd does for `loss_gen.backward()`, updater.py:387).  The upstream gradient is
 * gy * (*gy_dev): gy_dev (nullable) is a DEVICE scalar, so a FunctionNode can hand over the
 * gradient array it was given without a device->host read.  g_new_zp (nullable, (2B,HW,3)) is
 * the upstream gradient of the second output.  g_img / g_img_rot are overwritten.            */
RGBD_API int rgbd_consistency_bwd(const float *img, const float *img_rot, const float *M, const float *c,
                         const float *Mi, const float *ci, int B, int C, int H, int W,
                         const rgbd_loss_opts *opts, float gy, const float *gy_dev,
                         const float *g_new_zp, float *g_img, float *g_img_rot, void *workspace,
                         size_t workspace_bytes, void *stream);

/* Forward + backward in one pass for an upstream gradient known in advance (in the
 * reference's training loop gy is the constant lambda_rotate, updater.py:363-365):
 * same outputs as rgbd_consistency_fwd followed by rgbd_consistency_bwd(gy); new_zp nullable. */
RGBD_API int rgbd_consistency_fwd_bwd(const float *img, const float *img_rot, const float *M, const float *c,
                             const float *Mi, const float *ci, int B, int C, int H, int W,
                             const rgbd_loss_opts *opts, float gy, float *loss_parts,
                             float *new_zp, float *g_img, float *g_img_rot, void *workspace,
                             size_t workspace_bytes, void *stream);

/* Companion of rgbd_consistency_fwd_bwd for autograd bindings: when the upstream gradient
 * that finally arrives (*gy_dev) differs from gy_expected, multiply both stashed gradients by
 * (*gy_dev / gy_expected) in place; when it is equal (the normal case) the kernel exits
 * without touching memory.  n_elems = elements per gradient tensor (multiple of 4).        */
RGBD_API int rgbd_consistency_rescale(float *g_img, float *g_img_rot, size_t n_elems, const float *gy_dev,
                             float gy_expected, void *stream);

/* ---- multi-GPU: the only exchange of the path is the sum of the four loss means over the
 * ranks that shard the pairs (the reference's analogue is ChainerMN's pure_nccl communicator,
 * train_rgbd.py:103-113).  Instead of a separate NCCL launch, the finalize kernel of
 * rgbd_consistency_fwd / _fwd_bwd exchanges the 16 bytes itself over NVLink peer memory:
 *   1. every rank: rgbd_peer_comm_create(rank, world, &comm, handle)   (allocates a 10 KB mailbox)
 *   2. all-gather the 64-byte handles with any host-side transport (torch.distributed, MPI)
 *   3. every rank: rgbd_peer_comm_connect(comm, all_handles (world*64 bytes, rank order))
 *   4. set rgbd_loss_opts.peer_comm = comm and n_pairs_global = total pairs; every rank must make
 *      the same sequence of loss calls.  loss_parts[0..4] then hold the GLOBAL values on every
 *      rank, summed in rank order (bit-identical on all ranks and from run to run).
 * One box only (CUDA IPC), world <= 16.                                                     */
RGBD_API int rgbd_peer_comm_create(int rank, int world, void **comm_out, unsigned char *ipc_handle_out);
RGBD_API int rgbd_peer_comm_connect(void *comm, const unsigned char *all_handles);
RGBD_API int rgbd_peer_comm_destroy(void *comm);
/* make `stream` wait for the most recent (deferred) loss exchange of this comm; after defer_loss == 2 calls: launch, ON
 * `stream`, the kernel that sums the latest call's published parts into its loss_parts (`stream` must be the stream of
 * those loss calls or ordered after them) */
RGBD_API int rgbd_peer_comm_wait(void *comm, void *stream);
/* Health of the exchange: every wait for a peer's flag inside the finalize kernel is bounded (2 s, env
 * RGBD_B200_PEER_TIMEOUT_MS); a wait that expires -- a rank died, or made a different sequence of loss calls --
 * raises a sticky flag instead of hanging every GPU of the box, and the loss parts of that call are undefined.
 * Synchronises the comm's side stream and `stream`; *status_host = 0 ok, 1 a wait timed out (or, defer_loss == 2, a
 * peer overwrote an epoch before it was summed: the ranks made different call sequences). */
RGBD_API int rgbd_peer_comm_status(void *comm, void *stream, int *status_host);
/* Test aid: instead of rgbd_peer_comm_connect, point every OTHER rank's mailbox at a local buffer nobody writes
 * (single process, single GPU): the next sharded loss call then exercises the bounded wait. */
RGBD_API int rgbd_debug_peer_comm_loopback(void *comm);

/* ---- "next" row: the pose pipeline on the device -------------------------------------------------
 * CameraParamPrior.sample (train_rgbd.py:192-217) -> get_camera_matries (updater.py:26-60) -> R, inv_R, t
 * (common/loss_functions.py:85-91) -> the constant factors of warp / inv_warp (:174, :181), one thread per pair, one
 * launch, no host round trip.  Every product is evaluated in the rounding order of the reference's CPU path (NumPy over
 * OpenBLAS; csrc/poses.cu lists the orders), so cam2world, M, c, Mi, ci are BIT-IDENTICAL to the host path on the
 * golden vectors; the exception is NumPy's fp32 cos / sin, see rgbd_pose_camera_matrices.
 *   thetas     (2B,6) fp32: rows [0,B) the first view, rows [B,2B) the rotated view (x,y,z rotation | x,y,z translation)
 *   cam2world  (2B,4,4) fp32, same row convention
 *   M, c, Mi, ci: (B,9) (B,3) (B,9) (B,3) as taken by rgbd_consistency_*  (ci = -(K t))
 * K, inv_K are HOST pointers to 9 floats (LossFuncRotate.K / inv_K, :39-56).                                          */
typedef struct {
    double camera_param_range[6]; /* config.{x,y,z}_rotate, config.{x,y,z}_translate (train_rgbd.py:194-197) */
    int uniform_distribution;     /* config.uniform_distribution                                             */
} rgbd_pose_prior;
/* CameraParamPrior.sample(2B) (:199-217) in float64 like NumPy, cast to fp32 at the end.  draws: DEVICE (B,15) float64 =
 * per pair the reference's raw draws [uniform(-1,1) x6 | uniform(0,0.5) x6 | choice(2) x3] (replaying np.random: results
 * bit-identical to the reference's), or NULL: Philox4x32-10 keyed by (seed, step, pair) -- same distribution, no state. */
RGBD_API int rgbd_pose_sample(const rgbd_pose_prior *prior, int B, const double *draws, unsigned long long seed,
                              unsigned long long step, float *thetas, void *stream);
/* get_camera_matries(thetas, order) for n_rows independent rows.  cos_sin: optional DEVICE (n_rows,6) = cos of the three
 * angles | sin of the three angles as the caller's array library computed them (then the result is bit-identical to
 * the reference's); NULL: evaluated in double and rounded once (<= 1 ulp from NumPy's fp32 routines).  order: HOST int[3]
 * or NULL = (0,1,2).                                                                                                  */
RGBD_API int rgbd_pose_camera_matrices(const float *thetas, const float *cos_sin, int n_rows, const int *order,
                                       float *cam2world, void *stream);
/* loss_functions.py:85-91 + :174 / :181 from DEVICE cam2world matrices theta, theta_rot (B,4,4): what
 * LossFuncRotate.__call__ evaluates with ~10 tiny cuBLAS launches (or, on the NumPy path, on the host).              */
RGBD_API int rgbd_pose_algebra(const float *theta, const float *theta_rot, int B, const float *K, const float *inv_K,
                               float *M, float *c, float *Mi, float *ci, void *stream);
/* all three stages in one launch; thetas / cam2world may be NULL when the caller does not need them */
RGBD_API int rgbd_pose_pipeline(const rgbd_pose_prior *prior, int B, const double *draws, unsigned long long seed,
                                unsigned long long step, const int *order, const float *K, const float *inv_K, float *thetas,
                                float *cam2world, float *M, float *c, float *Mi, float *ci, void *stream);

/* ---- free functions of common/loss_functions.py ------------------------------------------ */

/* ---- "next" row: the generators' depth head, net.py:294-299 / :756-761 -----------------------
 *   depth = 1 / (F.softplus(h[:, -1:]) + 1e-4);  h = F.concat([h[:, :3], depth])
 * h, out, g_out, g_h: (B,C,H,W); the last channel is transformed, the others are copied.  out may alias h and g_h may
 * alias g_out (in place: only the depth plane is touched).  fp32, within 1e-5 of the reference expression. */
RGBD_API int rgbd_depth_head_fwd(const float *h, int B, int C, int H, int W, float *out, void *stream);
RGBD_API int rgbd_depth_head_bwd(const float *h, const float *g_out, int B, int C, int H, int W, float *g_h, void *stream);

/* warp (:171-175) / inv_warp (:178-182): new_zp[b,n,:] = M[b] (z[b,n] * p[:,n]) - cv[b]
 * with p = (col,row,1) (:59-61).  z: (B,HW); new_zp: (B,HW,3).  inv_warp passes cv = -(K t). */
RGBD_API int rgbd_warp_fwd(const float *z, const float *M, const float *cv, int B, int H, int W,
                  float *new_zp, void *stream);
/* g_z (B,HW) = sum_k (M[b]^T g_new_zp[b,n,:])_k p_k[n]   (autograd of F.matmul and z * p) */
RGBD_API int rgbd_warp_bwd(const float *g_new_zp, const float *M, int B, int H, int W, float *g_z,
                  void *stream);

/* bilinear(img, zp) (:185-228): warped (B*HW,C), mask (B*HW) uint8 = not_getting_out */
RGBD_API int rgbd_bilinear_fwd(const float *img, const float *zp, int B, int C, int H, int W,
                      float *warped, uint8_t *mask, void *stream);
/* autograd of bilinear: g_img (B,C,H,W) and g_zp (B,HW,3), both overwritten */
RGBD_API int rgbd_bilinear_bwd(const float *img, const float *zp, const float *g_warped, int B, int C,
                      int H, int W, float *g_img, float *g_zp, void *stream);

/* ---- DeepVoxels projection: deepvoxel/projection.py:48-105 and
 *      deepvoxel/deepvoxel.py:388-428 ------------------------------------------------------- */

typedef struct {
    int W, H, D;           /* projection_image_dims[0], [1]; frustrum_depth (projection.py:56)   */
    int G;                 /* grid_dims (cubic)                                                */
    float fx, fy, cx, cy;  /* projection_intrinsic[0][0], [1][1], [0][2], [1][2] (:78-79)        */
    float voxel_size;      /* :73,:87                                                          */
    float near_plane;      /* :74 (fp32, as on the reference's NumPy-1.x / CuPy)               */
} rgbd_dv_params;

/* ProjectionHelper.compute_proj_idcs for one camera: writes the kept frustum indices in
 * ascending order to lin_ind (capacity W*H*D int32) and their voxel coordinates to
 * voxel_coords (3 rows with row stride W*H*D).  *M_host receives the count (0 is the
 * reference's `None`).  Synchronises `stream` (the count sizes the caller's arrays, like
 * the reference's host-side `.any()` / boolean indexing, :98-103).  workspace: at least
 * rgbd_dv_workspace_bytes(p).                                                               */
RGBD_API size_t rgbd_dv_workspace_bytes(const rgbd_dv_params *p);
RGBD_API int rgbd_dv_compute_proj_idcs(const rgbd_dv_params *p, const float *cam2world, int32_t *lin_ind,
                              float *voxel_coords, int *M_host, void *workspace,
                              size_t workspace_bytes, void *stream);
/* the same with the reference's optional grid2world argument (projection.py:48,53-54,83-84; no caller in the reference):
 * world2grid = inv(grid2world), DEVICE, 16 floats; grid_coords = world2grid . (cam2world . coords), both products in the
 * sgemm order, so lin_ind / voxel_coords are bit-identical to the reference's (folding the two matrices on the host is
 * not: 2e-6 on the coordinates) */
RGBD_API int rgbd_dv_compute_proj_idcs_g2w(const rgbd_dv_params *p, const float *cam2world, const float *world2grid,
                                           int32_t *lin_ind, float *voxel_coords, int *M_host, void *workspace,
                                           size_t workspace_bytes, void *stream);

/* interpolate_trilinear (deepvoxel.py:388-428) for ONE sample from explicit index lists:
 * grid (F,G,G,G) -> frustum (F,D,H,W), zero where not listed.  ld = row stride of voxel_coords. */
RGBD_API int rgbd_dv_trilinear_fwd(const float *grid, const int32_t *lin_ind, const float *voxel_coords,
                          int ld, int M, int F, const rgbd_dv_params *p, float *frustum, void *stream);
/* its autograd (the "lift" direction): g_frustum (F,D,H,W) -> g_grid (F,G,G,G), overwritten */
RGBD_API int rgbd_dv_trilinear_bwd(const float *g_frustum, const int32_t *lin_ind, const float *voxel_coords,
                          int ld, int M, int F, const rgbd_dv_params *p, float *g_grid, void *stream);

/* Fused batch path used by the generator forward (deepvoxels_generator.py:287-299 ->
 * deepvoxel.py:879-884): compute_proj_idcs + interpolate_trilinear per sample without the
 * index lists, the compaction or the host sync.  grid (B,F,G,G,G), cam2world (B,16),
 * frustum (B,F,D,H,W).  workspace (rgbd_dv_project_workspace_bytes) holds a channels-last copy
 * of a chunk of grids / grid gradients; with workspace == NULL a slower planar kernel runs.   */
RGBD_API size_t rgbd_dv_project_workspace_bytes(const rgbd_dv_params *p, int B, int F);
RGBD_API int rgbd_dv_project_fwd(const rgbd_dv_params *p, const float *grid, const float *cam2world, int B,
                        int F, float *frustum, void *workspace, size_t workspace_bytes, void *stream);
RGBD_API int rgbd_dv_project_bwd(const rgbd_dv_params *p, const float *g_frustum, const float *cam2world,
                        int B, int F, float *g_grid, void *workspace, size_t workspace_bytes, void *stream);

/* ---- "next" row: the per-sample render tail of DeepVoxels.forward with the shipped
 *      `occlusion_type: accumulative` (deepvoxel.py:879-892,903-904 around AccumulativeOcclusionNet.forward,
 *      :574-587), fused with the projection: the (B,F,D,H,W) canonical view volume is never materialised.
 *      The occlusion module's two 1x1x1 EqualizedConv3d layers (deepvoxel.py:560-567, pggan.py:27-38) are
 *      passed as plain arrays: W1 (nf, F+1) [input channel 0 = depth coordinate], b1 (nf), W2 (nf), b2 (1).   */
typedef struct {
    int nf;              /* occnet_nf; only 4 (deepvoxel.py:830) is built                                  */
    int depth_steps;     /* int(ceil(sqrt(3) * grid_dims[-1])) of the depth rescale (:903), normally == D  */
    float threshold;     /* accmulative_threshold (:556, default 4)                                        */
    float inv_c1, inv_c2; /* EqualizedConv3d input scales sqrt(2) * sqrt(1 / in_ch) (pggan.py:31)           */
} rgbd_dv_render_params;

/* workspace: channels-last staging of a chunk of grids, the same for their gradients, weight-gradient partials */
RGBD_API size_t rgbd_dv_render_workspace_bytes(const rgbd_dv_params *p, int B, int F);
/* grid (B,F,G,G,G), cam2world (B,16) -> novel (B,F,H,W), depth (B,H,W) [rescaled, :903-904], fg (B,H,W) or NULL
 * [= F.sum(weights, axis=2), the `foreground_weight` of :892].  F multiple of 4, F <= 32, D <= 128.            */
/* saved (nullable, rgbd_dv_render_saved_bytes = B*H*W*(D+1)*4): the running sums c_d of every ray, kept for the
 * backward pass (4 B per depth step where the reference keeps the F*4 B view volume and its temporaries).         */
RGBD_API size_t rgbd_dv_render_saved_bytes(const rgbd_dv_params *p, int B);
RGBD_API int rgbd_dv_render_fwd(const rgbd_dv_params *p, const rgbd_dv_render_params *r, const float *grid,
                       const float *cam2world, const float *W1, const float *b1, const float *W2, const float *b2,
                       int B, int F, float *novel, float *depth, float *fg, float *saved, void *workspace,
                       size_t workspace_bytes, void *stream);
/* its autograd by recomputation: upstream g_novel (B,F,H,W), g_depth (B,H,W), g_fg (B,H,W) or NULL ->
 * g_grid (B,F,G,G,G), g_W1 (nf,F+1), g_b1 (nf), g_W2 (nf), g_b2 (1); all overwritten.  saved: what the forward
 * call for the SAME inputs wrote, or NULL (the kernel then repeats the forward walk first).                      */
RGBD_API int rgbd_dv_render_bwd(const rgbd_dv_params *p, const rgbd_dv_render_params *r, const float *grid,
                       const float *cam2world, const float *W1, const float *b1, const float *W2, const float *b2,
                       int B, int F, const float *saved, const float *g_novel, const float *g_depth, const float *g_fg,
                       float *g_grid, float *g_W1, float *g_b1, float *g_W2, float *g_b2, void *workspace,
                       size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RGBDGAN_B200_H */
