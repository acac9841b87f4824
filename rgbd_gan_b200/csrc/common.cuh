// common.cuh -- shared device helpers for librgbdgan_b200 (sm_100a only).
//
// Exactness contract: every quantity that feeds a discrete decision
// of the reference (truncated pixel index, in-bounds mask, occlusion mask, sign of the
// residual) is evaluated with the reference's fp32 operation order, one IEEE rounding per
// written operation.  Those operations use the __f*_rn intrinsics, which nvcc never
// contracts into FMAs; the K=3 / K=4 matrix products use explicit __fmaf_rn chains because
// that is what the reference's BLAS sgemm evaluates on its NumPy path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rgbdgan_b200.h"

namespace rgbd {

constexpr int kThreads = 256;

void set_error(const char *fmt, ...);
int check_launch(const char *what);
void count_launch(int n = 1);                       // bookkeeping for rgbd_launch_count()
extern thread_local cudaEvent_t g_hook_start, g_hook_stop;   // rgbd_profile_hook()
struct rgbd_peer_comm;
int launch_peer_collect(rgbd_peer_comm *pc, cudaStream_t st);   // consistency.cu: finishes a publish-only exchange

// ---- peer-memory mailbox for the fused loss all-reduce (one per rank, cudaMalloc + CUDA IPC)
constexpr int kMaxPeers = 16;
constexpr int kLazyDepth = 8;         // publish-only mode: ring of epochs a rank may run ahead of the slowest peer
struct rgbd_mailbox {
    float slot[2][kMaxPeers][8];      // [epoch parity][writer rank][4 loss means, depth hinge, 3 spare]
    unsigned flag[2][kMaxPeers];      // epoch published by each writer
    unsigned epoch;                   // local: last completed epoch
    unsigned error;                   // local, sticky: 1 = a wait for a peer's flag ran into the time limit, 2 = a peer
                                      // overwrote an epoch before it was collected (publish-only mode)
    unsigned lepoch;                  // local: last epoch this rank published in publish-only mode
    unsigned pad[29];
    float lslot[kLazyDepth][kMaxPeers][8];   // publish-only mode: [epoch % kLazyDepth][writer rank][values]
    unsigned lflag[kLazyDepth][kMaxPeers];   // epoch held by that slot
};
struct PeerArgs {
    rgbd_mailbox *box[kMaxPeers];     // box[r] = rank r's mailbox, mapped into this process
    int rank, world;
    int lazy;                         // 1: publish-only (rgbd_loss_opts.defer_loss == 2), summed by rgbd_peer_comm_wait
    unsigned long long timeout_ns;    // bound of every wait for a peer (a dead / desynchronised rank must not hang the GPU)
};
struct rgbd_peer_comm {               // host object behind the opaque handle of the C-ABI
    PeerArgs args;
    rgbd_mailbox *mine;
    bool connected;
    // deferred-loss mode: the finalize + exchange kernel runs on `side`, forked after the main kernel
    cudaStream_t side;
    cudaEvent_t ev_main_done, ev_fin_done;
    bool fin_pending;                 // ev_fin_done has been recorded and not yet waited on by the main stream
    unsigned long long calls;         // parity selects one of two partial-sum buffers
    // publish-only mode: what rgbd_peer_comm_wait has to finish (the latest call's output pointer and constants)
    bool lazy_pending;
    float *lazy_loss_parts;
    float lazy_lambda;
    rgbd_mailbox *loopback[kMaxPeers]; // test aid (rgbd_debug_peer_comm_loopback): local stand-ins for absent peers
};

struct Pose {          // one warp direction of one pair
    float m[9];        // K R K^-1, row major
    float c[3];        // subtracted vector
};

__device__ __forceinline__ Pose load_pose(const float *__restrict__ M, const float *__restrict__ c, int b)
{
    Pose p;
#pragma unroll
    for (int k = 0; k < 9; ++k) p.m[k] = __ldg(M + 9 * b + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) p.c[k] = __ldg(c + 3 * b + k);
    return p;
}

struct Px {
    float q0, q1, q2, zc, vcol, urow;
    int u0, v0;            // masked indices (0 where !m); v1 == v0 + 1 where m, else 0
    float a, bb, cc, d;    // unmasked 1-D weights
    float w1, w2, w3, w4;  // masked 2-D weights (0 where !m)
    bool m;
};

// coordinate part of bilinear(): common/loss_functions.py:199-225
__device__ __forceinline__ void coords_from_q(float q0, float q1, float q2, int H, int W, Px &o)
{
    o.q0 = q0; o.q1 = q1; o.q2 = q2;
    const float zc = fminf(fmaxf(q2, 1e-4f), 10000.0f);        // F.clip(zp2, 1e-4, 10000)  :199
    o.zc = zc;
    o.vcol = __fdiv_rn(q0, zc);                                 // :199 (renamed at :202)
    o.urow = __fdiv_rn(q1, zc);                                 // :200
    const int u0 = __float2int_rz(o.urow), v0 = __float2int_rz(o.vcol);   // astype(int32) :203-206
    const int u1 = (int)((unsigned)u0 + 1u), v1 = (int)((unsigned)v0 + 1u);
    o.a = __fsub_rn((float)u1, o.urow);  o.bb = __fsub_rn(o.urow, (float)u0);   // :209-212
    o.cc = __fsub_rn((float)v1, o.vcol); o.d = __fsub_rn(o.vcol, (float)v0);
    o.m = (o.urow >= 0.0f) && (o.urow < (float)(H - 1)) && (o.vcol >= 0.0f) &&
          (o.vcol < (float)(W - 1)) && (q2 > 1e-4f);            // :215-216
    if (o.m) {
        o.w1 = __fmul_rn(o.a, o.cc);  o.w2 = __fmul_rn(o.bb, o.cc);     // :222-225 (mask factor is 1)
        o.w3 = __fmul_rn(o.a, o.d);   o.w4 = __fmul_rn(o.bb, o.d);
        o.u0 = u0; o.v0 = v0;
    } else {
        o.w1 = o.w2 = o.w3 = o.w4 = 0.0f;
        o.u0 = 0; o.v0 = 0;                                      // :218-221
    }
}

// warp()/inv_warp(): common/loss_functions.py:171-182, then coords_from_q
__device__ __forceinline__ void project(const Pose &P, float z, int i, int j, int H, int W, Px &o)
{
    const float x = (float)j, y = (float)i;                     // p = (col,row,1)  :59-61
    const float P0 = __fmul_rn(z, x), P1 = __fmul_rn(z, y);     // z * p            :174
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {                               // F.matmul (sgemm, K=3): fma chain
        float t = __fmul_rn(P.m[3 * r], P0);
        t = __fmaf_rn(P.m[3 * r + 1], P1, t);
        t = __fmaf_rn(P.m[3 * r + 2], z, t);
        q[r] = __fsub_rn(t, P.c[r]);
    }
    coords_from_q(q[0], q[1], q[2], H, W, o);
}

// 4-term blend of :226-227; both row taps read row u0 (reference quirk, :219)
__device__ __forceinline__ float blend(const Px &p, float A, float Bv)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.w1, A), __fmul_rn(p.w2, A)), __fmul_rn(p.w3, Bv)),
                     __fmul_rn(p.w4, Bv));
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rgbd
