// kintinuous_b200 -- projective point-to-plane ICP: per-pixel residual / Jacobian row and the
// 6x6 [J^T J | J^T r] reduction, with the Gauss-Newton solve fused into the reduction's tail.
//
// Replaces (reference, src/frontend/):
//   icpStep + ICPReduction::search/getProducts + icpKernel + reduceSum   cuda/reduce.cu:186-419
//   struct JtJJtrSE3 (29 floats)                                         cuda/internal.h:98-149
//   the host half of ICPOdometry::getIncrementalTransformation           ICPOdometry.cpp:86-180
//
// B200 design: the reference launches 64 CTAs x 128 threads (8 192 threads on
// a 148-SM part), a second 1-CTA kernel, then cudaDeviceSynchronize + 116-byte D2H + host LDLT, 19x
// per frame.  Here ONE launch per iteration covers the image with a grid sized to the SM count
// (148 x k CTAs of 256 threads, grid-stride), reduces the 29 sums with warp shuffles -> shared memory ->
// one 128-byte partial per CTA, and the LAST CTA to finish (ticket counter) sums the partials in a
// fixed order (deterministic run to run), solves the 6x6 system in FP64 and writes the new pose into
// device memory, where the next iteration's launch picks it up: no host round trip inside a frame.
// Bound: L2-resident streaming (48 B/pixel: 24 streamed + 24 gathered) -- latency-, not HBM-bound at
// 640x480 (14.7 MB per level-0 iteration).
#include "kt_ops.h"
#include "kt_solve.cuh"
#include "kt_reduce.cuh"

namespace kt {

namespace {

enum { ICP_THREADS = RED_THREADS };

struct IcpParams {
    IcpLevelArgs a;
    OdomState* st;
    float* partials;       // [gridDim.x][32]
    float* trace;          // [iter][44] or null
    int mode;              // 0 reduce only, 1 reduce + solve (ICP-only odometry)
};

__global__ void __launch_bounds__(ICP_THREADS)
icp_kernel(const IcpParams p)
{
    __shared__ float s_pose[24];
    __shared__ float s_red[ICP_THREADS / 32][32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    if (tid < 9) { s_pose[tid] = p.st->Rcurr[tid]; s_pose[12 + tid] = p.st->Rprev_inv[tid]; }
    if (tid < 3) { s_pose[9 + tid] = p.st->tcurr[tid]; s_pose[21 + tid] = p.st->tprev[tid]; }
    __syncthreads();
    Mat33 Rcurr, Rprev_inv; float3 tcurr, tprev;
    Rcurr.r0 = make_float3(s_pose[0], s_pose[1], s_pose[2]); Rcurr.r1 = make_float3(s_pose[3], s_pose[4], s_pose[5]); Rcurr.r2 = make_float3(s_pose[6], s_pose[7], s_pose[8]);
    tcurr = make_float3(s_pose[9], s_pose[10], s_pose[11]);
    Rprev_inv.r0 = make_float3(s_pose[12], s_pose[13], s_pose[14]); Rprev_inv.r1 = make_float3(s_pose[15], s_pose[16], s_pose[17]); Rprev_inv.r2 = make_float3(s_pose[18], s_pose[19], s_pose[20]);
    tprev = make_float3(s_pose[21], s_pose[22], s_pose[23]);

    const int cols = p.a.cols, rows = p.a.rows, N = cols * rows;
    const float* __restrict__ vmap_curr = p.a.vmap_curr;
    const float* __restrict__ nmap_curr = p.a.nmap_curr;
    const float* __restrict__ vmap_g_prev = p.a.vmap_g_prev;
    const float* __restrict__ nmap_g_prev = p.a.nmap_g_prev;
    const Intr intr = p.a.k;

    float sum[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) sum[k] = 0.f;

    for (int i = blockIdx.x * ICP_THREADS + tid; i < N; i += gridDim.x * ICP_THREADS) {
        float3 vcurr;
        vcurr.x = vmap_curr[i];
        if (isnan(vcurr.x)) continue;                       // Q16: the reference rejects these through NaN propagation
        vcurr.y = vmap_curr[i + N];
        vcurr.z = vmap_curr[i + 2 * N];

        float3 vcurr_g = add3(mul33(Rcurr, vcurr), tcurr);
        float3 vcurr_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));

        int2 ukr;
        ukr.x = __float2int_rn(vcurr_cp.x * intr.fx / vcurr_cp.z + intr.cx);
        ukr.y = __float2int_rn(vcurr_cp.y * intr.fy / vcurr_cp.z + intr.cy);
        if (ukr.x < 0 || ukr.y < 0 || ukr.x >= cols || ukr.y >= rows || vcurr_cp.z < 0) continue;

        const int j = ukr.y * cols + ukr.x;
        float3 vprev_g, nprev_g, ncurr;
        vprev_g.x = __ldg(&vmap_g_prev[j]);
        nprev_g.x = __ldg(&nmap_g_prev[j]);
        ncurr.x = nmap_curr[i];
        if (isnan(vprev_g.x) || isnan(nprev_g.x) || isnan(ncurr.x)) continue;
        vprev_g.y = __ldg(&vmap_g_prev[j + N]); vprev_g.z = __ldg(&vmap_g_prev[j + 2 * N]);
        nprev_g.y = __ldg(&nmap_g_prev[j + N]); nprev_g.z = __ldg(&nmap_g_prev[j + 2 * N]);
        ncurr.y = nmap_curr[i + N]; ncurr.z = nmap_curr[i + 2 * N];

        float3 ncurr_g = mul33(Rcurr, ncurr);
        float dist = norm3(sub3(vprev_g, vcurr_g));
        float sine = norm3(cross3(ncurr_g, nprev_g));
        if (!(sine < p.a.angle_thres && dist <= p.a.dist_thres)) continue;

        float3 s_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));
        float3 d_cp = mul33(Rprev_inv, sub3(vprev_g, tprev));
        float3 n_cp = mul33(Rprev_inv, nprev_g);
        float3 sxn = cross3(s_cp, n_cp);
        float row[7] = {n_cp.x, n_cp.y, n_cp.z, sxn.x, sxn.y, sxn.z, dot3(n_cp, sub3(s_cp, d_cp))};
        accumulate_row(sum, row);
    }

    if (!grid_reduce29(sum, p.partials, &p.st->blocks_done, s_red, &s_last)) return;

    // ---- last CTA: the Gauss-Newton step ----
    if (tid < NSUM) p.st->sums_icp[tid] = s_red[0][tid];
    if (tid == 0) {
        OdomState* st = p.st;
        float A[36], b[6];
        unpack_normal_equations(s_red[0], A, b);
        if (p.trace) {
            float* t = p.trace + (size_t)st->iter * TRACE_STRIDE;
            for (int k = 0; k < 36; ++k) t[k] = A[k];
            for (int k = 0; k < 6; ++k) t[36 + k] = b[k];
            t[42] = s_red[0][27]; t[43] = s_red[0][28];
        }
        if (p.mode == 1) {
            double dA[36], db[6];
            for (int k = 0; k < 36; ++k) dA[k] = A[k];
            for (int k = 0; k < 6; ++k) db[k] = b[k];
            gauss_newton_update(dA, db, st);
            st->iter += 1;
        }
    }
}

__global__ void odom_begin_kernel(OdomState* st, const float* pose12)
{
    // pose12: Rprev (9) tprev (3), uploaded by the host (it owns rmats_/tvecs_ like the reference tracker)
    if (threadIdx.x == 0) {
        for (int k = 0; k < 9; ++k) { st->Rprev[k] = pose12[k]; st->Rcurr[k] = pose12[k]; }
        for (int k = 0; k < 3; ++k) { st->tprev[k] = pose12[9 + k]; st->tcurr[k] = pose12[9 + k]; }
        mat3f_inverse(st->Rprev, st->Rprev_inv);                       // Rprev.inverse(), ICPOdometry.cpp:81
        for (int k = 0; k < 16; ++k) st->resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
        st->iter = 0; st->blocks_done = 0; st->blocks_done_rgb = 0;
        st->rgb_count = 0; st->rgb_sigma = 0;
    }
}

} // namespace

static int g_sm_count = 0;
static int sm_count()
{
    if (!g_sm_count) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

int reduce_grid_for(int n_items)
{
    // two items per thread are enough to cover latency; cap at 4 CTAs per SM (a multiple of the SM count) and MAX_PARTIALS
    int want = div_up(n_items, RED_THREADS * 2);
    int cap = sm_count() * 4;
    if (cap > MAX_PARTIALS) cap = MAX_PARTIALS;
    return want < cap ? (want < 1 ? 1 : want) : cap;
}

int icp_iteration(const IcpLevelArgs& a, OdomState* state, float* partials, float* trace, int mode, cudaStream_t s)
{
    IcpParams p; p.a = a; p.st = state; p.partials = partials; p.trace = trace; p.mode = mode;
    int grid = reduce_grid_for(a.rows * a.cols);
    icp_kernel<<<grid, ICP_THREADS, 0, s>>>(p);
    KT_LAUNCH_CHECK();
    return 0;
}

int odom_begin_frame(OdomState* state, const float* pose12_dev, cudaStream_t s)
{
    odom_begin_kernel<<<1, 32, 0, s>>>(state, pose12_dev);
    KT_LAUNCH_CHECK();
    return 0;
}

} // namespace kt
