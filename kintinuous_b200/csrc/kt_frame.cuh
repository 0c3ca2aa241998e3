// kintinuous_b200 -- device helpers shared by the whole-frame (persistent, cooperative) odometry kernels:
// grid barrier, TMA bulk-copy staging, warp transpose-reduce, the per-pixel point-to-plane row.
#pragma once
#include "kt_ops.h"
#include "kt_reduce.cuh"

namespace kt {

enum { FRAME_THREADS = 512, STAGE_MAX_K = 17 };   // 17 passes x 6 planes x 2 KB = 204 KB of shared memory: 1280x960 level 0 still fits

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(*((volatile unsigned int*)bar) - target) < 0) { }
        __threadfence();
    }
    __syncthreads();
}

// ---- TMA (bulk async copy engine) helpers: global -> shared 1-D bulk copies completing on an mbarrier ----
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KT_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KT_WAIT_DONE;\n"
        "bra KT_WAIT_LOOP;\n"
        "KT_WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// 32 per-lane values -> lane l holds the warp total of value l (31 shuffles instead of 32 x 5; fixed tree)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane)
{
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int j = 0; j < step; ++j) {
            const float send = upper ? v[j] : v[j + step];
            const float keep = upper ? v[j + step] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

// one pixel whose current vertex / normal were staged in shared memory
__device__ __forceinline__ void icp_pixel_staged(const float3& vcurr, const float3& ncurr, int N, int cols, int rows,
                                                 const float* __restrict__ vmap_g_prev, const float* __restrict__ nmap_g_prev,
                                                 const Intr& intr, const Mat33& Rcurr, const float3& tcurr, const Mat33& Rprev_inv, const float3& tprev,
                                                 float dist_thres, float angle_thres, float (&sum)[32])
{
    if (isnan(vcurr.x)) return;
    float3 vcurr_g = add3(mul33(Rcurr, vcurr), tcurr);
    float3 vcurr_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));
    int2 ukr;
    ukr.x = __float2int_rn(vcurr_cp.x * intr.fx / vcurr_cp.z + intr.cx);
    ukr.y = __float2int_rn(vcurr_cp.y * intr.fy / vcurr_cp.z + intr.cy);
    if (ukr.x < 0 || ukr.y < 0 || ukr.x >= cols || ukr.y >= rows || vcurr_cp.z < 0) return;
    const int j = ukr.y * cols + ukr.x;
    float3 vprev_g, nprev_g;
    vprev_g.x = __ldg(&vmap_g_prev[j]); vprev_g.y = __ldg(&vmap_g_prev[j + N]); vprev_g.z = __ldg(&vmap_g_prev[j + 2 * N]);
    nprev_g.x = __ldg(&nmap_g_prev[j]); nprev_g.y = __ldg(&nmap_g_prev[j + N]); nprev_g.z = __ldg(&nmap_g_prev[j + 2 * N]);
    if (isnan(vprev_g.x) || isnan(nprev_g.x) || isnan(ncurr.x)) return;
    float3 ncurr_g = mul33(Rcurr, ncurr);
    float dist = norm3(sub3(vprev_g, vcurr_g));
    float sine = norm3(cross3(ncurr_g, nprev_g));
    if (!(sine < angle_thres && dist <= dist_thres)) return;
    float3 s_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));
    float3 d_cp = mul33(Rprev_inv, sub3(vprev_g, tprev));
    float3 n_cp = mul33(Rprev_inv, nprev_g);
    float3 sxn = cross3(s_cp, n_cp);
    const float row[7] = {n_cp.x, n_cp.y, n_cp.z, sxn.x, sxn.y, sxn.z, dot3(n_cp, sub3(s_cp, d_cp))};
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 7; ++b) sum[k++] += row[a] * row[b];
    sum[27] += row[6] * row[6];
    sum[28] += 1.f;
}

// The same pixel split in two so that the model-map gathers of several pixels can be in flight together (the per-iteration time of
// the whole-frame kernel is the latency of these dependent L2 loads, not their bandwidth):
//   icp_pixel_project  current vertex -> index of the model pixel it projects to, or -1      (reduce.cu:222-240)
//   icp_pixel_finish   tests + row products from the six gathered floats                     (reduce.cu:241-316)
// Expression order is the one of icp_pixel_staged, so the sums are bit-identical.
__device__ __forceinline__ int icp_pixel_project(const float3& vcurr, int cols, int rows, const Intr& intr,
                                                 const Mat33& Rcurr, const float3& tcurr, const Mat33& Rprev_inv, const float3& tprev)
{
    if (isnan(vcurr.x)) return -1;
    float3 vcurr_g = add3(mul33(Rcurr, vcurr), tcurr);
    float3 vcurr_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));
    int2 ukr;
    ukr.x = __float2int_rn(vcurr_cp.x * intr.fx / vcurr_cp.z + intr.cx);
    ukr.y = __float2int_rn(vcurr_cp.y * intr.fy / vcurr_cp.z + intr.cy);
    if (ukr.x < 0 || ukr.y < 0 || ukr.x >= cols || ukr.y >= rows || vcurr_cp.z < 0) return -1;
    return ukr.y * cols + ukr.x;
}

__device__ __forceinline__ void icp_pixel_finish(const float3& vcurr, const float3& ncurr, const float3& vprev_g, const float3& nprev_g,
                                                 const Mat33& Rcurr, const float3& tcurr, const Mat33& Rprev_inv, const float3& tprev,
                                                 float dist_thres, float angle_thres, float (&sum)[32])
{
    if (isnan(vprev_g.x) || isnan(nprev_g.x) || isnan(ncurr.x)) return;
    float3 vcurr_g = add3(mul33(Rcurr, vcurr), tcurr);
    float3 ncurr_g = mul33(Rcurr, ncurr);
    float dist = norm3(sub3(vprev_g, vcurr_g));
    float sine = norm3(cross3(ncurr_g, nprev_g));
    if (!(sine < angle_thres && dist <= dist_thres)) return;
    float3 s_cp = mul33(Rprev_inv, sub3(vcurr_g, tprev));
    float3 d_cp = mul33(Rprev_inv, sub3(vprev_g, tprev));
    float3 n_cp = mul33(Rprev_inv, nprev_g);
    float3 sxn = cross3(s_cp, n_cp);
    const float row[7] = {n_cp.x, n_cp.y, n_cp.z, sxn.x, sxn.y, sxn.z, dot3(n_cp, sub3(s_cp, d_cp))};
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 7; ++b) sum[k++] += row[a] * row[b];
    sum[27] += row[6] * row[6];
    sum[28] += 1.f;
}


} // namespace kt
