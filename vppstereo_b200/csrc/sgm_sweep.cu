// sgm_sweep.cu -- the 8-path SGM aggregation of compute_rsgm as FOUR sweeps with on-chip path state (sm_100a),
// bit-exact with the reference's raster recursion (RSGM/StereoSGM_SSE.hpp:13-515 via RSGM/pyrSGM.cpp:504-637;
// semantics in SURVEY.md A.4, restated per path line in sgm.cu).
//
// Why.  sgm.cu runs one launch per path (8) and every launch read-modify-writes the aggregated volume S in HBM:
// 8x the algorithmic bytes (profiles/r01_summary_v1.md).  Here S is touched once per sweep:
//   h-sweep fwd  : r0 of pass 0, one warp per image row, path state in registers, S  = L          (store)
//   v-sweep down : r1+r2+r3 of pass 0 together, S += L1+L2+L3                                      (one RMW)
//   v-sweep up   : r1+r2+r3 of pass 1 together, S += L1+L2+L3                                      (one RMW)
//   h-sweep bwd  : r0 of pass 1, S += L  (stand-alone), or fused into the WTA row sweep (rsgm_ops.cu) so that the
//                  final S is never written
// v-sweep: a thread-block CLUSTER owns one frame; CTA c owns a strip of columns and keeps the three paths' previous-row
// state L_r(.,d) for its strip in shared memory (3 x strip x D x 2 B, ~180 KB at D=192 / 156 columns).  Rows are swept
// in order; a warp handles one pixel at a time (lane l holds disparities [2*NW*l, 2*NW*(l+1)) as u16x2 words).
// Diagonal state is stored per LINE in a ring (slot = (column -/+ row) mod strip) so a line's state never moves;
// only the line that leaves the strip is pushed into the neighbour CTA's halo through distributed shared memory,
// one cluster barrier per row (arrive after the row, wait before the next one).
//
// Arithmetic: the "fast" domain of sgm.cu -- uint8 costs, the effective default parameters (P1=7, P2min=17,
// Alpha=0.25, Gamma=50 => P2 in [17,50]), so no uint16 saturation is reachable (L <= 255+50, S <= 8*305).
// State is kept NORMALISED (L - min_d L): L_new = C + min(L'[d], min(L'[d-1], L'[d+1]) + P1, P2) is the reference's
// C + min(L[d], L[d+-1]+P1, minL+P2) - minL.
//
// Internal "plane" layouts (per pixel, NW = ceil(D/64) words per lane): word index k*32 + lane holds disparities
// d = 2*NW*lane + 2k (low half) and d+1 (high half); costs 1 byte per disparity (u16 words), S 2 bytes (u32 words).
// Every warp-wide access is one contiguous 64 B / 128 B segment.  For D = 64*NW this is a permutation of the
// reference's xyd order inside a pixel; vppb200 un-permutes it for the test tap only.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace vppb200 {

#define SW_BIG2 0x3FFF3FFFu
static constexpr int SW_P1 = 7, SW_P2MIN = 17, SW_GAMMA = 50;
static constexpr float SW_ALPHA = 0.25f;
static constexpr uint32_t SW_P1X2 = 0x00070007u;

__device__ __forceinline__ int sw_adapt_p2(int ip, int ipr)
{
    // (sint32)(-alpha * abs(I_p - I_pr) + gamma), clamped below by P2min  (RSGM/StereoSGM.hpp:92-99)
    const int r = (int)__fadd_rn(__fmul_rn(-SW_ALPHA, (float)abs(ip - ipr)), (float)SW_GAMMA);
    return r < SW_P2MIN ? SW_P2MIN : r;
}

// one SGM step on normalised state: nw = c + min(w[d], min(w[d-1], w[d+1]) + P1, P2);  p2m = (P2 - P1) * 0x10001
template <int NW>
__device__ __forceinline__ void sw_step(const uint32_t (&w)[NW], const uint32_t (&c)[NW], uint32_t p2m, int lane,
                                        uint32_t (&nw)[NW])
{
    uint32_t up = __shfl_up_sync(0xFFFFFFFFu, w[NW - 1], 1);
    uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, w[0], 1);
    if (lane == 0) up = SW_BIG2;
    if (lane == 31) dn = SW_BIG2;
    uint32_t ext[NW + 2];
    ext[0] = up;
#pragma unroll
    for (int k = 0; k < NW; k++) ext[k + 1] = w[k];
    ext[NW + 1] = dn;
    uint32_t p[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; k++) p[k] = __byte_perm(ext[k], ext[k + 1], 0x5432);   // (hi(ext[k]), lo(ext[k+1]))
#pragma unroll
    for (int k = 0; k < NW; k++) {
        uint32_t t = __vimin3_u16x2(p[k], p[k + 1], p2m);       // min(L[d-1], L[d+1], P2 - P1)
        t = __viaddmin_u16x2(t, SW_P1X2, w[k]);                 // min(. + P1, L[d])
        nw[k] = t + c[k];
    }
}

// min over all disparities of a warp's packed values
template <int NW>
__device__ __forceinline__ uint32_t sw_min(const uint32_t (&v)[NW])
{
    uint32_t mm = v[0];
#pragma unroll
    for (int k = 1; k < NW; k++) mm = __vminu2(mm, v[k]);
    mm = min(mm & 0xFFFFu, mm >> 16);
    return __reduce_min_sync(0xFFFFFFFFu, mm);
}

// ------------------------------------------------------------------------------------------------------------
// cost volume in plane layout: popc(L ^ R[x-d]) for d <= x on rows 2..H-3, 12 elsewhere (RSGM/StereoBMHelper.cpp:29-140)
// One warp produces PX adjacent pixels of a row; lane l owns disparities [2*NW*l, 2*NW*(l+1)) and keeps the
// PX + 2*NW - 1 right-census words it needs in registers.
// ------------------------------------------------------------------------------------------------------------
template <int NW, int PX>
__global__ void __launch_bounds__(256) cost_plane_kernel(const uint32_t *__restrict__ cl, const uint32_t *__restrict__ cr,
                                                         uint16_t *__restrict__ cost, int W, int H, int D, long total_groups)
{
    const long grp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (grp >= total_groups) return;
    const int lane = threadIdx.x & 31;
    const int gpr = W / PX;                        // W % 16 == 0, PX divides 16
    const long row = grp / gpr;                    // row over all frames
    const int x0 = (int)(grp % gpr) * PX;
    const int y = (int)(row % H);
    const int d0 = 2 * NW * lane;
    uint16_t *out = cost + (row * W + x0) * (long)(NW * 32) + lane;
    if (y < 2 || y >= H - 2) {
#pragma unroll
        for (int p = 0; p < PX; p++)
#pragma unroll
            for (int k = 0; k < NW; k++) out[(p * NW + k) * 32] = 0x0C0Cu;
        return;
    }
    const uint32_t *lrow = cl + row * W, *rrow = cr + row * W;
    // window: rr[t] = R[x0 - d0 - (2NW-1) + t], t in [0, PX + 2NW - 1)
    uint32_t rr[PX + 2 * NW - 1];
    const int base = x0 - d0 - (2 * NW - 1);
#pragma unroll
    for (int t = 0; t < PX + 2 * NW - 1; t++) rr[t] = (base + t >= 0) ? rrow[base + t] : 0u;
#pragma unroll
    for (int p = 0; p < PX; p++) {
        const int x = x0 + p;
        const uint32_t l = lrow[x];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            // d = d0 + 2k (+1): R[x - d] = rr[p + 2NW-1 - 2k (-1)]
            const int dlo = d0 + 2 * k;
            uint32_t vlo = dlo > x ? 12u : (uint32_t)__popc(l ^ rr[p + 2 * NW - 1 - 2 * k]);
            uint32_t vhi = dlo + 1 > x ? 12u : (uint32_t)__popc(l ^ rr[p + 2 * NW - 2 - 2 * k]);
            if (dlo >= D) { vlo = 0; vhi = 0; }
            out[(p * NW + k) * 32] = (uint16_t)(vlo | (vhi << 8));
        }
    }
}

int launch_cost_plane(const uint32_t *cl, const uint32_t *cr, uint8_t *cost, int W, int H, int D, int n, cudaStream_t st)
{
    constexpr int PX = 8;
    const long groups = (long)n * H * (W / PX);
    const int blocks = cdiv(groups * 32, 256);
    uint16_t *c16 = reinterpret_cast<uint16_t *>(cost);
    switch ((D + 63) / 64) {
        case 1: cost_plane_kernel<1, PX><<<blocks, 256, 0, st>>>(cl, cr, c16, W, H, D, groups); break;
        case 2: cost_plane_kernel<2, PX><<<blocks, 256, 0, st>>>(cl, cr, c16, W, H, D, groups); break;
        case 3: cost_plane_kernel<3, PX><<<blocks, 256, 0, st>>>(cl, cr, c16, W, H, D, groups); break;
        default: cost_plane_kernel<4, PX><<<blocks, 256, 0, st>>>(cl, cr, c16, W, H, D, groups); break;
    }
    VPP_LAUNCH_CHECK("cost_plane_kernel");
    return VPPB200_OK;
}

// _guided_dsi (models/rsgm/rsgm.py:115-127) on the plane-layout u8 volume; arithmetic as guided_u8_kernel (rsgm_ops.cu)
__global__ void guided_plane_kernel(uint8_t *__restrict__ cost, const float *__restrict__ hints, const float *__restrict__ valid,
                                    RsgmDims d, int NW, long total)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int dd = (int)(t % d.D);
    const long px = t / d.D;
    const int x = (int)(px % d.Wp) - d.pl, y = (int)((px / d.Wp) % d.Hp) - d.pt;
    const long f = px / ((long)d.Wp * d.Hp);
    if (x < 0 || x >= d.W || y < 0 || y >= d.H) return;
    const long s = (f * d.H + y) * d.W + x;
    if (!(valid[s] > 0)) return;
    const double tt = __dsub_rn((double)hints[s], (double)dd);
    const float w = (float)__dmul_rn(10.0, __dsub_rn(1.0, exp(__ddiv_rn(-__dmul_rn(tt, tt), 2.0))));
    const int lane = dd / (2 * NW), k = (dd % (2 * NW)) >> 1;
    uint8_t *p = cost + px * (long)(NW * 64) + (k * 32 + lane) * 2 + (dd & 1);
    *p = (uint8_t)(uint16_t)__dmul_rn((double)*p, (double)w);
}
int launch_guided_plane(uint8_t *cost, const float *hints, const float *valid, const RsgmDims &d, int n, cudaStream_t st)
{
    long total = (long)n * d.Hp * d.Wp * d.D;
    guided_plane_kernel<<<cdiv(total, 256), 256, 0, st>>>(cost, hints, valid, d, (d.D + 63) / 64, total);
    VPP_LAUNCH_CHECK("guided_plane_kernel");
    return VPPB200_OK;
}

// plane-layout S -> the reference's xyd order (test tap only)
__global__ void unplane_s_kernel(const uint32_t *__restrict__ Sp, uint16_t *__restrict__ S, int D, int NW, long total)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;        // over pixels * D/2
    if (t >= total) return;
    const int h = D / 2;
    const int dp = (int)(t % h) * 2;
    const long px = t / h;
    const int lane = dp / (2 * NW), k = (dp % (2 * NW)) >> 1;
    reinterpret_cast<uint32_t *>(S)[t] = Sp[px * (long)(NW * 32) + k * 32 + lane];
}
int launch_unplane_s(const uint32_t *Sp, uint16_t *S, int W, int H, int D, int n, cudaStream_t st)
{
    long total = (long)n * W * H * (D / 2);
    unplane_s_kernel<<<cdiv(total, 256), 256, 0, st>>>(Sp, S, D, (D + 63) / 64, total);
    VPP_LAUNCH_CHECK("unplane_s_kernel");
    return VPPB200_OK;
}

// ------------------------------------------------------------------------------------------------------------
// per-warp operand ring: the next pixels' costs (and S words) are brought into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) several steps ahead of the warp that consumes them, so that no sweep step
// waits on HBM latency.  One lane produces, the whole warp consumes; a slot is refilled right after it was read.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}

// ------------------------------------------------------------------------------------------------------------
// h-sweep: path r0 of one pass, one warp per image row (RSGM/StereoSGM_SSE.hpp:100-113,:219-236 for the recursion;
// the line starts with L = C at the pass's first column and every pixel is summed into S).
// Operands arrive through the ring in chunks of HCH pixels; P2 of 32 consecutive pixels is computed lane-parallel one
// block ahead and broadcast per step.
// ------------------------------------------------------------------------------------------------------------
static constexpr int HCH = 4;      // pixels per ring slot
static constexpr int HST = 2;      // ring slots per warp
static constexpr int HWARPS = 4;   // warps per CTA

template <int NW, bool STORE>
__host__ __device__ constexpr int h_slot_bytes() { return HCH * NW * 32 * (STORE ? 2 : 6); }

template <int NW, bool PAD, bool STORE>
__global__ void __launch_bounds__(HWARPS * 32) sgm_h_kernel(const uint8_t *__restrict__ img, const uint16_t *__restrict__ cost,
                                                            uint32_t *__restrict__ S, int W, int D, int dirn, long total_rows)
{
    extern __shared__ __align__(16) uint8_t hsm[];
    constexpr int PW = NW * 32;
    constexpr int SLOTB = h_slot_bytes<NW, STORE>();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * HWARPS + warp;
    if (row >= total_rows) return;
    uint8_t *ring = hsm + warp * (HST * SLOTB);
    const uint32_t ring_a = smem_u32(ring);
    const uint32_t bars = smem_u32(hsm + HWARPS * HST * SLOTB) + warp * (HST * 8);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < HST; q++) mbar_init(bars + q * 8, 1);
        mbar_init_fence();
    }
    __syncwarp();
    const uint8_t *irow = img + row * W;
    const uint16_t *crow = cost + row * (long)W * PW;
    uint32_t *srow = S + row * (long)W * PW;
    bool wv[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) wv[k] = !PAD || (2 * NW * lane + 2 * k < D);
    const int nchunks = W / HCH;
    const int xs = dirn > 0 ? 0 : W - 1;
    // chunk c covers pixels [xlo, xlo + HCH): forwards xlo = c*HCH, backwards xlo = W - (c+1)*HCH
    auto produce = [&](int c) {
        if (lane == 0 && c < nchunks) {
            const int xlo = dirn > 0 ? c * HCH : W - (c + 1) * HCH;
            const uint32_t slot = (uint32_t)c % HST;
            const uint32_t dst = ring_a + slot * SLOTB, bar = bars + slot * 8;
            mbar_expect_tx(bar, SLOTB);
            tma_load_1d(dst, crow + (long)xlo * PW, HCH * PW * 2, bar);
            if (!STORE) tma_load_1d(dst + HCH * PW * 2, srow + (long)xlo * PW, HCH * PW * 4, bar);
        }
    };
#pragma unroll
    for (int q = 0; q < HST; q++) produce(q);

    auto p2_block = [&](int t0) -> uint32_t {
        // (P2 - P1) * 0x10001 for step t0 + lane: |I(x) - I(x - dirn)| inside the row (step 0 has no predecessor)
        const int t = t0 + lane;
        int p2 = SW_P2MIN;
        if (t > 0 && t < W) {
            const int xx = xs + dirn * t;
            p2 = sw_adapt_p2(irow[xx], irow[xx - dirn]);
        }
        return (uint32_t)(p2 - SW_P1) * 0x10001u;
    };
    uint32_t p2v = p2_block(0), p2n = p2_block(32);
    uint32_t w[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = 0;
    int t = 0;
    for (int c = 0; c < nchunks; c++) {
        const uint32_t slot = (uint32_t)c % HST;
        mbar_wait(bars + slot * 8, ((uint32_t)c / HST) & 1u);
        const uint16_t *rc = reinterpret_cast<const uint16_t *>(ring + slot * SLOTB) + lane;
        const uint32_t *rs = reinterpret_cast<const uint32_t *>(ring + slot * SLOTB + HCH * PW * 2) + lane;
        const int xlo = dirn > 0 ? c * HCH : W - (c + 1) * HCH;
#pragma unroll
        for (int pp = 0; pp < HCH; pp++, t++) {
            const int p = dirn > 0 ? pp : HCH - 1 - pp;
            if ((t & 31) == 0 && t > 0) { p2v = p2n; p2n = p2_block(t + 32); }
            uint32_t cc[NW], sv[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
                cc[k] = __byte_perm((uint32_t)rc[p * PW + k * 32], 0, 0x4140);      // two uint8 costs -> u16x2
                if (!STORE) sv[k] = rs[p * PW + k * 32];
            }
            const uint32_t p2m = __shfl_sync(0xFFFFFFFFu, p2v, t & 31);
            uint32_t nw[NW];
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < NW; k++) nw[k] = cc[k];
            } else {
                sw_step<NW>(w, cc, p2m, lane, nw);
            }
            if (PAD) {
#pragma unroll
                for (int k = 0; k < NW; k++) if (!wv[k]) nw[k] = SW_BIG2;
            }
            const uint32_t m2 = sw_min<NW>(nw) * 0x10001u;
            uint32_t *so = srow + (long)(xlo + p) * PW + lane;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                w[k] = nw[k] - m2;
                if (wv[k]) so[k * 32] = STORE ? nw[k] : sv[k] + nw[k];
            }
        }
        __syncwarp();
        produce(c + HST);
    }
}

template <int NW, bool PAD, bool STORE>
static int run_h_t(const uint8_t *img, const uint16_t *cost, uint32_t *S, int W, int H, int D, int dirn, int n, cudaStream_t st)
{
    const long rows = (long)n * H;
    const int blocks = cdiv(rows, HWARPS);
    const size_t smem = (size_t)HWARPS * HST * h_slot_bytes<NW, STORE>() + HWARPS * HST * 8;
    auto kern = sgm_h_kernel<NW, PAD, STORE>;
    if (smem > 48 * 1024) VPP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, HWARPS * 32, smem, st>>>(img, cost, S, W, D, dirn, rows);
    VPP_LAUNCH_CHECK("sgm_h_kernel");
    return VPPB200_OK;
}

template <int NW>
static int run_h(const uint8_t *img, const uint16_t *cost, uint32_t *S, int W, int H, int D, int dirn, bool store, int n,
                 cudaStream_t st)
{
    const bool pad = (D != 64 * NW);
    if (store) return pad ? run_h_t<NW, true, true>(img, cost, S, W, H, D, dirn, n, st)
                          : run_h_t<NW, false, true>(img, cost, S, W, H, D, dirn, n, st);
    return pad ? run_h_t<NW, true, false>(img, cost, S, W, H, D, dirn, n, st)
               : run_h_t<NW, false, false>(img, cost, S, W, H, D, dirn, n, st);
}

// ------------------------------------------------------------------------------------------------------------
// v-sweep: paths r1, r2, r3 of one pass, cluster per frame, state in (distributed) shared memory
// ------------------------------------------------------------------------------------------------------------
struct VArgs {
    int W, H, D, n;
    int pass;          // 0: top-down (di = dj = +1), 1: bottom-up (di = dj = -1)
    int csize;         // CTAs per cluster
    int SC;            // strip width (columns per CTA), ceil(W / csize)
};

static constexpr int SWV_WARPS = 16;
static constexpr int VRING = 4;    // operand ring slots (pixels in flight) per warp

template <int NW, bool PAD>
__global__ void __launch_bounds__(SWV_WARPS * 32, 1) sgm_v_kernel(const uint8_t *__restrict__ img_all,
                                                                  const uint16_t *__restrict__ cost_all,
                                                                  uint32_t *__restrict__ S_all, VArgs a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / a.csize, nclusters = gridDim.x / a.csize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PW = NW * 32;                    // words per pixel-path state
    constexpr int SLOTB = 6 * PW;                  // ring slot: PW u16 costs + PW u32 S words
    const int W = a.W, H = a.H;
    const int x0 = rank * a.SC;
    const int nc = min(a.SC, W - x0);              // columns of this strip (>= 1, checked by the launcher)
    const int dj = a.pass == 0 ? 1 : -1, di = dj;
    const int i1 = a.pass == 0 ? 0 : H - 1;

    uint32_t *st = smem;                            // [3][SC][PW]
    uint32_t *halo1 = st + 3 * a.SC * PW;           // [2][PW]  r1 state of the line entering this strip
    uint32_t *halo3 = halo1 + 2 * PW;               // [2][PW]  r3 state of the line entering this strip
    uint4 *p2tab = reinterpret_cast<uint4 *>(halo3 + 2 * PW);   // [2][SC]  (P2 - P1) * 0x10001 for r1, r2, r3
    uint8_t *ring = reinterpret_cast<uint8_t *>(p2tab + 2 * a.SC) + warp * (VRING * SLOTB);
    const uint32_t ring_a = smem_u32(ring);
    const uint32_t bars = smem_u32(reinterpret_cast<uint8_t *>(p2tab + 2 * a.SC) + SWV_WARPS * VRING * SLOTB) + warp * (VRING * 8);
    // r1 lines move by +dj per row, r3 lines by -dj: where a leaving line's state goes
    uint32_t *push1 = (rank + dj >= 0 && rank + dj < a.csize) ? cluster.map_shared_rank(halo1, rank + dj) : nullptr;
    uint32_t *push3 = (rank - dj >= 0 && rank - dj < a.csize) ? cluster.map_shared_rank(halo3, rank - dj) : nullptr;

    bool wv[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) wv[k] = !PAD || (2 * NW * lane + 2 * k < a.D);

    const long npx = (long)W * H;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < VRING; q++) mbar_init(bars + q * 8, 1);
        mbar_init_fence();
    }
    __syncwarp();
    // this warp's pixel stream: frames f = cid, cid + nclusters, ...; rows in sweep order; columns warp, warp + 16, ...
    int pf = cid, ps = 0, pjl = warp;               // producer position
    unsigned pq = 0, cq = 0;                        // pixels produced / consumed
    auto produce = [&]() {
        if (pf < a.n) {
            if (lane == 0) {
                const long px = (long)pf * npx + (long)(i1 + ps * di) * W + x0 + pjl;
                const uint32_t slot = pq % VRING;
                const uint32_t dst = ring_a + slot * SLOTB, bar = bars + slot * 8;
                mbar_expect_tx(bar, SLOTB);
                tma_load_1d(dst, cost_all + px * PW, 2 * PW, bar);
                tma_load_1d(dst + 2 * PW, S_all + px * PW, 4 * PW, bar);
            }
            pq++;
            pjl += SWV_WARPS;
            if (pjl >= nc) {
                pjl = warp;
                if (++ps == H) { ps = 0; pf += nclusters; }
            }
        }
    };
    if (warp < nc) {
#pragma unroll
        for (int q = 0; q < VRING; q++) produce();
    }

    unsigned gstep = 0;                             // rows processed by this cluster so far (halo / P2 double-buffer parity)
    for (int f = cid; f < a.n; f += nclusters) {
        const uint8_t *img = img_all + f * npx;
        uint32_t *S = S_all + f * npx * PW + lane;
        int sh = 0;                                 // s mod nc
        for (int s = 0; s < H; s++, gstep++) {
            const int i = i1 + s * di;
            if (gstep > 0) cluster.barrier_wait();  // row s-1 (state, halos, P2 table) complete everywhere
            const unsigned par = gstep & 1u;
            const uint4 *p2row = p2tab + par * a.SC;
            const uint32_t *h1in = halo1 + (par ^ 1u) * PW, *h3in = halo3 + (par ^ 1u) * PW;

            for (int jl = warp; jl < nc; jl += SWV_WARPS) {
                const uint32_t slot = cq % VRING;
                mbar_wait(bars + slot * 8, (cq / VRING) & 1u);
                cq++;
                const uint16_t *rc = reinterpret_cast<const uint16_t *>(ring + slot * SLOTB) + lane;
                const uint32_t *rs = reinterpret_cast<const uint32_t *>(ring + slot * SLOTB + 2 * PW) + lane;
                uint32_t c[NW], sv[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    c[k] = __byte_perm((uint32_t)rc[k * 32], 0, 0x4140);
                    sv[k] = rs[k * 32];
                }
                __syncwarp();
                produce();                          // refill the slot that was just read
                const int j = x0 + jl;
                // ring slots: a line moving +1 column per row sits in slot (jl - s) mod nc, one moving -1 in (jl + s) mod nc
                int slotA = jl - sh; if (slotA < 0) slotA += nc;
                int slotB = jl + sh; if (slotB >= nc) slotB -= nc;
                const int slot1 = dj > 0 ? slotA : slotB;
                const int slot3 = dj > 0 ? slotB : slotA;
                uint32_t *s1 = st + (0 * a.SC + slot1) * PW + lane;
                uint32_t *s2 = st + (1 * a.SC + jl) * PW + lane;
                uint32_t *s3 = st + (2 * a.SC + slot3) * PW + lane;
                uint32_t n1[NW], n2[NW], n3[NW];
                if (s == 0) {
                    // first row of the pass: L = C on all three paths, nothing is summed (StereoSGM_SSE.hpp:116-218)
#pragma unroll
                    for (int k = 0; k < NW; k++) n1[k] = n2[k] = n3[k] = c[k];
                } else {
                    const uint4 pm = p2row[jl];
                    uint32_t w[NW];
                    // r2: predecessor (i - di, j)
#pragma unroll
                    for (int k = 0; k < NW; k++) w[k] = s2[k * 32];
                    sw_step<NW>(w, c, pm.y, lane, n2);
                    // r1: predecessor (i - di, j - dj)
                    const int jp1 = j - dj, jlp1 = jl - dj;
                    if (jp1 < 0 || jp1 >= W) {
                        // enters through the image border: min(65535, 65535 + P1, 0 + P2) - 0 = P2  (:48-58,:69-72)
#pragma unroll
                        for (int k = 0; k < NW; k++) n1[k] = c[k] + pm.x + SW_P1X2;
                    } else {
                        const uint32_t *src = (jlp1 >= 0 && jlp1 < nc) ? s1 : h1in + lane;
#pragma unroll
                        for (int k = 0; k < NW; k++) w[k] = src[k * 32];
                        sw_step<NW>(w, c, pm.x, lane, n1);
                    }
                    // r3: predecessor (i - di, j + dj)
                    const int jp3 = j + dj, jlp3 = jl + dj;
                    if (jp3 < 0 || jp3 >= W) {
#pragma unroll
                        for (int k = 0; k < NW; k++) n3[k] = c[k] + pm.z + SW_P1X2;
                    } else {
                        const uint32_t *src = (jlp3 >= 0 && jlp3 < nc) ? s3 : h3in + lane;
#pragma unroll
                        for (int k = 0; k < NW; k++) w[k] = src[k * 32];
                        sw_step<NW>(w, c, pm.z, lane, n3);
                    }
                    const int o = (i * W + j) * PW;
#pragma unroll
                    for (int k = 0; k < NW; k++)
                        if (wv[k]) S[o + k * 32] = sv[k] + n1[k] + n2[k] + n3[k];
                }
                if (PAD) {
#pragma unroll
                    for (int k = 0; k < NW; k++) if (!wv[k]) { n1[k] = SW_BIG2; n2[k] = SW_BIG2; n3[k] = SW_BIG2; }
                }
                // normalise, keep as the state of this row; push a line that leaves the strip to the neighbour's halo
                const uint32_t m1 = sw_min<NW>(n1) * 0x10001u;
                const uint32_t m2 = (s == 0) ? m1 : sw_min<NW>(n2) * 0x10001u;
                const uint32_t m3 = (s == 0) ? m1 : sw_min<NW>(n3) * 0x10001u;
                const bool leave1 = (jl + dj < 0 || jl + dj >= nc) && push1 != nullptr;
                const bool leave3 = (jl - dj < 0 || jl - dj >= nc) && push3 != nullptr;
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint32_t v1 = n1[k] - m1, v2 = n2[k] - m2, v3 = n3[k] - m3;
                    s1[k * 32] = v1;
                    s2[k * 32] = v2;
                    s3[k * 32] = v3;
                    if (leave1) push1[par * PW + k * 32 + lane] = v1;
                    if (leave3) push3[par * PW + k * 32 + lane] = v3;
                }
            }
            // P2 of the next row for this strip: intensities from the FLAT image stream (wrap across row ends); on the
            // row right after the pass's first row the "previous line" is that same row (StereoSGM_SSE.hpp:221,:238-243)
            if (s + 1 < H) {
                const int in = i + di;
                const int il = (s == 0) ? in : i;
                uint4 *p2next = p2tab + (par ^ 1u) * a.SC;
                for (int t = tid; t < nc; t += SWV_WARPS * 32) {
                    const int j = x0 + t;
                    const int ip = img[in * W + j];
                    long q1 = (long)il * W + j - dj, q2 = (long)il * W + j, q3 = (long)il * W + j + dj;
                    q1 = q1 < 0 ? 0 : (q1 >= npx ? npx - 1 : q1);
                    q3 = q3 < 0 ? 0 : (q3 >= npx ? npx - 1 : q3);
                    uint4 e;
                    e.x = (uint32_t)(sw_adapt_p2(ip, img[q1]) - SW_P1) * 0x10001u;
                    e.y = (uint32_t)(sw_adapt_p2(ip, img[q2]) - SW_P1) * 0x10001u;
                    e.z = (uint32_t)(sw_adapt_p2(ip, img[q3]) - SW_P1) * 0x10001u;
                    e.w = 0;
                    p2next[t] = e;
                }
            }
            if (++sh == nc) sh = 0;
            cluster.barrier_arrive();
        }
    }
    if (gstep > 0) cluster.barrier_wait();          // nobody leaves while a neighbour may still push into its halo
}

static size_t v_fixed_bytes(int NW) { return (size_t)4 * NW * 32 * 4 + (size_t)SWV_WARPS * VRING * (6 * NW * 32 + 8); }
static size_t v_smem_bytes(int NW, int SC) { return (size_t)3 * SC * NW * 32 * 4 + (size_t)2 * SC * 16 + v_fixed_bytes(NW); }

// tuning / test hook: upper bound on the strip width (0 = as wide as shared memory allows)
static int g_max_strip = 0;
void sweep_set_max_strip(int cols) { g_max_strip = cols < 0 ? 0 : cols; }

struct VPlan { int csize, SC, nclusters; size_t smem; };

template <int NW, bool PAD>
static int plan_v(int W, int n, VPlan *plan)
{
    auto kern = sgm_v_kernel<NW, PAD>;
    int dev = 0, smem_optin = 0;
    VPP_CUDA_TRY(cudaGetDevice(&dev));
    VPP_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if ((size_t)smem_optin <= v_fixed_bytes(NW)) return 1;
    int sc_max = (int)(((size_t)smem_optin - v_fixed_bytes(NW)) / ((size_t)3 * NW * 32 * 4 + 32));
    if (g_max_strip > 0 && g_max_strip < sc_max) sc_max = g_max_strip;
    if (sc_max < 1) return VPPB200_ERR_ARG;
    int csize = 1;
    while (csize * sc_max < W && csize < 16) csize *= 2;
    if (csize * sc_max < W) return 1;               // does not fit a cluster: caller falls back to the per-path kernels
    const int SC = (W + csize - 1) / csize;
    if ((csize - 1) * SC >= W) return 1;            // an empty strip
    const size_t smem = v_smem_bytes(NW, SC);
    VPP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) VPP_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(csize * n));
    cfg.blockDim = dim3(SWV_WARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) { cudaGetLastError(); return 1; }
    plan->csize = csize; plan->SC = SC; plan->smem = smem;
    plan->nclusters = n < max_clusters ? n : max_clusters;
    return VPPB200_OK;
}

template <int NW, bool PAD>
static int run_v(const uint8_t *img, const uint16_t *cost, uint32_t *S, int W, int H, int D, int pass, int n, const VPlan &p,
                 cudaStream_t st)
{
    VArgs a;
    a.W = W; a.H = H; a.D = D; a.n = n; a.pass = pass; a.csize = p.csize; a.SC = p.SC;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.csize * p.nclusters));
    cfg.blockDim = dim3(SWV_WARPS * 32);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VPP_CUDA_TRY(cudaLaunchKernelEx(&cfg, sgm_v_kernel<NW, PAD>, img, cost, S, a));
    note_launch();
    return VPPB200_OK;
}

template <int NW, bool PAD>
static int aggregate_plane_t(const uint8_t *img, const uint8_t *cost8, uint16_t *S16, int W, int H, int D, int n, int h_bwd,
                             cudaStream_t st)
{
    VPlan plan;
    int rc = plan_v<NW, PAD>(W, n, &plan);
    if (rc) return rc;
    const uint16_t *cost = reinterpret_cast<const uint16_t *>(cost8);
    uint32_t *S = reinterpret_cast<uint32_t *>(S16);
    if ((rc = run_h<NW>(img, cost, S, W, H, D, +1, true, n, st))) return rc;
    if ((rc = run_v<NW, PAD>(img, cost, S, W, H, D, 0, n, plan, st))) return rc;
    if ((rc = run_v<NW, PAD>(img, cost, S, W, H, D, 1, n, plan, st))) return rc;
    if (h_bwd && (rc = run_h<NW>(img, cost, S, W, H, D, -1, false, n, st))) return rc;
    return VPPB200_OK;
}

// does the cluster sweep cover this shape on the current device?  (strip state must fit one cluster's shared memory)
static int g_sweep_off = 0;
void sweep_set_enabled(int on) { g_sweep_off = !on; }
bool aggregate_plane_supported(int W, int H, int D, int n)
{
    if (g_sweep_off || (long)W * H * 128 >= (1L << 31) || H < 3) return false;
    const int nw = (D + 63) / 64;
    const bool pad = (D != 64 * nw);
    VPlan plan;
    int rc;
    switch (nw) {
        case 1: rc = pad ? plan_v<1, true>(W, n, &plan) : plan_v<1, false>(W, n, &plan); break;
        case 2: rc = pad ? plan_v<2, true>(W, n, &plan) : plan_v<2, false>(W, n, &plan); break;
        case 3: rc = pad ? plan_v<3, true>(W, n, &plan) : plan_v<3, false>(W, n, &plan); break;
        default: rc = pad ? plan_v<4, true>(W, n, &plan) : plan_v<4, false>(W, n, &plan); break;
    }
    return rc == VPPB200_OK;
}

// 0 = done; 1 = this shape does not fit the cluster sweep (caller uses sgm.cu); < 0 = error.
// h_bwd = 0 leaves out r0 of pass 1 (the caller fuses it into the WTA sweep).
int launch_aggregate_plane(const uint8_t *img, const uint8_t *cost, uint16_t *S, int W, int H, int D, int n, int h_bwd,
                           cudaStream_t st)
{
    if ((long)W * H * 128 >= (1L << 31) || H < 3) return 1;
    const int nw = (D + 63) / 64;
    const bool pad = (D != 64 * nw);
    switch (nw) {
        case 1: return pad ? aggregate_plane_t<1, true>(img, cost, S, W, H, D, n, h_bwd, st)
                           : aggregate_plane_t<1, false>(img, cost, S, W, H, D, n, h_bwd, st);
        case 2: return pad ? aggregate_plane_t<2, true>(img, cost, S, W, H, D, n, h_bwd, st)
                           : aggregate_plane_t<2, false>(img, cost, S, W, H, D, n, h_bwd, st);
        case 3: return pad ? aggregate_plane_t<3, true>(img, cost, S, W, H, D, n, h_bwd, st)
                           : aggregate_plane_t<3, false>(img, cost, S, W, H, D, n, h_bwd, st);
        default: return pad ? aggregate_plane_t<4, true>(img, cost, S, W, H, D, n, h_bwd, st)
                            : aggregate_plane_t<4, false>(img, cost, S, W, H, D, n, h_bwd, st);
    }
}

}  // namespace vppb200
