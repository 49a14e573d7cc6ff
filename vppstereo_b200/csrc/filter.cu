// filter.cu -- the occlusion heuristic that produces the VPP occlusion mask g_occ (filter.py:246-292, called from
// test.py:154), as exact parallel restatements of the reference's sequential numba scans (sm_100a).
//
//   left_warp      (filter.py:7-48)    "largest disparity wins" per right-view pixel  -> atomicMax on the float bits
//                                      (disparities are positive, so the IEEE order is the integer order)
//   weighted_conf  (filter.py:113-164) scatter "foreground rejects nearby background" -> gather per target pixel over the
//                                      mirrored 9x8 window; rejections never change what later pixels read, so order-free
//   filter         (filter.py:167-194) point-wise
//   left_unwarp / conf_unwarp (filter.py:50-111)  later x overwrites earlier x        -> atomicMax on the source column
//   interpolate_disparity(dmap, 3) (filter.py:196-243)  a filled gap is never read as a neighbour by a later pixel (its
//                                      neighbours are non-zero already), so every pixel can look at the un-filled map
// Arithmetic follows numba's typing for float32 maps: round half to even, float32 differences, float64 penalties/thresholds.
#include "common.cuh"

namespace vppb200 {

struct OccWs {
    uint32_t *omap;      // [n][H][W] float bits of the warped map (right view)
    uint32_t *src;       // [n][H][W] 1 + source column of the un-warp (0 = none)
    float *omapf;        // [n][H][W] warped map after the confidence filter
    uint8_t *conf;       // [n][H][W] confidence in the right view
};

static inline size_t occ_align(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t occ_ws_layout(int H, int W, int n, void *base, OccWs *ws)
{
    size_t off = 0;
    char *b = (char *)base;
    const size_t np = (size_t)n * H * W;
    auto take = [&](size_t bytes) { size_t o = off; off += occ_align(bytes); return b ? b + o : (char *)nullptr; };
    uint32_t *omap = (uint32_t *)take(np * 4);
    uint32_t *src = (uint32_t *)take(np * 4);          // directly behind omap: one memset clears both
    float *omapf = (float *)take(np * 4);
    uint8_t *conf = (uint8_t *)take(np);
    if (ws) { ws->omap = omap; ws->src = src; ws->omapf = omapf; ws->conf = conf; }
    return off;
}

__global__ void occ_warp_kernel(const float *__restrict__ dmap, uint32_t *__restrict__ omap, int W, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float d = dmap[i];
    if (!(d > 0.0f)) return;
    const int x = (int)(i % W);
    const int xd = x - __float2int_rn(d);                 // int(round(float32)): half to even
    if (0 <= xd && xd <= W - 1) atomicMax(omap + (i - x + xd), __float_as_uint(d));
}

__global__ void occ_conf_kernel(const uint32_t *__restrict__ omap, float *__restrict__ omapf, uint8_t *__restrict__ conf, int W,
                                int H, int rx, int ry, double l, double g, double th_conf, double th_filter, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const float dn = __uint_as_float(omap[i]);
    int c = 1;                                            // empty pixels are "rejected" (filter.py:161-162)
    if (dn > 0.0f) {
        c = 0;
        const uint32_t *frame = omap + (i - (long)y * W - x);
        // the source (ys, xs) reaches this pixel with yw = y - ys in [-ry-1, ry] and xw = x - xs in [-rx, rx]
        for (int yw = -ry - 1; yw <= ry && !c; yw++) {
            const int ys = y - yw;
            if (ys < 0 || ys > H - 1) continue;
            const double pen_y = __dmul_rn(__dsub_rn(1.0, g), (double)abs(yw));
            for (int xw = -rx; xw <= rx; xw++) {
                const int xs = x - xw;
                if (xs < 0 || xs > W - 1) continue;
                const float dc = __uint_as_float(frame[(long)ys * W + xs]);
                if (dc > 0.0f && dn < dc) {
                    const float diff = __fsub_rn(dc, dn);
                    const double pen = __dmul_rn(l, __dadd_rn(__dmul_rn(g, (double)abs(xw)), pen_y));
                    if (__dsub_rn((double)diff, pen) > th_conf) { c = 1; break; }
                }
            }
        }
    }
    conf[i] = (uint8_t)c;
    omapf[i] = (dn > 0.0f && !((double)c > th_filter)) ? dn : 0.0f;
}

__global__ void occ_unwarp_kernel(const float *__restrict__ omapf, uint32_t *__restrict__ src, int W, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float d = omapf[i];
    if (!(d > 0.0f)) return;
    const int x = (int)(i % W);
    const int xd = x + __float2int_rn(d);
    if (0 <= xd && xd <= W - 1) atomicMax(src + (i - x + xd), (uint32_t)x + 1u);       // the last writer in scan order wins
}

__global__ void occ_finish_kernel(const float *__restrict__ omapf, const uint8_t *__restrict__ conf, const uint32_t *__restrict__ src,
                                  float *__restrict__ dmap_out, uint8_t *__restrict__ conf_out, int W, int H, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long row = i - x;
    auto unwarped = [&](long r, int xx) -> float {
        const uint32_t s = src[r + xx];
        return s ? omapf[r + s - 1] : 0.0f;
    };
    const uint32_t s = src[i];
    float v = s ? omapf[row + s - 1] : 0.0f;
    if (conf_out) conf_out[i] = s ? conf[row + s - 1] : (uint8_t)1;
    if (dmap_out) {
        if (v == 0.0f) {
            // neighbours in the reference's addressing: column -1 wraps inside the row, column W is the flat successor
            const float nl = x > 0 ? unwarped(row, x - 1) : unwarped(row, W - 1);
            const float nr = x + 1 < W ? unwarped(row, x + 1) : (y + 1 < H ? unwarped(row + W, 0) : 0.0f);
            if (nl > 0.0f && nr > 0.0f) {
                const double a = (double)nl, b = (double)nr;
                if (fabs(__dsub_rn(a, b)) < 1.0) {
                    const double m = __ddiv_rn(__dsub_rn(b, a), 2.0);
                    v = (float)__dadd_rn(__dmul_rn(m, 0.0), __dsub_rn(a, __dmul_rn(m, -1.0)));
                }
            }
        }
        dmap_out[i] = v;
    }
}

}  // namespace vppb200

using namespace vppb200;

extern "C" size_t vppb200_occlusion_workspace_bytes(int H, int W, int n)
{
    if (H <= 0 || W <= 0 || n <= 0) return 0;
    return occ_ws_layout(H, W, n, nullptr, nullptr);
}

extern "C" int vppb200_occlusion_heuristic(const float *dmap, float *dmap_out, uint8_t *conf_out, int W, int H, int rx, int ry,
                                           double l, double g, double th_conf, double th_filter, void *workspace,
                                           size_t workspace_bytes, int n, void *stream)
{
    if (!dmap || (!dmap_out && !conf_out) || W <= 0 || H <= 0 || n <= 0 || rx < 0 || ry < 0) return VPPB200_ERR_ARG;
    if (!workspace || workspace_bytes < occ_ws_layout(H, W, n, nullptr, nullptr)) return VPPB200_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    OccWs ws;
    occ_ws_layout(H, W, n, workspace, &ws);
    const long total = (long)n * H * W;
    VPP_CUDA_TRY(cudaMemsetAsync(ws.omap, 0, (size_t)((char *)ws.omapf - (char *)ws.omap), st));
    const int blocks = cdiv(total, 256);
    occ_warp_kernel<<<blocks, 256, 0, st>>>(dmap, ws.omap, W, total);
    VPP_LAUNCH_CHECK("occ_warp_kernel");
    occ_conf_kernel<<<blocks, 256, 0, st>>>(ws.omap, ws.omapf, ws.conf, W, H, rx / 2, ry / 2, l, g, th_conf, th_filter, total);
    VPP_LAUNCH_CHECK("occ_conf_kernel");
    occ_unwarp_kernel<<<blocks, 256, 0, st>>>(ws.omapf, ws.src, W, total);
    VPP_LAUNCH_CHECK("occ_unwarp_kernel");
    occ_finish_kernel<<<blocks, 256, 0, st>>>(ws.omapf, ws.conf, ws.src, dmap_out, conf_out, W, H, total);
    VPP_LAUNCH_CHECK("occ_finish_kernel");
    return VPPB200_OK;
}
