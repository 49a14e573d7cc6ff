/*
 * filter_oracle.c -- CPU restatement (plain C) of the reference's occlusion heuristic, the producer of the VPP occlusion
 * mask g_occ (filter.py:246-292, called from test.py:154 as occlusion_heuristic(hints)[1]).
 *
 * TEST INFRASTRUCTURE ONLY (see rsgm_oracle.c header).  Parity status: PINNED against the reference's own numba code
 * (oracle/_ref/filter_ref.pycode = byte-compiled filter.py) in tests/test_oracle_vs_ref.py and by golden vectors in
 * tests/golden/.
 *
 * numba typing that matters (float32 disparity maps, as test.py passes them):
 *   round(float32) is half-to-even; dmap differences are float32; the window penalty l*(g*|xw|+(1-g)*|yw|) and every
 *   comparison with a threshold are float64; n_left / n_right of interpolate_disparity unify to float64.
 * Out-of-bounds reads of the reference (interpolate_disparity has no bounds checks, filter.py:218-226): column -1 wraps to
 * column w-1 of the same row (numba wraparound); column w is the flat successor, i.e. column 0 of the next row; for the last
 * row that is one element past the buffer (undefined in the reference) and is DEFINED here as 0 (no neighbour).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

static inline int round_half_even_f32(float v) { return (int)nearbyintf(v); }   /* default rounding mode = to nearest even */

/* filter.py:7-48  left_warp: omap[y, x-d] = max of the colliding disparities */
static void left_warp(const float *dmap, float *omap, int w, int h)
{
    memset(omap, 0, sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float d = dmap[(size_t)y * w + x];
            if (d > 0) {
                const int xd = x - round_half_even_f32(d);
                if (0 <= xd && xd <= w - 1 && omap[(size_t)y * w + xd] < d) omap[(size_t)y * w + xd] = d;
            }
        }
}

/* filter.py:113-164  weighted_conf (rx, ry already halved by the caller as in :143-144) */
static void weighted_conf(const float *dmap, uint8_t *conf, int w, int h, int rx, int ry, double l, double g, double th)
{
    memset(conf, 0, (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float dc = dmap[(size_t)y * w + x];
            if (dc > 0) {
                for (int xw = -rx; xw <= rx; xw++)
                    for (int yw = -ry - 1; yw <= ry; yw++) {
                        if (0 <= y + yw && y + yw <= h - 1 && 0 <= x + xw && x + xw <= w - 1) {
                            const float dn = dmap[(size_t)(y + yw) * w + x + xw];
                            if (dn > 0 && dn < dc) {
                                const float diff = dc - dn;                         /* float32 - float32 */
                                if ((double)diff - l * (g * (double)abs(xw) + (1.0 - g) * (double)abs(yw)) > th)
                                    conf[(size_t)(y + yw) * w + x + xw] = 1;
                            }
                        }
                    }
            } else {
                conf[(size_t)y * w + x] = 1;
            }
        }
}

/* filter.py:167-194 filter, :50-79 left_unwarp, :81-111 conf_unwarp, :196-243 interpolate_disparity(n=3, th=1) */
ORC_API void orc_occlusion_heuristic(const float *dmap_in, float *dmap_out, uint8_t *conf_out, int w, int h, int rx, int ry,
                                     double l, double g, double th_conf, double th_filter)
{
    const size_t np = (size_t)w * h;
    float *omap = (float *)malloc(np * sizeof(float));
    uint8_t *conf = (uint8_t *)malloc(np);
    left_warp(dmap_in, omap, w, h);
    weighted_conf(omap, conf, w, h, rx / 2, ry / 2, l, g, th_conf);
    for (size_t i = 0; i < np; i++)
        if (omap[i] > 0 && (double)conf[i] > th_filter) omap[i] = 0;
    memset(dmap_out, 0, np * sizeof(float));
    memset(conf_out, 1, np);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float d = omap[(size_t)y * w + x];
            if (d > 0) {
                const int xd = x + round_half_even_f32(d);
                if (0 <= xd && xd <= w - 1) {
                    dmap_out[(size_t)y * w + xd] = d;
                    conf_out[(size_t)y * w + xd] = conf[(size_t)y * w + x];
                }
            }
        }
    /* interpolate_disparity(dmap, 3): n = 1, th = 1; sequential and in place like the reference */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float *row = dmap_out + (size_t)y * w;
            if (row[x] == 0) {
                float *pl = x - 1 >= 0 ? &row[x - 1] : &row[w - 1];                 /* negative index wraps inside the row */
                float *pr = (x + 1 < w || y + 1 < h) ? &row[x + 1] : NULL;          /* flat successor; past the buffer: none */
                const double n_left = *pl > 0 ? (double)*pl : 0.0;
                const double n_right = (pr && *pr > 0) ? (double)*pr : 0.0;
                if (n_left > 0 && n_right > 0 && fabs(n_left - n_right) < 1.0) {
                    const double m = (n_right - n_left) / 2.0, q = n_left - m * -1.0;
                    *pl = (float)(m * -1.0 + q);
                    row[x] = (float)(m * 0.0 + q);
                    *pr = (float)(m * 1.0 + q);
                }
            }
        }
    free(omap);
    free(conf);
}
