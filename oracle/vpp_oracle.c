/*
 * vpp_oracle.c -- CPU restatement (plain C) of the reference's Virtual Pattern Projection scans.
 *
 * TEST INFRASTRUCTURE ONLY (see rsgm_oracle.c header).  Parity status: PINNED against the compiled reference
 * Cython module (oracle/_ref/vpp_core_opt*.so, arithmetic mode 0) and against the reference's numba twin
 * (vpp_standalone.py, arithmetic mode 1) in tests/test_oracle_vs_ref.py, plus golden vectors in tests/golden/.
 *
 * Citations relative to /root/reference.
 *   mode 0 = Cython vpp_core/vpp_core_opt.pyx  (c, c_occ, d1_blending are float32; products with them follow the C
 *            usual arithmetic conversions of the generated code; (int)round() is half-away-from-zero)
 *   mode 1 = numba vpp_standalone.py:14-369    (everything float64; round() is half-to-even; histogram bins uint8,
 *            book-keeping always outside the range test)
 * The random pattern is NOT drawn here: the caller passes the pre-drawn stream (`rand()%256` after `srand(seed)` for
 * mode 0, numba's generator for mode 1) and the scan consumes it in the reference's call order
 * (vpp_core_opt.pyx:92-93,:101-102).  Compile with -ffp-contract=off: the reference build has no FMA.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    int mode;               /* 0 cython, 1 numba */
    float c32, cocc32;      /* mode 0 */
    double c64, cocc64;     /* mode 1 (python floats) */
} blend_t;

static inline uint8_t tr(double v) { return (uint8_t)v; }   /* (uint8_t) cast: truncation, values in [0,256) */

/* colour term rc = pattern * c.  rnd mode 0: float32 product widened (vpp_core_opt.pyx:107); everything else double */
static inline double colour_term(const blend_t *B, int occ, int is_rnd, double pv)
{
    if (B->mode == 0) {
        float c = occ ? B->cocc32 : B->c32;
        if (is_rnd) return (double)((float)pv * c);           /* uint8 * float -> float */
        return pv * (double)c;                                /* ((pa+pb)/2) * c : double * float -> double (:318) */
    }
    return pv * (occ ? B->cocc64 : B->c64);
}
static inline double one_minus_c(const blend_t *B, int occ)
{
    if (B->mode == 0) return 1.0 - (double)(occ ? B->cocc32 : B->c32);   /* (1-c): int - float -> double in C */
    return 1.0 - (occ ? B->cocc64 : B->c64);
}

/* the splat of one patch pixel, shared by both scans (vpp_core_opt.pyx:104-124 / :315-335) */
static inline void splat_pixel(const blend_t *B, int is_rnd, double pv, uint8_t *lrow, uint8_t *rrow, int W, int C, int j,
                               int xl, int x0, int x1, int xr, int occluded, int discard, int interpolate,
                               float b32, double b64)
{
#define LP(x) lrow[(size_t)(x) * C + j]
#define RP(x) rrow[(size_t)(((x) < 0) ? (x) + W : (x)) * C + j]   /* Cython/numba negative-index wraparound */
    if (0 <= x0 && x0 <= W - 1) {
        if (!occluded) {
            double rc = colour_term(B, 0, is_rnd, pv), omc = one_minus_c(B, 0);
            LP(xl) = tr(rc + (double)LP(xl) * omc);
            if (interpolate) {
                if (B->mode == 0) {
                    /* :109  ((rv*c + r*(1-c)) * (1-b)) + r*b ; r*b is uint8*float -> float */
                    uint8_t r0 = RP(x0);
                    RP(x0) = tr((rc + (double)r0 * omc) * (1.0 - (double)b32) + (double)((float)r0 * b32));
                    if (0 <= x1 && x1 <= W - 1) {
                        uint8_t r1 = RP(x1);
                        /* :111  ((..)*b) + r*(1-b): r*(1-b) is uint8 * double */
                        RP(x1) = tr((rc + (double)r1 * omc) * (double)b32 + (double)r1 * (1.0 - (double)b32));
                    }
                } else {
                    uint8_t r0 = RP(x0);
                    RP(x0) = tr((rc + (double)r0 * omc) * (1.0 - b64) + (double)r0 * b64);
                    if (0 <= x1 && x1 <= W - 1) {
                        uint8_t r1 = RP(x1);
                        RP(x1) = tr((rc + (double)r1 * omc) * b64 + (double)r1 * (1.0 - b64));
                    }
                }
            } else {
                RP(xr) = tr(rc + (double)RP(xr) * omc);
            }
        } else if (!discard) {
            double rc = colour_term(B, 1, is_rnd, pv), omo = one_minus_c(B, 1), omc = one_minus_c(B, 0);
            if (interpolate) {
                if (B->mode == 0) {
                    uint8_t r0 = RP(x0);
                    RP(x0) = tr((rc + (double)r0 * omo) * (1.0 - (double)b32) + (double)((float)r0 * b32));
                    if (0 <= x1 && x1 <= W - 1) {
                        uint8_t r1 = RP(x1);
                        RP(x1) = tr((rc + (double)r1 * omo) * (double)b32 + (double)r1 * (1.0 - (double)b32));
                    }
                    /* :119  (r0'*(1-b) + r1'*b) * c + l*(1-c);  r1'*b is uint8*float -> float; *c is double*float */
                    double mix = (double)RP(x0) * (1.0 - (double)b32) + (double)((float)RP(x1) * b32);
                    LP(xl) = tr(mix * (double)B->c32 + (double)LP(xl) * omc);
                } else {
                    uint8_t r0 = RP(x0);
                    RP(x0) = tr((rc + (double)r0 * omo) * (1.0 - b64) + (double)r0 * b64);
                    if (0 <= x1 && x1 <= W - 1) {
                        uint8_t r1 = RP(x1);
                        RP(x1) = tr((rc + (double)r1 * omo) * b64 + (double)r1 * (1.0 - b64));
                    }
                    double mix = (double)RP(x0) * (1.0 - b64) + (double)RP(x1) * b64;
                    LP(xl) = tr(mix * B->c64 + (double)LP(xl) * omc);
                }
            } else {
                RP(xr) = tr(rc + (double)RP(xr) * omo);
                /* :122  r*c + l*(1-c): uint8*float -> float in mode 0 */
                if (B->mode == 0) LP(xl) = tr((double)((float)RP(xr) * B->c32) + (double)LP(xl) * omc);
                else LP(xl) = tr((double)RP(xr) * B->c64 + (double)LP(xl) * omc);
            }
        }
    } else {
        double rc = colour_term(B, 0, is_rnd, pv), omc = one_minus_c(B, 0);
        LP(xl) = tr(rc + (double)LP(xl) * omc);               /* left-side occlusion :123-124 */
    }
#undef LP
#undef RP
}

static inline int round_mode(const blend_t *B, float g)
{
    if (B->mode == 0) return (int)round((double)g);            /* C round: half away from zero (:82) */
    return (int)nearbyint((double)g);                          /* numba round: half to even */
}

/* ---- TPAMI extensions of vpp_standalone.py (numba only) ------------------------------------- */
typedef struct {
    int use_distance, use_bilateral;
    float dmin, dmax;            /* numpy float32 scalars (vpp_standalone.py:410-411) */
    double gamma;                /* python float */
    const float *filled;         /* [H][W] bilateral-filled hints (only read when use_bilateral) */
} adapt_t;

/* _get_patch_size_based_on_distance  vpp_standalone.py:6-11: float32 ratio, float64 pow (libm), round half to even */
static inline int adapt_radius(const adapt_t *A, float gv, int wsize)
{
    if (!A || !A->use_distance) return (wsize - 1) / 2;
    const float ratio = (gv - A->dmin) / (A->dmax - A->dmin);
    const double w = pow((double)ratio, 1.0 / A->gamma);
    const long ws = (long)nearbyint(w * (double)(wsize - 1) + 1.0);
    /* Python floor division of (ws - 1) by 2 */
    long q = (ws - 1) / 2;
    if ((ws - 1) % 2 != 0 && (ws - 1) < 0) q -= 1;
    return (int)q;
}
/* vpp_standalone.py:153-154 / :334-335: abs(g[y,x] - filled_g[yy,xx]) < 0.1.  _bilateral_filling returns FLOAT64 (numba
 * types np.where(cmap>th, aug_dmap, 0) as float64), so the difference is float32 - float64 -> float64 (exact). */
static inline int adapt_keep(const adapt_t *A, float gv, int W, int yy, int xx)
{
    if (!A || !A->use_bilateral) return 1;
    return fabs((double)gv - (double)A->filled[(size_t)yy * W + xx]) < 0.1;
}

/* _bilateral_filling  vpp_standalone.py:371-394: img is the uint8 gray context, cmap float32, weights float64 (libm exp);
 * o_xy / o_i are passed as doubles (numba types the defaults as int64; 2*(o**2) is exact either way for integers). */
ORC_API void orc_bilateral_filling(const float *dmap, const uint8_t *img, int W, int H, int n, double o_xy, double o_i, double th,
                                   float *out)
{
    float *cmap = (float *)calloc((size_t)W * H, sizeof(float));
    memcpy(out, dmap, (size_t)W * H * sizeof(float));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float d_ref = dmap[(size_t)y * W + x];
            if (!(d_ref > 0)) continue;
            const int i_ref = img[(size_t)y * W + x];
            for (int yw = -n; yw <= n; yw++)
                for (int xw = -n; xw <= n; xw++) {
                    if (y + yw < 0 || y + yw > H - 1 || x + xw < 0 || x + xw > W - 1) continue;
                    const long di = (long)img[(size_t)(y + yw) * W + x + xw] - i_ref;
                    const double weight = exp(-((double)(yw * yw + xw * xw) / (2.0 * (o_xy * o_xy)) + (double)(di * di) / (2.0 * (o_i * o_i))));
                    float *cm = cmap + (size_t)(y + yw) * W + x + xw;
                    if ((double)*cm < weight) { *cm = (float)weight; out[(size_t)(y + yw) * W + x + xw] = d_ref; }
                }
        }
    for (size_t i = 0; i < (size_t)W * H; i++)
        if (!((double)cmap[i] > th)) out[i] = 0.0f;
    free(cmap);
}

/* virtual_projection_scan_rnd  vpp_core_opt.pyx:53-131 / vpp_standalone.py:243-369.
 * `stream` holds the pre-drawn pattern values; returns the number of hints; *consumed = values used. */
static int scan_rnd(uint8_t *l, uint8_t *r, const float *g, int W, int H, int C, int uniform_color, int wsize,
                    int direction, double c, double c_occ, const uint8_t *g_occ, int discard, int interpolate,
                    int mode, const uint8_t *stream, long stream_len, long *consumed, const adapt_t *A)
{
    blend_t B = { mode, (float)c, (float)c_occ, c, c_occ };
    long si = 0;
    int count = 0;
    for (int y = 0; y < H; y++) {
        int x = direction == 0 ? W - 1 : 0;
        while ((direction != 0 && x < W) || (direction == 0 && x >= 0)) {
            const float gv = g[(size_t)y * W + x];
            if (gv > 0) {
                const int d = round_mode(&B, gv), d0 = (int)floor((double)gv), d1 = (int)ceil((double)gv);
                const float b32 = gv - (float)d0;                      /* float32 (:85) */
                const double b64 = (double)gv - (double)d0;            /* numba: float32 - int64 -> float64 */
                const int xd = x - d, xd0 = x - d0, xd1 = x - d1;
                const int occ = g_occ[(size_t)y * W + x] != 0;
                const int n = adapt_radius(A, gv, wsize);
                for (int j = 0; j < C; j++) {
                    uint8_t rv = 0;
                    if (uniform_color) rv = si < stream_len ? stream[si] : 0, si++;
                    for (int yw = -n; yw <= n; yw++)
                        for (int xw = -n; xw <= n; xw++) {
                            if (y + yw < 0 || y + yw > H - 1 || x + xw < 0 || x + xw > W - 1) continue;
                            if (!adapt_keep(A, gv, W, y + yw, x + xw)) continue;
                            if (!uniform_color) rv = si < stream_len ? stream[si] : 0, si++;
                            splat_pixel(&B, 1, (double)rv, l + (size_t)(y + yw) * W * C, r + (size_t)(y + yw) * W * C, W, C, j,
                                        x + xw, xd0 + xw, xd1 + xw, xd + xw, occ, discard, interpolate, b32, b64);
                        }
                }
                count++;
            }
            x = direction == 0 ? x - 1 : x + 1;
        }
    }
    if (consumed) *consumed = si;
    return count;
}

ORC_API int orc_vpp_scan_rnd(uint8_t *l, uint8_t *r, const float *g, int W, int H, int C, int uniform_color, int wsize,
                             int direction, double c, double c_occ, const uint8_t *g_occ, int discard, int interpolate,
                             int mode, const uint8_t *stream, long stream_len, long *consumed)
{
    return scan_rnd(l, r, g, W, H, C, uniform_color, wsize, direction, c, c_occ, g_occ, discard, interpolate, mode, stream,
                    stream_len, consumed, NULL);
}

/* the numba scan with the adaptive-patch flags (vpp_standalone.py:243-369 with :318-322 and :334-335) */
ORC_API int orc_vpp_scan_rnd_adaptive(uint8_t *l, uint8_t *r, const float *g, const float *filled_g, int W, int H, int C,
                                      int uniform_color, int wsize, int direction, double c, double c_occ, const uint8_t *g_occ,
                                      int discard, int interpolate, const uint8_t *stream, long stream_len, long *consumed,
                                      int use_distance, int use_bilateral, float dmin, float dmax, double gamma)
{
    adapt_t A = { use_distance, use_bilateral, dmin, dmax, gamma, filled_g };
    return scan_rnd(l, r, g, W, H, C, uniform_color, wsize, direction, c, c_occ, g_occ, discard, interpolate, 1, stream,
                    stream_len, consumed, &A);
}

/* histogram "max distance" colour of one window (vpp_core_opt.pyx:216-260 uniform, :269-313 per pixel;
 * vpp_standalone.py:100-144, :158-202).  (cy,cx) is the window centre in left coordinates, shift = x - xd. */
static double max_dist_colour(const uint8_t *l, const uint8_t *r, int W, int H, int C, int j, int cy, int cx, int shift,
                              int nax, int nay, int occ, int uniform_branch, int mode)
{
    int pa = 0, pb = 255, n_bins = 256;
    int bins32[256];
    uint8_t bins8[256];
    memset(bins32, 0, sizeof bins32);
    memset(bins8, 0, sizeof bins8);
    for (int ya = -nay; ya <= nay; ya++)
        for (int xa = -nax; xa <= nax; xa++) {
            const int yy = cy + ya, xx = cx + xa;
            if (yy < 0 || yy > H - 1 || xx < 0 || xx > W - 1) continue;
            const int xr = xx - shift, r_in = (0 <= xr && xr <= W - 1);
            for (int side = 0; side < 2; side++) {
                int v;
                if (side == 0) {
                    if (!(occ == 0 || !r_in)) continue;
                    v = l[((size_t)yy * W + xx) * C + j];
                } else {
                    if (!r_in) continue;
                    v = r[((size_t)yy * W + xr) * C + j];
                }
                const int inside = v > pa && v < pb;
                if (inside) {
                    if (v - pa > pb - v) pb = v;
                    else if (v - pa < pb - v) pa = v;
                }
                /* book-keeping: Cython uniform branch only inside the range test (:235-237,:248-250);
                 * Cython per-pixel branch (:288-290,:301-303) and numba (both branches) always */
                if (inside || !(mode == 0 && uniform_branch)) {
                    if (v == 0) n_bins -= 1;
                    bins32[v] += 1;
                    bins8[v] = (uint8_t)(bins8[v] + 1);
                }
            }
        }
    if (n_bins == 0) {
        int mb = 0;
        if (mode == 0) { int mv = bins32[0]; for (int k = 0; k < 256; k++) if (mv > bins32[k]) { mb = k; mv = bins32[k]; } }
        else { int mv = bins8[0]; for (int k = 0; k < 256; k++) if (mv > bins8[k]) { mb = k; mv = bins8[k]; } }
        pa = pb = mb;
    }
    return (double)(pa + pb) / 2.0;
}

/* virtual_projection_scan_max_dist  vpp_core_opt.pyx:133-341 / vpp_standalone.py:14-232 */
static int scan_max_dist(uint8_t *l, uint8_t *r, const float *g, int W, int H, int C, int uniform_color, int wsize,
                         int wsize_agg_x, int wsize_agg_y, int direction, double c, double c_occ,
                         const uint8_t *g_occ, int discard, int interpolate, int mode, const adapt_t *A)
{
    blend_t B = { mode, (float)c, (float)c_occ, c, c_occ };
    const int nax = (wsize_agg_x - 1) / 2, nay = (wsize_agg_y - 1) / 2;
    int count = 0;
    for (int y = 0; y < H; y++) {
        int x = direction == 0 ? W - 1 : 0;
        while ((direction != 0 && x < W) || (direction == 0 && x >= 0)) {
            const float gv = g[(size_t)y * W + x];
            if (gv > 0) {
                const int d = round_mode(&B, gv), d0 = (int)floor((double)gv), d1 = (int)ceil((double)gv);
                const float b32 = gv - (float)d0;
                const double b64 = (double)gv - (double)d0;
                const int xd = x - d, xd0 = x - d0, xd1 = x - d1;
                const int occ = g_occ[(size_t)y * W + x] != 0;
                const int n = adapt_radius(A, gv, wsize);
                for (int j = 0; j < C; j++) {
                    double pv = 0;
                    if (uniform_color) pv = max_dist_colour(l, r, W, H, C, j, y, x, x - xd, nax, nay, occ, 1, mode);
                    for (int yw = -n; yw <= n; yw++)
                        for (int xw = -n; xw <= n; xw++) {
                            if (y + yw < 0 || y + yw > H - 1 || x + xw < 0 || x + xw > W - 1) continue;
                            if (!adapt_keep(A, gv, W, y + yw, x + xw)) continue;
                            if (!uniform_color)
                                pv = max_dist_colour(l, r, W, H, C, j, y + yw, x + xw, x - xd, nax, nay, occ, 0, mode);
                            splat_pixel(&B, 0, pv, l + (size_t)(y + yw) * W * C, r + (size_t)(y + yw) * W * C, W, C, j,
                                        x + xw, xd0 + xw, xd1 + xw, xd + xw, occ, discard, interpolate, b32, b64);
                        }
                }
                count++;
            }
            x = direction == 0 ? x - 1 : x + 1;
        }
    }
    return count;
}

ORC_API int orc_vpp_scan_max_dist(uint8_t *l, uint8_t *r, const float *g, int W, int H, int C, int uniform_color, int wsize,
                                  int wsize_agg_x, int wsize_agg_y, int direction, double c, double c_occ,
                                  const uint8_t *g_occ, int discard, int interpolate, int mode)
{
    return scan_max_dist(l, r, g, W, H, C, uniform_color, wsize, wsize_agg_x, wsize_agg_y, direction, c, c_occ, g_occ, discard,
                         interpolate, mode, NULL);
}

/* the numba scan with the adaptive-patch flags (vpp_standalone.py:14-232 with :93-96 and :153-154) */
ORC_API int orc_vpp_scan_max_dist_adaptive(uint8_t *l, uint8_t *r, const float *g, const float *filled_g, int W, int H, int C,
                                           int uniform_color, int wsize, int wsize_agg_x, int wsize_agg_y, int direction, double c,
                                           double c_occ, const uint8_t *g_occ, int discard, int interpolate, int use_distance,
                                           int use_bilateral, float dmin, float dmax, double gamma)
{
    adapt_t A = { use_distance, use_bilateral, dmin, dmax, gamma, filled_g };
    return scan_max_dist(l, r, g, W, H, C, uniform_color, wsize, wsize_agg_x, wsize_agg_y, direction, c, c_occ, g_occ, discard,
                         interpolate, 1, &A);
}

/* gt_reshape  vpp_core_opt.pyx:352-371: raster-order compaction to (x, y, d, 1) */
ORC_API int orc_gt_reshape(const float *gt, int W, int H, float *out /* [W*H][4] */)
{
    int i = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (gt[(size_t)y * W + x] > 0) {
                out[4 * i] = (float)x; out[4 * i + 1] = (float)y; out[4 * i + 2] = gt[(size_t)y * W + x]; out[4 * i + 3] = 1.0f;
                i++;
            }
    return i;
}
