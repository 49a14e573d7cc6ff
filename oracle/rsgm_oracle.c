/*
 * rsgm_oracle.c -- CPU restatement (plain C) of the reference's rSGM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this file.  The product (vppstereo_b200/) never does; it fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the compiled, unmodified reference
 * (oracle/_ref, built by oracle/build_ref.py) in tests/test_oracle_vs_ref.py (runs where /root/reference exists)
 * and against committed golden vectors generated from that reference (tests/golden/, tests/test_golden.py).
 * The reference leaves some census pixels unwritten (uninitialised malloc, non-deterministic); this oracle
 * defines them as 0 and the golden generator zeroes them in the reference's census output.
 *
 * All file:line citations are relative to /root/reference; RSGM/ = thirdparty/stereo-vision/reconstruction/base/rSGM/.
 * Third-party arithmetic that is NOT under /root/reference (un-vendored OpenCV 4.13.0: cvtColor, copyMakeBorder,
 * filterSpeckles; numba 0.65 round()) is restated from its published behaviour and pinned by running the installed
 * library side by side in tests/test_oracle_thirdparty.py.
 *
 * Layouts: images row-major [H][W]; cost volumes [H][W][D] with d fastest (RSGM/StereoBMHelper.h:138-141).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <xmmintrin.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * OpenCV pieces (third-party, un-vendored; pinned against cv2 4.13.0 in tests)
 * ---------------------------------------------------------------------------------------------- */

/* cv2.cvtColor(img, COLOR_RGB2GRAY) for uint8 (models/rsgm/rsgm.py:11-12): 15-bit fixed point. */
ORC_API void orc_rgb2gray(const uint8_t *rgb, uint8_t *gray, int n_px)
{
    for (int i = 0; i < n_px; i++) {
        uint32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        gray[i] = (uint8_t)((r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15);
    }
}

/* cv2.cvtColor(img, COLOR_BGR2GRAY) (vpp_standalone.py:415): same weights, channel order swapped. */
ORC_API void orc_bgr2gray(const uint8_t *bgr, uint8_t *gray, int n_px)
{
    for (int i = 0; i < n_px; i++) {
        uint32_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15);
    }
}

static int reflect_idx(int p, int len)
{
    /* BORDER_REFLECT: fedcba|abcdefgh|hgfedcb (edge pixel repeated) */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p - 1;
        else p = 2 * len - 1 - p;
    }
    return p;
}

/* cv2.copyMakeBorder(src, top, bottom, left, right, BORDER_REFLECT) on `elem` bytes per pixel (rsgm.py:258-260) */
ORC_API void orc_pad_reflect(const uint8_t *src, uint8_t *dst, int H, int W, int elem,
                             int top, int bottom, int left, int right)
{
    int Hp = H + top + bottom, Wp = W + left + right;
    for (int y = 0; y < Hp; y++) {
        int sy = reflect_idx(y - top, H);
        for (int x = 0; x < Wp; x++) {
            int sx = reflect_idx(x - left, W);
            memcpy(dst + ((size_t)y * Wp + x) * elem, src + ((size_t)sy * W + sx) * elem, (size_t)elem);
        }
    }
}

/* cv2.filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) on the int16 view the reference uses after
 * astype(uint8) (rsgm.py:284-285; cv2 accepts CV_8U in 4.x).  4-connected components under |a-b|<=maxDiff,
 * pixels equal to newVal are skipped, components with <= maxSpeckleSize pixels are set to newVal. */
ORC_API void orc_filter_speckles_u8(uint8_t *img, int H, int W, int newVal, int maxSize, int maxDiff)
{
    int n = H * W;
    int *label = (int *)calloc((size_t)n, sizeof(int));
    int *stack = (int *)malloc((size_t)n * sizeof(int));
    uint8_t *small = (uint8_t *)calloc((size_t)n + 1, 1);
    int cur = 0;
    for (int p = 0; p < n; p++) {
        if (img[p] == newVal) continue;
        if (label[p]) {
            if (small[label[p]]) img[p] = (uint8_t)newVal;
            continue;
        }
        cur++;
        int sp = 0, count = 0;
        stack[sp++] = p;
        label[p] = cur;
        while (sp) {
            int q = stack[--sp];
            count++;
            int y = q / W, x = q % W, v = img[q];
            int nb[4] = { x + 1 < W ? q + 1 : -1, x > 0 ? q - 1 : -1, y + 1 < H ? q + W : -1, y > 0 ? q - W : -1 };
            for (int k = 0; k < 4; k++) {
                int t = nb[k];
                if (t < 0 || label[t] || img[t] == newVal) continue;
                if (abs((int)img[t] - v) <= maxDiff) {
                    label[t] = cur;
                    stack[sp++] = t;
                }
            }
        }
        if (count <= maxSize) {
            small[cur] = 1;
            img[p] = (uint8_t)newVal;
        }
    }
    free(label);
    free(stack);
    free(small);
}

/* ------------------------------------------------------------------------------------------------
 * census5x5_SSE  (RSGM/FastFilters.cpp:181-442)
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_census5x5(const uint8_t *src, uint32_t *dst, int W, int H)
{
    const long n = (long)W * H;
    memset(dst, 0, (size_t)n * sizeof(uint32_t)); /* unwritten pixels are DEFINED as 0 (reference: uninitialised) */
    /* SSE body: flat centres c in [2W+2, W(H-2)-17]  (:214-421); neighbours wrap across row ends (flat stream) */
    const long c_lo = 2L * W + 2, c_hi = (long)W * (H - 2) - 17;
    for (long c = c_lo; c <= c_hi; c++) {
        uint32_t v = 0;
        int k = 0;
        const uint8_t cv = src[c];
        for (int dy = -2; dy <= 2; dy++)
            for (int dx = -2; dx <= 2; dx++) {
                if (dy == 0 && dx == 0) continue;
                if (src[c + (long)dy * W + dx] < cv)          /* _mm_cmplt_epi8 after ^0x80 (:204-221) */
                    v |= 1u << (8 * (k / 8) + (7 - k % 8));    /* byte k/8, MSB first (:273-281,:320-328,:352-359) */
                k++;
            }
        dst[c] = v;
    }
    /* dst[2W], dst[2W+1] = lastResult = 0 (:212,:406): already zero */
    /* scalar tail (:424-441): row H-3, columns W-14..W-3, different bit order (first neighbour = MSB of 24) */
    {
        const int i = H - 3;
        for (int j = W - 16 + 2; j < W - 2; j++) {
            const int cv = src[(long)i * W + j];
            uint32_t v = 0;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++)
                    if (dy != 0 || dx != 0) {
                        v *= 2;
                        if (cv > src[(long)(i + dy) * W + (j + dx)]) v += 1;
                    }
            dst[(long)i * W + j] = v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * costMeasureCensus5x5_xyd_SSE  (RSGM/StereoBMHelper.cpp:29-140), invalidDispValue = 12 (RSGM/pyrSGM.cpp:190)
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_cost_census(const uint32_t *cl, const uint32_t *cr, uint16_t *dsi, int W, int H, int D)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            uint16_t *p = dsi + ((size_t)i * W + j) * D;
            if (i < 2 || i >= H - 2) {
                for (int d = 0; d < D; d++) p[d] = 12;          /* :82-88, :133-139 */
                continue;
            }
            const uint32_t l = cl[(size_t)i * W + j];
            for (int d = 0; d < D; d++)
                p[d] = d > j ? 12 : (uint16_t)__builtin_popcount(l ^ cr[(size_t)i * W + j - d]); /* :45-71 */
        }
}

/* ------------------------------------------------------------------------------------------------
 * aggregate_SSE -> StereoSGM<uint8>::accumulateVariableParamsSSE<8>  (RSGM/StereoSGM_SSE.hpp:13-515)
 * Effective parameters are the struct defaults (RSGM/StereoSGM.h:33-47) because RSGM/pyrSGM.cpp:519 copies the
 * params before :557-560 mutate the local: P1=7, Alpha=0.25, Gamma=50, P2min=17.  They are arguments here so that
 * tests can also exercise other values; the parity entry point passes the defaults.
 * ---------------------------------------------------------------------------------------------- */
static inline uint16_t sat_add16(uint32_t a, uint32_t b) { uint32_t s = a + b; return (uint16_t)(s > 65535u ? 65535u : s); }

static inline int adapt_p2(float alpha, int a, int b, int gamma, int p2min)
{
    int r = (int)(-alpha * (float)abs(a - b) + (float)gamma);      /* RSGM/StereoSGM.hpp:92-99 */
    return r < p2min ? p2min : r;
}

/* one path step, SSE semantics (:296-308): L = C (+sat) subs( min(Lp[d], Lp[d-1]+P1, Lp[d+1]+P1, P2+mp), mp ) */
static void path_step(const uint16_t *C, const uint16_t *Lp /* [-1..D] valid */, uint16_t mp, int P1, int P2, int D,
                      uint16_t *Lout, uint16_t *mout)
{
    uint16_t cur_p2 = sat_add16((uint32_t)P2, mp);
    uint16_t m = 65535;
    for (int d = 0; d < D; d++) {
        uint16_t a = sat_add16(Lp[d - 1], (uint32_t)P1), b = sat_add16(Lp[d + 1], (uint32_t)P1);
        uint16_t t = Lp[d];
        if (a < t) t = a;
        if (b < t) t = b;
        if (cur_p2 < t) t = cur_p2;
        t = (uint16_t)(t > mp ? t - mp : 0);
        uint16_t v = sat_add16(C[d], t);
        Lout[d] = v;
        if (v < m) m = v;
    }
    *mout = m;
}

ORC_API void orc_sgm_aggregate(const uint8_t *img, const uint16_t *dsi, uint16_t *S, int W, int H, int D,
                               int P1, int P2min, float alpha, int gamma)
{
    const int DP = D + 2;
    /* per-column path state of the previous row, slots [-1] and [D] hold 65535; column slots -1 and W are the
     * image border: L = 65535 for all d, min = 0 (:48-58, :69-72) */
    size_t colsz = (size_t)DP;
    uint16_t *L1a = (uint16_t *)malloc((W + 2) * colsz * 2), *L1b = (uint16_t *)malloc((W + 2) * colsz * 2);
    uint16_t *L2 = (uint16_t *)malloc((W + 2) * colsz * 2), *L3 = (uint16_t *)malloc((W + 2) * colsz * 2);
    uint16_t *L0a = (uint16_t *)malloc(colsz * 2), *L0b = (uint16_t *)malloc(colsz * 2), *tmp = (uint16_t *)malloc(colsz * 2);
    uint16_t *m1a = (uint16_t *)calloc(W + 2, 2), *m1b = (uint16_t *)calloc(W + 2, 2), *m2 = (uint16_t *)calloc(W + 2, 2),
             *m3 = (uint16_t *)calloc(W + 2, 2);
    memset(L1a, 0xFF, (W + 2) * colsz * 2); memset(L1b, 0xFF, (W + 2) * colsz * 2);
    memset(L2, 0xFF, (W + 2) * colsz * 2);  memset(L3, 0xFF, (W + 2) * colsz * 2);
    memset(L0a, 0xFF, colsz * 2); memset(L0b, 0xFF, colsz * 2); memset(tmp, 0xFF, colsz * 2);
#define COL(buf, j) ((buf) + ((size_t)((j) + 1)) * colsz + 1)   /* pointer to d=0 of column j, j in [-1, W] */
    uint16_t *L1 = L1a, *L1last = L1b, *m1 = m1a, *m1last = m1b;
    uint16_t *L0 = L0a + 1, *L0last = L0b + 1;
    uint16_t m0last = 0;

    for (int pass = 0; pass < 2; pass++) {
        const int i1 = pass == 0 ? 0 : H - 1, i2 = pass == 0 ? H : -1, di = pass == 0 ? 1 : -1;
        const int j1 = pass == 0 ? 0 : W - 1, j2 = pass == 0 ? W : -1, dj = di;
        const uint8_t *img_line = img + (size_t)i1 * W;
        /* first pixel of the first line (:116-149) */
        {
            const uint16_t *C = dsi + ((size_t)i1 * W + j1) * D;
            uint16_t *Sp = S + ((size_t)i1 * W + j1) * D;
            uint16_t mc = 65535;
            for (int d = 0; d < D; d++) {
                uint16_t c = C[d];
                if (c == 255) c = 12;                                             /* :120 */
                L0last[d] = c; COL(L1last, j1)[d] = c; COL(L2, j1)[d] = c; COL(L3, j1)[d] = c;
                if (c < mc) mc = c;
                if (pass == 0) Sp[d] = c; else Sp[d] = (uint16_t)(Sp[d] + c);     /* :129 / :143 (wrapping +=) */
            }
            m0last = mc; m1last[j1 + 1] = mc; m2[j1 + 1] = mc; m3[j1 + 1] = mc;
        }
        /* rest of the first line (:152-218): only r0 aggregates */
        for (int j = j1 + dj; j != j2; j += dj) {
            const uint16_t *C = dsi + ((size_t)i1 * W + j) * D;
            uint16_t *Sp = S + ((size_t)i1 * W + j) * D;
            uint16_t mc = 65535, m0 = 65535;
            const int P2 = adapt_p2(alpha, img_line[j], img_line[j - dj], gamma, P2min);
            for (int d = 0; d < D; d++) {
                uint16_t c = C[d];
                if (c == 255) c = 12;
                COL(L1last, j)[d] = c; COL(L2, j)[d] = c; COL(L3, j)[d] = c;
                if (c < mc) mc = c;
                int32_t mp = L0last[d];
                int32_t a = (int32_t)L0last[d - 1] + P1; if (mp > a) mp = a;
                int32_t b = (int32_t)L0last[d + 1] + P1; if (mp > b) mp = b;
                int32_t p2 = (int32_t)m0last + P2;      if (mp > p2) mp = p2;
                mp -= m0last;
                int32_t nc = (int32_t)c + mp;
                uint16_t v = (uint16_t)(nc < 0 ? 0 : nc > 65535 ? 65535 : nc);    /* saturate_cast<uint16> */
                L0[d] = v;
                if (v < m0) m0 = v;
                if (pass == 0) Sp[d] = v; else Sp[d] = (uint16_t)(Sp[d] + v);
            }
            m1last[j + 1] = mc; m2[j + 1] = mc; m3[j + 1] = mc;
            { uint16_t *t = L0; L0 = L0last; L0last = t; }
            m0last = m0;
        }
        /* remaining lines (:224-499) */
        int il = i1 + di;                                   /* img_line_last == img_line on the 2nd line (:221) */
        for (int i = i1 + di; i != i2; i += di) {
            memset(L0last, 0, (size_t)D * 2);               /* :226-227 */
            m0last = 0;
            img_line = img + (size_t)i * W;
            const uint8_t *img_last = img + (size_t)il * W;
            for (int j = j1; j != j2; j += dj) {
                const uint16_t *C = dsi + ((size_t)i * W + j) * D;
                uint16_t *Sp = S + ((size_t)i * W + j) * D;
                const int p2_0 = adapt_p2(alpha, img_line[j], img_line[j - dj], gamma, P2min);   /* flat-indexed */
                const int p2_1 = adapt_p2(alpha, img_line[j], img_last[j - dj], gamma, P2min);
                const int p2_2 = adapt_p2(alpha, img_line[j], img_last[j], gamma, P2min);
                const int p2_3 = adapt_p2(alpha, img_line[j], img_last[j + dj], gamma, P2min);
                uint16_t m0, mm1, mm2, mm3;
                /* r0: predecessor (i, j-dj), in place */
                memcpy(tmp + 1, L0last, (size_t)D * 2);
                path_step(C, tmp + 1, m0last, P1, p2_0, D, L0last, &m0);
                for (int d = 0; d < D; d++) L0[d] = L0last[d];
                /* r1: predecessor (i-di, j-dj) from the last-row buffer into the current-row buffer */
                path_step(C, COL(L1last, j - dj), m1last[j - dj + 1], P1, p2_1, D, COL(L1, j), &mm1);
                /* r2: predecessor (i-di, j), in place */
                memcpy(tmp + 1, COL(L2, j), (size_t)D * 2);
                path_step(C, tmp + 1, m2[j + 1], P1, p2_2, D, COL(L2, j), &mm2);
                /* r3: predecessor (i-di, j+dj) (not yet overwritten on this line), written at j */
                path_step(C, COL(L3, j + dj), m3[j + dj + 1], P1, p2_3, D, COL(L3, j), &mm3);
                for (int d = 0; d < D; d++) {
                    uint16_t s = sat_add16(sat_add16(sat_add16(L0last[d], COL(L1, j)[d]), COL(L2, j)[d]), COL(L3, j)[d]);
                    Sp[d] = pass == 0 ? s : sat_add16(Sp[d], s);                  /* :397-402 */
                }
                m0last = m0; m1[j + 1] = mm1; m2[j + 1] = mm2; m3[j + 1] = mm3;
            }
            il = i;                                                              /* :492 */
            { uint16_t *t = L1; L1 = L1last; L1last = t; t = m1; m1 = m1last; m1last = t; }
            /* the swapped-in buffers keep their border columns: L=65535, min=0 */
        }
    }
#undef COL
    free(L1a); free(L1b); free(L2); free(L3); free(L0a); free(L0b); free(tmp);
    free(m1a); free(m1b); free(m2); free(m3);
}

/* ------------------------------------------------------------------------------------------------
 * matchWTA_SSE / matchWTARight_SSE  (RSGM/StereoBMHelper.cpp:634-750, :893-1015): first arg-min; the uniqueness
 * test is dead code (:717,:745,:981,:1009).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_wta_left(const uint16_t *S, float *disp, int W, int H, int D)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            const uint16_t *p = S + ((size_t)i * W + j) * D;
            int end = j < D - 1 ? j : D - 1, best = 0;
            uint32_t mc = p[0];
            for (int d = 1; d <= end; d++)
                if (p[d] < mc) { mc = p[d]; best = d; }
            disp[(size_t)i * W + j] = (float)best;
        }
}

ORC_API void orc_wta_right(const uint16_t *S, float *disp, int W, int H, int D)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            int end = (W - 1 - j) < D - 1 ? (W - 1 - j) : D - 1, best = 0;
            const uint16_t *p = S + ((size_t)i * W + j) * D;
            uint32_t mc = p[0];
            for (int k = 1; k <= end; k++) {
                uint16_t c = p[(size_t)k * D + k];       /* S[i][j+k][k]  (:918-923, :988) */
                if (c < mc) { mc = c; best = k; }
            }
            disp[(size_t)i * W + j] = (float)best;
        }
}

/* rcp_nz_ss (RSGM/StereoBMHelper.cpp:752-756): hardware RCPSS, 0 -> 0.  CPU-vendor specific table instruction:
 * the product uploads a LUT of exactly these values computed on the host it runs on. */
ORC_API float orc_rcp_nz(float v)
{
    if (v == 0.0f) return 0.0f;
    return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(v)));
}

/* LUT[k] = rcp_nz(-2k), k = 0..65535  (lowerMin = min(c1-c0, c1-c2) in {0,-1,...,-65535}) */
ORC_API void orc_rcp_lut(float *lut)
{
    for (int k = 0; k < 65536; k++) lut[k] = orc_rcp_nz(2.0f * (float)(-k));
}

/* subPixelRefine (RSGM/StereoBMHelper.cpp:1065-1135).  lut may be NULL (use this CPU's RCPSS) or a 65536-entry table
 * as produced by orc_rcp_lut (used with golden vectors recorded on another CPU). */
ORC_API void orc_subpixel(const uint16_t *S, float *disp, int W, int H, int D, int method, const float *lut)
{
    for (int y = 0; y < H; y++)
        for (int x = 1; x < W - 1; x++) {
            float *dp = disp + (size_t)y * W + x;
            if (*dp > 0.0f) {
                int dm = (int)*dp;
                const uint16_t *c = S + ((size_t)y * W + x) * D + dm;       /* dm = D-1 reads the next pixel's d=0 */
                int c0 = c[-1], c1 = c[0], c2 = c[1];
                if (method == 0) {
                    float left = (float)(c1 - c0), right = (float)(c1 - c2);
                    float lower = left < right ? left : right;           /* _mm_min_ss */
                    float r = lut ? lut[(int)(-lower)] : orc_rcp_nz(2.0f * lower);
                    *dp = (float)dm + (float)(c2 - c0) * r;
                } else {
                    int a = c0 + c0 - 4 * c1 + c2 + c2, b = c0 - c2;     /* :1119-1124 */
                    *dp = (float)dm + (float)b / (float)a;
                }
            } else {
                *dp = -10.0f;
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * median3x3_SSE  (RSGM/FastFilters.cpp:701-757): flat 3x3 median for c in [W+1, WH-W-5], copy elsewhere
 * ---------------------------------------------------------------------------------------------- */
static int cmp_f(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

ORC_API void orc_median3x3(const float *src, float *dst, int W, int H)
{
    const long n = (long)W * H;
    memcpy(dst, src, (size_t)n * sizeof(float));
    for (long c = W + 1; c <= n - W - 5; c++) {
        float v[9];
        int k = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) v[k++] = src[c + (long)dy * W + dx];
        qsort(v, 9, sizeof(float), cmp_f);
        dst[c] = v[4];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Python tail of compute_rsgm (models/rsgm/rsgm.py)
 * ---------------------------------------------------------------------------------------------- */

/* _linear_interpolate(dmap, n, th)  rsgm.py:66-113  (numba: neighbours are float64 copies of float32 values,
 * m and q float64, store rounds to float32) */
ORC_API void orc_linear_interpolate(float *dmap, int W, int H, int n_arg, float th)
{
    const int n = n_arg / 2;
    for (int y = 0; y < H; y++) {
        float *row = dmap + (size_t)y * W;
        for (int x = 0; x < W; x++) {
            if (row[x] <= 0) {
                double nl = 0, nr = 0;
                int nlx = 0, nrx = 0;
                for (int xw = -1; xw >= -n; xw--)
                    if (x + xw >= 0 && x + xw < W && row[x + xw] > 0) { nl = row[x + xw]; nlx = xw; break; }
                for (int xw = 1; xw <= n; xw++)
                    if (x + xw >= 0 && x + xw < W && row[x + xw] > 0) { nr = row[x + xw]; nrx = xw; break; }
                if (nl > 0 && nr > 0 && fabs(nl - nr) < (double)th) {
                    double m = (nr - nl) / (double)(nrx - nlx);
                    double q = nl - m * (double)nlx;
                    for (int xw = nlx; xw <= nrx; xw++) row[x + xw] = (float)(m * (double)xw + q);
                }
            }
        }
    }
}

/* numba round() on float32 -> int: round half to even */
static int round_half_even(float v) { return (int)nearbyint((double)v); }

/* _left_right_check rsgm.py:229-248 */
ORC_API void orc_lr_check(const float *dl, const float *dr, uint8_t *mask, int W, int H, float th)
{
    memset(mask, 0, (size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float v = dl[(size_t)y * W + x];
            if (v > 0) {
                int d = round_half_even(v), xd = x - d;
                if (xd >= 0 && xd <= W - 1) {
                    float r = dr[(size_t)y * W + xd];
                    if (r > 0) mask[(size_t)y * W + x] = fabsf(v - r) > th ? 128 : 255;
                } else {
                    mask[(size_t)y * W + x] = 128;
                }
            }
        }
}

/* _interpolate_background rsgm.py:184-227 */
ORC_API void orc_interpolate_background(float *dmap, int W, int H)
{
    for (int v = 0; v < H; v++) {
        float *row = dmap + (size_t)v * W;
        int count = 0;
        for (int u = 0; u < W; u++) {
            if (row[u] > 0) {
                if (count >= 1) {
                    int u1 = u - count, u2 = u - 1;
                    if (u1 > 0 && u2 < W - 1) {
                        float d = row[u1 - 1] < row[u2 + 1] ? row[u1 - 1] : row[u2 + 1];
                        for (int c = u1; c <= u2; c++) row[c] = d;
                    }
                }
                count = 0;
            } else {
                count++;
            }
        }
        for (int u = 0; u < W; u++)
            if (row[u] > 0) { for (int u2 = 0; u2 < u; u2++) row[u2] = row[u]; break; }
        for (int u = W - 1; u >= 0; u--)
            if (row[u] > 0) { for (int u2 = u + 1; u2 < W; u2++) row[u2] = row[u]; break; }
    }
    for (int u = 0; u < W; u++) {
        for (int v = 0; v < H; v++)
            if (dmap[(size_t)v * W + u] > 0) { for (int v2 = 0; v2 < v; v2++) dmap[(size_t)v2 * W + u] = dmap[(size_t)v * W + u]; break; }
        for (int v = H - 1; v >= 0; v--)
            if (dmap[(size_t)v * W + u] > 0) { for (int v2 = v + 1; v2 < H; v2++) dmap[(size_t)v2 * W + u] = dmap[(size_t)v * W + u]; break; }
    }
}

/* _guided_dsi rsgm.py:115-127.  numba fuses `k * (1-np.exp((-(hints[y,x]-np.arange(dmax))**2)/(2*c**2)))` into one
 * element loop evaluated in float64 whose RESULT ARRAY is float32 (float32 scalar with int64 array), verified against
 * numba 0.65: the weight is rounded to float32 once, then multiplies the float64 cost; the cast back to uint16
 * truncates (k=10, c=1). */
ORC_API void orc_guided_dsi(uint16_t *dsi, const float *hints, const float *valid, int W, int H, int D)
{
    for (size_t p = 0; p < (size_t)W * H; p++)
        if (valid[p] > 0)
            for (int d = 0; d < D; d++) {
                double t = (double)hints[p] - (double)d;
                float w = (float)(10.0 * (1.0 - exp(-(t * t) / 2.0)));
                dsi[p * D + d] = (uint16_t)((double)dsi[p * D + d] * (double)w);
            }
}

/* ------------------------------------------------------------------------------------------------
 * compute_rsgm  (models/rsgm/rsgm.py:250-294), whole pipeline on one frame.
 *   left      : guide image for adaptive P2, uint8 [H][W][C]   (colour: first W*H bytes of the PADDED buffer
 *               are read as gray, RSGM/pyrSGM.cpp:586-588)
 *   left_vpp, right_vpp : matching images uint8 [H][W][C], C in {1,3}
 *   out       : float32 [H][W]
 *   stage outputs (optional, may be NULL): padded census L/R, S, are returned for stage-wise tests.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_compute_rsgm(const uint8_t *left, const uint8_t *left_vpp, const uint8_t *right_vpp,
                             int H, int W, int C, int D, int subpixel, const float *rcp_lut,
                             const float *hints, const float *validhints, float *out)
{
    if (D % 8 != 0 || D > 256) return -2;
    const int pad_h = (((H / 16) + 1) * 16 - H) % 16, pad_w = (((W / 16) + 1) * 16 - W) % 16;
    const int pl = pad_w / 2, pr = pad_w - pl, pt = pad_h / 2, pb = pad_h - pt;
    const int Hp = H + pad_h, Wp = W + pad_w;
    const size_t np = (size_t)Hp * Wp;
    uint8_t *lp = (uint8_t *)malloc(np * C), *lvp = (uint8_t *)malloc(np * C), *rvp = (uint8_t *)malloc(np * C);
    orc_pad_reflect(left, lp, H, W, C, pt, pb, pl, pr);
    orc_pad_reflect(left_vpp, lvp, H, W, C, pt, pb, pl, pr);
    orc_pad_reflect(right_vpp, rvp, H, W, C, pt, pb, pl, pr);
    uint8_t *gl = (uint8_t *)malloc(np), *gr = (uint8_t *)malloc(np);
    if (C == 3) { orc_rgb2gray(lvp, gl, (int)np); orc_rgb2gray(rvp, gr, (int)np); }
    else { memcpy(gl, lvp, np); memcpy(gr, rvp, np); }
    uint32_t *cl = (uint32_t *)malloc(np * 4), *cr = (uint32_t *)malloc(np * 4);
    orc_census5x5(gl, cl, Wp, Hp);
    orc_census5x5(gr, cr, Wp, Hp);
    uint16_t *dsi = (uint16_t *)malloc(np * D * 2), *S = (uint16_t *)malloc(np * D * 2);
    orc_cost_census(cl, cr, dsi, Wp, Hp, D);
    if (hints && validhints) {
        float *hp = (float *)calloc(np, 4), *vp = (float *)calloc(np, 4);
        for (int y = 0; y < H; y++) {
            memcpy(hp + (size_t)(y + pt) * Wp + pl, hints + (size_t)y * W, (size_t)W * 4);
            memcpy(vp + (size_t)(y + pt) * Wp + pl, validhints + (size_t)y * W, (size_t)W * 4);
        }
        orc_guided_dsi(dsi, hp, vp, Wp, Hp, D);
        free(hp); free(vp);
    }
    orc_sgm_aggregate(lp /* first Wp*Hp bytes */, dsi, S, Wp, Hp, D, 7, 17, 0.25f, 50);
    float *dl = (float *)malloc(np * 4), *dlf = (float *)malloc(np * 4), *dr = (float *)malloc(np * 4), *drf = (float *)malloc(np * 4);
    orc_wta_left(S, dl, Wp, Hp, D);
    orc_subpixel(S, dl, Wp, Hp, D, 0, rcp_lut);
    orc_median3x3(dl, dlf, Wp, Hp);
    orc_linear_interpolate(dlf, Wp, Hp, 15, 3.0f);
    orc_wta_right(S, dr, Wp, Hp, D);
    orc_median3x3(dr, drf, Wp, Hp);
    orc_linear_interpolate(drf, Wp, Hp, 15, 3.0f);
    for (size_t p = 0; p < np; p++) { if (dlf[p] < 0) dlf[p] = 0; if (drf[p] < 0) drf[p] = 0; }
    /* crop */
    float *fl = (float *)malloc((size_t)H * W * 4), *fr = (float *)malloc((size_t)H * W * 4);
    for (int y = 0; y < H; y++) {
        memcpy(fl + (size_t)y * W, dlf + (size_t)(y + pt) * Wp + pl, (size_t)W * 4);
        memcpy(fr + (size_t)y * W, drf + (size_t)(y + pt) * Wp + pl, (size_t)W * 4);
    }
    uint8_t *mask = (uint8_t *)malloc((size_t)H * W), *u8 = (uint8_t *)malloc((size_t)H * W);
    orc_lr_check(fl, fr, mask, W, H, 1.0f);
    for (size_t p = 0; p < (size_t)H * W; p++) {
        float v = mask[p] == 128 ? 0.0f : fl[p];
        u8[p] = (uint8_t)v;                                   /* astype(uint8): values are in [0,256) */
    }
    orc_filter_speckles_u8(u8, H, W, 0, 200, 10);
    for (size_t p = 0; p < (size_t)H * W; p++) {
        float v = (float)u8[p];
        if (subpixel && v != 0) v = fl[p];                    /* rsgm.py:289-290 */
        out[p] = v;
    }
    orc_interpolate_background(out, W, H);
    free(lp); free(lvp); free(rvp); free(gl); free(gr); free(cl); free(cr); free(dsi); free(S);
    free(dl); free(dlf); free(dr); free(drf); free(fl); free(fr); free(mask); free(u8);
    return 0;
}
