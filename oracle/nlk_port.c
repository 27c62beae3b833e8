/* TEST INFRASTRUCTURE ONLY (oracle/) -- see nlk_port.h.
 *
 * Plain-C restatement of the reference's per-frame path.  Each function cites
 * the reference lines (src/nlkalman.c unless noted) whose behaviour it follows.
 * It is organised the way the CUDA path is organised -- (1) search every grid
 * patch, (2) resolve the order-dependent "already processed" mask in raster
 * order, (3) filter the active groups and aggregate, (4) normalise -- which is
 * equivalent to the reference's single loop run with one thread, because the
 * search of a patch never depends on the mask.
 *
 * Arithmetic is fp32 with the reference's evaluation order (compiled with
 * -ffp-contract=off, no fast-math):  patch distances are a sequential
 * sum over (hy, hx, c) of separately rounded squares, statistics are Welford
 * updates in sorted-candidate order, the posterior variance is a sequential sum
 * over (n, c, hy, hx).  The DCT is the exact orthonormal DCT-II/III evaluated in
 * double and rounded once (the reference goes through FFTW + fp32 rescaling).
 */
#include "nlk_port.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ---- colour transform (:92-130) ------------------------------------------------ */

void port_rgb2opp(float *im, int w, int h, int ch)
{
    if (ch != 3) return; /* :94 */
    const float a = 1.f / sqrtf(3.f), b = 1.f / sqrtf(2.f), c = 2.f * a * sqrtf(2.f);
    for (long k = 0; k < (long)w * h; ++k) {
        float *p = im + 3 * k;
        const float Y = a * (p[0] + p[1] + p[2]);
        const float U = b * (p[0] - p[2]);
        const float V = c * (0.25f * p[0] - 0.5f * p[1] + 0.25f * p[2]);
        p[0] = Y; p[1] = U; p[2] = V;
    }
}

void port_opp2rgb(float *im, int w, int h, int ch)
{
    if (ch != 3) return; /* :114 */
    const float a = 1.f / sqrtf(3.f), b = 1.f / sqrtf(2.f), c = a / b;
    for (long k = 0; k < (long)w * h; ++k) {
        float *p = im + 3 * k;
        const float R = a * p[0] + b * p[1] + 0.5f * c * p[2];
        const float G = a * p[0] - c * p[2];
        const float B = a * p[0] - b * p[1] + 0.5f * c * p[2];
        p[0] = R; p[1] = G; p[2] = B;
    }
}

/* ---- bicubic warp with NaN outside / occluded (:29-88) ---------------------------- */

/* Keys a=-1/2 cubic in Horner form; the literals are double in the reference (:38-40),
 * so the polynomial is evaluated in double and rounded to float on return */
static float cubic1(const float v[4], float x)
{
    const double xd = x;
    return (float)(v[1] + 0.5 * xd * (v[2] - v[0]
                 + xd * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3]
                 + xd * (3.0 * (v[1] - v[2]) + v[3] - v[0]))));
}

void port_warp_bicubic(float *imw, const float *im, const float *of, const float *msk,
                       int w, int h, int ch)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float *o = imw + ((long)y * w + x) * ch;
            if (msk && msk[(long)y * w + x] != 0) { /* :77, :84-85 */
                for (int c = 0; c < ch; ++c) o[c] = NAN;
                continue;
            }
            float xw = x + of[((long)y * w + x) * 2 + 0]; /* :79-80 */
            float yw = y + of[((long)y * w + x) * 2 + 1];
            xw -= 1; yw -= 1;                             /* :56-57 */
            const int ix = (int)floor(xw), iy = (int)floor(yw);
            const float fx = xw - ix, fy = yw - iy;
            for (int c = 0; c < ch; ++c) {
                float col[4];
                for (int i = 0; i < 4; ++i) { /* i: x tap, j: y tap (:64-66, :45-50) */
                    float tap[4];
                    for (int j = 0; j < 4; ++j) {
                        const int sx = ix + i, sy = iy + j;
                        tap[j] = (sx < 0 || sx >= w || sy < 0 || sy >= h)
                               ? NAN : im[((long)sy * w + sx) * ch + c]; /* :32 */
                    }
                    col[i] = cubic1(tap, fy);
                }
                o[c] = cubic1(col, fx);
            }
        }
}

/* ---- default parameters (:426-487) -------------------------------------------- */

void port_default_params(port_params *p, float sigma, int mode)
{
    if (p->patch_sz < 0) p->patch_sz = 8;
    if (p->search_sz_x < 0) p->search_sz_x = 10;
    if (p->search_sz_t < 0) p->search_sz_t = 5;
    if (p->dista_lambda < 0) p->dista_lambda = 1.0f;
    switch (mode) {
    case PORT_FLT1: /* :458-464 */
        if (p->npatches_x < 0) p->npatches_x = (int)(0.5 * sigma + 40.);
        if (p->beta_x < 0) p->beta_x = (float)(-0.04 * sigma + 3.91);
        if (p->npatches_t < 0) p->npatches_t = 30;
        if (p->npatches_tagg < 0) p->npatches_tagg = 20;
        if (p->beta_t < 0) p->beta_t = (float)(-0.005 * sigma + 2.05);
        break;
    case PORT_FLT2: /* :466-472 */
        if (p->npatches_x < 0) p->npatches_x = (int)(0.5 * sigma + 10.);
        if (p->beta_x < 0) p->beta_x = (float)(0.004 * sigma + 0.21);
        if (p->npatches_t < 0) p->npatches_t = (int)(5.f > sigma ? 5.f : sigma);
        if (p->npatches_tagg < 0) p->npatches_tagg = 1;
        if (p->beta_t < 0) p->beta_t = (float)(0.014 * sigma + 1.38);
        break;
    case PORT_SMO1: /* :474-480 */
        if (p->npatches_x < 0) p->npatches_x = 0;
        if (p->beta_x < 0) p->beta_x = 0;
        if (p->npatches_t < 0) {
            const float v = 3 * sigma - 15;
            p->npatches_t = (int)(5.f > v ? 5.f : v);
        }
        if (p->npatches_tagg < 0) p->npatches_tagg = p->npatches_t;
        if (p->beta_t < 0) {
            const double v = -0.14 * sigma + 8.0;
            p->beta_t = (float)(1.0 > v ? 1.0 : v);
        }
        break;
    }
}

/* ---- aggregation window (:365-419, "gaussian" :401-407, outer product :413-416) ---- */

void port_window(float *w2, int psz)
{
    float w1[64];
    const float N = (float)psz;
    const float N2 = (float)((N - 1.) / 2.);
    for (int n = 0; n < psz; ++n) {
        const float s = .4f;
        const float x = ((float)n - N2) / N2 / s;
        w1[n] = (float)exp(-.5 * x * x);
    }
    for (int i = 0; i < psz; ++i)
        for (int j = 0; j < psz; ++j) w2[i * psz + j] = w1[i] * w1[j];
}

/* ---- orthonormal 2-D DCT-II / DCT-III of n psz x psz tiles (:248-360) -------------- */

static void dct_table(double *t, int n) /* t[k*n+j] = c(k) sqrt(2/n) cos(pi (j+1/2) k / n) */
{
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            t[k * n + j] = sqrt((k ? 2.0 : 1.0) / n) * cos(pi * (j + 0.5) * k / n);
}

static void dct2_tiles(float *tiles, int psz, int n, int inverse, const double *t)
{
    double a[64 * 64], b[64 * 64];
    const int pp = psz * psz;
    for (int s = 0; s < n; ++s) {
        float *x = tiles + (long)s * pp;
        for (int i = 0; i < pp; ++i) a[i] = x[i];
        /* rows */
        for (int y = 0; y < psz; ++y)
            for (int k = 0; k < psz; ++k) {
                double acc = 0;
                for (int j = 0; j < psz; ++j)
                    acc += (inverse ? t[j * psz + k] : t[k * psz + j]) * a[y * psz + j];
                b[y * psz + k] = acc;
            }
        /* columns */
        for (int xx = 0; xx < psz; ++xx)
            for (int k = 0; k < psz; ++k) {
                double acc = 0;
                for (int j = 0; j < psz; ++j)
                    acc += (inverse ? t[j * psz + k] : t[k * psz + j]) * b[j * psz + xx];
                a[k * psz + xx] = acc;
            }
        for (int i = 0; i < pp; ++i) x[i] = (float)a[i];
    }
}

void port_dct2(float *tiles, int psz, int n, int inverse)
{
    double t[64 * 64];
    dct_table(t, psz);
    dct2_tiles(tiles, psz, n, inverse, t);
}

/* ---- one pass ---------------------------------------------------------------------- */

typedef struct { float d; int idx; int x, y; } cand_t;

static int cand_cmp(const void *a, const void *b)
{
    /* the reference compares d only (:500-505) and relies on glibc's stable merge
     * sort; the scan index as second key states that order explicitly */
    const cand_t *p = (const cand_t *)a, *q = (const cand_t *)b;
    if (p->d < q->d) return -1;
    if (p->d > q->d) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

/* valid(q): prev0 given and no NaN in channel 0 of the patch at q (:605-609, :725-730) */
static int patch_valid(const float *prev0, int w, int ch, int psz, int qx, int qy)
{
    if (!prev0) return 0;
    for (int hy = 0; hy < psz; ++hy)
        for (int hx = 0; hx < psz; ++hx)
            if (isnan(prev0[((long)(qy + hy) * w + qx + hx) * ch])) return 0;
    return 1;
}

typedef struct {
    int nk;      /* kept candidates (0: no search was made) */
    int np0;     /* kept candidates whose previous patch is valid */
    int prev_p;
} group_hdr;

/* search for the patch at (px,py): all distances in the clamped window, stable
 * ascending sort, keep k (:630-707, :1521-1597).  cands must hold (2r+1)^2. */
static void search_patch(const float *S, const float *prev0, int w, int h, int ch, int psz,
                         int px, int py, int mode, const port_params *prm,
                         cand_t *cands, group_hdr *hdr, int *kx, int *ky, float *kd,
                         unsigned char *kprev)
{
    const int prev_p = patch_valid(prev0, w, ch, psz, px, py);
    hdr->prev_p = prev_p;
    hdr->nk = 0;
    hdr->np0 = 0;
    int k = prev_p ? prm->npatches_t : prm->npatches_x; /* :630, :1521 */
    if (k <= 1) return;
    const int r = (mode == PORT_PASS_SMOOTH) ? prm->search_sz_t                 /* :1527 */
                : (prev_p ? prm->search_sz_t : prm->search_sz_x);              /* :637 */
    const int x0 = imax(px - r, 0), x1 = imin(px + r, w - psz) + 1;
    const int y0 = imax(py - r, 0), y1 = imin(py + r, h - psz) + 1;
    const float npix = (float)psz * psz * ch;
    int n = 0;
    for (int qy = y0; qy < y1; ++qy)
        for (int qx = x0; qx < x1; ++qx, ++n) {
            float ww = 0;
            for (int hy = 0; hy < psz; ++hy)
                for (int hx = 0; hx < psz; ++hx)
                    for (int c = 0; c < ch; ++c) {
                        const float e = S[((long)(qy + hy) * w + qx + hx) * ch + c]
                                      - S[((long)(py + hy) * w + px + hx) * ch + c];
                        ww += e * e; /* :687-692 (dista_sigma2 = 0, :629) */
                    }
            const float d = ww / npix;
            cands[n].d = d > 0 ? d : 0; /* :701 */
            cands[n].idx = n;
            cands[n].x = qx;
            cands[n].y = qy;
        }
    qsort(cands, n, sizeof *cands, cand_cmp);
    k = imin(k, n);
    hdr->nk = k;
    for (int i = 0; i < k; ++i) {
        kx[i] = cands[i].x;
        ky[i] = cands[i].y;
        kd[i] = cands[i].d;
        kprev[i] = prev_p && patch_valid(prev0, w, ch, psz, kx[i], ky[i]); /* :723-732 */
        hdr->np0 += kprev[i];
    }
}

/* gather a patch into planar tiles [ch][psz][psz] (:734-741) */
static void gather(float *dst, const float *img, int w, int ch, int psz, int qx, int qy)
{
    for (int c = 0; c < ch; ++c)
        for (int hy = 0; hy < psz; ++hy)
            for (int hx = 0; hx < psz; ++hx)
                dst[(c * psz + hy) * psz + hx] = img[((long)(qy + hy) * w + qx + hx) * ch + c];
}

typedef struct {
    int nagg;
    float wgt;
    float vp;
    int marks;
} group_out;

/* statistics, gain, update, inverse transform for one group (:713-911, :1600-1826).
 * PGout receives nagg pixel-domain patches [nagg][ch][psz][psz], gx/gy their coords. */
static void filter_group(int mode, const float *in1, const float *prev0, const float *bsic1,
                         int w, int ch, int psz, int px, int py, float sigma,
                         const port_params *prm, const double *tab,
                         const group_hdr *hdr, const int *kx, const int *ky,
                         const unsigned char *kprev,
                         float *work, float *PGout, int *gx, int *gy, group_out *res)
{
    const int pp = psz * psz, cpp = ch * pp;
    const int tagg = prm->npatches_tagg;
    const float sigma2 = sigma * sigma;
    const float *S = bsic1 ? bsic1 : in1;
    float *M0 = work, *M0V = M0 + cpp, *V0 = M0V + cpp, *V01 = V0 + cpp, *M1 = V01 + cpp,
          *V1 = M1 + cpp, *ND = V1 + cpp /* [2ch][pp] */;
    float *PG = PGout;            /* current-frame group (PG / PG1) */
    float *PG0 = ND + 2 * cpp;    /* smoother: previous-frame group [tagg][cpp] */
    memset(work, 0, sizeof(float) * 6 * cpp);

    res->nagg = 0;
    res->vp = 0;
    res->wgt = 0;
    res->marks = 0;

    const int k = hdr->nk;
    int np0 = 0, np1 = 0;
    if (k > 1) {
        for (int i = 0; i < k; ++i) {
            const int qx = kx[i], qy = ky[i], prev = kprev[i];
            gather(ND, S, w, ch, psz, qx, qy);
            if (prev) gather(ND + cpp, prev0, w, ch, psz, qx, qy);
            else memset(ND + cpp, 0, sizeof(float) * cpp);
            dct2_tiles(ND, psz, 2 * ch, 0, tab); /* :744 */
            np1++;
            np0 += prev;
            const float inp0 = prev ? (float)(1. / (float)np0) : 0; /* :755-756 */
            const float inp1 = (float)(1. / (float)np1);
            for (int j = 0; j < cpp; ++j) {
                const float p = ND[j];
                const float delta = p - M1[j];
                M1[j] += delta * inp1;               /* :765-766 */
                V1[j] += delta * (p - M1[j]);
                if (prev) {
                    float q = ND[cpp + j];
                    if (mode == PORT_PASS_FILTER) {
                        const float d0 = q - M0V[j]; /* :770-775 */
                        M0V[j] += d0 * inp0;
                        V0[j] += d0 * (q - M0V[j]);
                    } else {
                        const float d0 = q - M0[j];  /* :1654-1659 */
                        M0[j] += d0 * inp0;
                        V0[j] += d0 * (q - M0[j]);
                    }
                    const float t = q - ND[j];
                    V01[j] += t * t;                 /* :777-778 */
                    if (np0 <= tagg) {               /* :780-787, :1663-1671 */
                        if (mode == PORT_PASS_FILTER) M0[j] += (q - M0[j]) * inp0;
                        else PG0[(np0 - 1) * cpp + j] = q;
                        PG[(np0 - 1) * cpp + j] = p; /* replaced below when bsic1 */
                    }
                } else if (mode == PORT_PASS_FILTER && np1 <= tagg) { /* :789-794 */
                    PG[(np1 - 1) * cpp + j] = p;
                }
            }
            if (prev && np0 <= tagg) { gx[np0 - 1] = qx; gy[np0 - 1] = qy; }
            else if (!prev && mode == PORT_PASS_FILTER && np1 <= tagg) { gx[np1 - 1] = qx; gy[np1 - 1] = qy; }
        }
        const float inp0 = np0 ? (float)(1. / (float)np0) : 0; /* :798-811 */
        const float inp1 = (float)(1. / (float)np1);
        for (int j = 0; j < cpp; ++j) {
            V1[j] *= inp1;
            if (np0) { V0[j] *= inp0; V01[j] *= inp0; }
        }
    } else if (mode == PORT_PASS_SMOOTH && hdr->prev_p) {
        /* reference :1699-1730 (single-patch estimate).  The reference aggregates at
         * uninitialised coordinates there (undefined behaviour); we restate the evident
         * intent: the group is the patch at p itself.  Not reachable with the CLI
         * defaults (needs --s1_nt <= 1). */
        np0 = 1;
        gather(ND, S, w, ch, psz, px, py);
        gather(ND + cpp, prev0, w, ch, psz, px, py);
        dct2_tiles(ND, psz, 2 * ch, 0, tab);
        for (int j = 0; j < cpp; ++j) {
            const float p = ND[j], q = ND[cpp + j];
            PG[j] = p;
            PG0[j] = q;
            V1[j] = p * p;
            V0[j] = q * q;
            V01[j] = (q - p) * (q - p);
        }
        gx[0] = px; gy[0] = py;
    }
    /* filter, k <= 1: the reference's point-estimate branch (:815-849) never sets
     * np0/np1, so nothing is aggregated for this patch */

    int nagg;
    if (mode == PORT_PASS_FILTER) nagg = imin(np0 ? np0 : np1, tagg); /* :857 */
    else nagg = imin(np0, tagg);                                      /* :1737 */

    /* with a basic estimate the group holds the noisy patches themselves (:785, :853) */
    if (bsic1 && nagg > 0) {
        for (int n = 0; n < nagg; ++n) gather(PG + n * cpp, in1, w, ch, psz, gx[n], gy[n]);
        dct2_tiles(PG, psz, nagg * ch, 0, tab);
    }

    float vp = 0;
    if (mode == PORT_PASS_FILTER) {
        const float bt = prm->beta_t, bx = prm->beta_x;
        const float s2 = bsic1 ? 0.f : sigma2;
        for (int n = 0; n < nagg; ++n)
            for (int j = 0; j < cpp; ++j) {
                if (np0 > 0) { /* :860-880 */
                    const float t = V01[j] - s2;
                    const float v = V0[j] + (t > 0.f ? t : 0.f);
                    const float a = v / (v + bt * sigma2);
                    vp += (1 - a * a) * v + a * a * sigma2;
                    PG[n * cpp + j] = a * PG[n * cpp + j] + (1 - a) * M0[j];
                } else {       /* :883-903 */
                    const float t = V1[j] - s2;
                    const float v = t > 0.f ? t : 0.f;
                    const float a = v / (v + bx * sigma2);
                    vp += a * v;
                    PG[n * cpp + j] = a * PG[n * cpp + j] + (1 - a) * M1[j];
                }
            }
    } else if (np0 > 0) {
        const float b = prm->beta_t;
        for (int n = 0; n < nagg; ++n)
            for (int j = 0; j < cpp; ++j) { /* :1763-1777 */
                const float a = V1[j] / (V1[j] + b * V01[j]);
                const float t = V0[j] - b * V01[j];
                vp += (1 - a * a) * V1[j] + a * a * (t > 0.f ? t : 0.f);
                PG[n * cpp + j] = (1 - a) * PG[n * cpp + j] + a * PG0[n * cpp + j];
            }
    }
    if (nagg > 0) dct2_tiles(PG, psz, nagg * ch, 1, tab); /* :906, :1793 */

    if (mode == PORT_PASS_SMOOTH && np0 == 0) { /* :1795-1804 */
        nagg = 1;
        gx[0] = px; gy[0] = py;
        gather(PG, in1, w, ch, psz, px, py);
    }
    res->nagg = nagg;
    res->vp = vp;
    res->wgt = 1.f / (vp > 1e-6f ? vp : 1e-6f); /* :911, :1824 */
    if (mode == PORT_PASS_FILTER) res->marks = !(prev0 && !np0); /* :931 */
    else res->marks = np0 ? 1 : 0;                               /* :1844 */
}

void port_pass(int mode, float *out, const float *in1, const float *prev0, const float *bsic1,
               int w, int h, int ch, float sigma, port_params prm, port_dump *dump)
{
    const int psz = prm.patch_sz, step = psz / 2; /* :524-525 */
    const int pp = psz * psz, cpp = ch * pp;
    const int gw = (w - psz) / step + 1, gh = (h - psz) / step + 1;
    const long G = (w >= psz && h >= psz) ? (long)gw * gh : 0;
    const int tagg = imax(prm.npatches_tagg, 1);
    const int kmax = imax(imax(prm.npatches_x, prm.npatches_t), 1);
    const int rmax = imax(prm.search_sz_x, prm.search_sz_t);
    const int ncand = (2 * rmax + 1) * (2 * rmax + 1);
    const float *S = bsic1 ? bsic1 : in1;

    float *aggr = (float *)calloc((size_t)w * h, sizeof(float));
    float W[64 * 64];
    double tab[64 * 64];
    port_window(W, psz);
    dct_table(tab, psz);
    memset(out, 0, sizeof(float) * (size_t)w * h * ch);

    /* (1) search all grid patches */
    group_hdr *hdr = (group_hdr *)calloc((size_t)(G ? G : 1), sizeof *hdr);
    int *kx = (int *)malloc(sizeof(int) * (size_t)(G ? G : 1) * kmax);
    int *ky = (int *)malloc(sizeof(int) * (size_t)(G ? G : 1) * kmax);
    float *kd = (float *)malloc(sizeof(float) * (size_t)(G ? G : 1) * kmax);
    unsigned char *kprev = (unsigned char *)malloc((size_t)(G ? G : 1) * kmax);
#pragma omp parallel
    {
        cand_t *cands = (cand_t *)malloc(sizeof(cand_t) * (size_t)ncand);
#pragma omp for schedule(dynamic, 8)
        for (long g = 0; g < G; ++g) {
            const int px = (int)(g % gw) * step, py = (int)(g / gw) * step;
            search_patch(S, prev0, w, h, ch, psz, px, py, mode, &prm, cands, &hdr[g],
                         kx + g * kmax, ky + g * kmax, kd + g * kmax, kprev + g * kmax);
        }
        free(cands);
    }

    /* (2)+(3): raster order; skip patches already marked (:597-600, :1490-1493) */
    unsigned char *mask = (unsigned char *)calloc((size_t)w * h, 1);
    float *work = (float *)malloc(sizeof(float) * (size_t)(8 * cpp + tagg * cpp));
    float *PG = (float *)malloc(sizeof(float) * (size_t)tagg * cpp);
    int *gx = (int *)malloc(sizeof(int) * tagg), *gy = (int *)malloc(sizeof(int) * tagg);
    for (long g = 0; g < G; ++g) {
        const int px = (int)(g % gw) * step, py = (int)(g / gw) * step;
        const int skip = mask[(long)py * w + px] != 0;
        if (dump) {
            if (dump->nk) dump->nk[g] = hdr[g].nk;
            if (dump->np0) dump->np0[g] = hdr[g].np0;
            if (dump->prev_p) dump->prev_p[g] = (unsigned char)hdr[g].prev_p;
            if (dump->active) dump->active[g] = !skip;
            if (dump->vp) dump->vp[g] = 0;
            for (int i = 0; i < hdr[g].nk && i < dump->kmax; ++i) {
                if (dump->knn_xy) {
                    dump->knn_xy[(g * dump->kmax + i) * 2 + 0] = kx[g * kmax + i];
                    dump->knn_xy[(g * dump->kmax + i) * 2 + 1] = ky[g * kmax + i];
                }
                if (dump->knn_d) dump->knn_d[g * dump->kmax + i] = kd[g * kmax + i];
            }
        }
        if (skip) continue;
        group_out res;
        filter_group(mode, in1, prev0, bsic1, w, ch, psz, px, py, sigma, &prm, tab, &hdr[g],
                     kx + g * kmax, ky + g * kmax, kprev + g * kmax, work, PG, gx, gy, &res);
        if (dump && dump->vp) dump->vp[g] = res.vp;
        for (int n = 0; n < res.nagg; ++n) { /* :916-932, :1829-1845 */
            const int qx = gx[n], qy = gy[n];
            for (int hy = 0; hy < psz; ++hy)
                for (int hx = 0; hx < psz; ++hx) {
                    const long pix = (long)(qy + hy) * w + qx + hx;
                    aggr[pix] += res.wgt * W[hy * psz + hx];
                    for (int c = 0; c < ch; ++c)
                        out[pix * ch + c] += res.wgt * W[hy * psz + hx] * PG[n * cpp + (c * psz + hy) * psz + hx];
                }
            if (res.marks) mask[(long)qy * w + qx] = 1;
        }
    }

    /* (4) normalise (:939-942, :1854-1856) */
    for (long i = 0; i < (long)w * h; ++i)
        for (int c = 0; c < ch; ++c) {
            if (aggr[i] > 1e-6f) out[i * ch + c] /= aggr[i];
            else out[i * ch + c] = in1[i * ch + c];
        }

    free(aggr); free(hdr); free(kx); free(ky); free(kd); free(kprev);
    free(mask); free(work); free(PG); free(gx); free(gy);
}

void port_filter_frame(float *deno1, const float *nisy1, const float *deno0, const float *bsic1,
                       int w, int h, int ch, float sigma, port_params prms)
{
    port_pass(PORT_PASS_FILTER, deno1, nisy1, deno0, bsic1, w, h, ch, sigma, prms, NULL);
}

void port_smooth_frame(float *smoo1, const float *filt1, const float *smoo0, const float *bsic1,
                       int w, int h, int ch, float sigma, port_params prms)
{
    port_pass(PORT_PASS_SMOOTH, smoo1, filt1, smoo0, bsic1, w, h, ch, sigma, prms, NULL);
}
