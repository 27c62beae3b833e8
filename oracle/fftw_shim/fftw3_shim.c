/* TEST INFRASTRUCTURE ONLY (oracle/): table-driven stand-in for the FFTW3f r2r
 * calls made by the reference (see fftw3.h in this directory for the call
 * sites and the definitions restated).  Transforms are evaluated separably in
 * double precision from precomputed cosine tables and rounded once to float,
 * so each output is within 1/2 ulp(float) + O(n * 2^-53) of the exact DCT;
 * real FFTW's single-precision codelets are about 1e-7 relative.
 *
 * Only what the reference uses is implemented: contiguous (stride 1) batches
 * of rank <= 3 with kinds REDFT10 / REDFT01, out-of-place.
 */
#include "fftw3.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define SHIM_MAX_RANK 3

struct fftwf_shim_plan {
    int rank;
    int n[SHIM_MAX_RANK];
    fftwf_r2r_kind kind[SHIM_MAX_RANK];
    int howmany, idist, odist;
    float *in, *out;
    double *tab[SHIM_MAX_RANK]; /* tab[d][k*n+j]: weight of input j in output k */
    double *work;               /* 2 * prod(n) doubles */
    int total;
};

void *fftwf_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64)) return NULL;
    return p;
}

void fftwf_free(void *p) { free(p); }

static double *shim_table(int n, fftwf_r2r_kind kind)
{
    double *t = (double *)malloc(sizeof(double) * (size_t)n * n);
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) {
            if (kind == FFTW_REDFT10)
                t[k * n + j] = 2.0 * cos(pi * (j + 0.5) * k / n);
            else /* FFTW_REDFT01 */
                t[k * n + j] = (j == 0) ? 1.0 : 2.0 * cos(pi * j * (k + 0.5) / n);
        }
    return t;
}

fftwf_plan fftwf_plan_many_r2r(int rank, const int *n, int howmany,
                               float *in, const int *inembed, int istride, int idist,
                               float *out, const int *onembed, int ostride, int odist,
                               const fftwf_r2r_kind *kind, unsigned flags)
{
    (void)flags;
    if (rank < 1 || rank > SHIM_MAX_RANK || inembed || onembed || istride != 1 || ostride != 1) {
        fprintf(stderr, "fftw3 shim: unsupported plan (rank %d, strides %d/%d)\n", rank, istride, ostride);
        exit(1);
    }
    struct fftwf_shim_plan *p = (struct fftwf_shim_plan *)calloc(1, sizeof *p);
    p->rank = rank;
    p->howmany = howmany;
    p->idist = idist;
    p->odist = odist;
    p->in = in;
    p->out = out;
    p->total = 1;
    for (int d = 0; d < rank; ++d) {
        if (kind[d] != FFTW_REDFT10 && kind[d] != FFTW_REDFT01) {
            fprintf(stderr, "fftw3 shim: unsupported r2r kind %d\n", (int)kind[d]);
            exit(1);
        }
        p->n[d] = n[d];
        p->kind[d] = kind[d];
        p->tab[d] = shim_table(n[d], kind[d]);
        p->total *= n[d];
    }
    p->work = (double *)malloc(sizeof(double) * 2 * (size_t)p->total);
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    const int total = p->total;
    for (int s = 0; s < p->howmany; ++s) {
        const float *x = p->in + (size_t)s * p->idist;
        float *y = p->out + (size_t)s * p->odist;
        double *a = p->work, *b = p->work + total;
        for (int i = 0; i < total; ++i) a[i] = x[i];

        /* transform one dimension at a time; the last dimension is contiguous */
        int inner = 1;
        for (int d = p->rank - 1; d >= 0; --d) {
            const int n = p->n[d];
            const int outer = total / (n * inner);
            const double *t = p->tab[d];
            if (n == 1) {
                const double g = t[0];
                for (int i = 0; i < total; ++i) b[i] = g * a[i];
            } else {
                for (int o = 0; o < outer; ++o)
                    for (int k = 0; k < n; ++k)
                        for (int in = 0; in < inner; ++in) {
                            const double *src = a + (size_t)o * n * inner + in;
                            const double *tk = t + (size_t)k * n;
                            double acc = 0.0;
                            for (int j = 0; j < n; ++j) acc += tk[j] * src[(size_t)j * inner];
                            b[(size_t)o * n * inner + (size_t)k * inner + in] = acc;
                        }
            }
            double *tmp = a; a = b; b = tmp;
            inner *= n;
        }
        for (int i = 0; i < total; ++i) y[i] = (float)a[i];
    }
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    for (int d = 0; d < p->rank; ++d) free(p->tab[d]);
    free(p->work);
    free(p);
}
