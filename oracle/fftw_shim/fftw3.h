/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the five single-precision
 * FFTW3 entry points that the reference's DCT handler calls
 * (reference src/nlkalman.c:201-220, :239-242, :278, :355).  FFTW is an
 * un-vendored, unpinned system dependency of the reference (libfftw3-dev,
 * README.md:25) and is not installed in this image, so the reference cannot be
 * linked against the real thing here.  This header and fftw3_shim.c are
 * written from FFTW's published r2r definitions:
 *
 *   REDFT10 (DCT-II):  Y[k] = 2 * sum_{j=0}^{n-1} X[j] cos(pi (j+1/2) k / n)
 *   REDFT01 (DCT-III): Y[k] = X[0] + 2 * sum_{j=1}^{n-1} X[j] cos(pi j (k+1/2) / n)
 *
 * applied separably over the dimensions of a rank-d transform.  Nothing in the
 * product (bwd_nlkalman_b200/) includes or links this file.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2,
    FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6,
    FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} fftwf_r2r_kind;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

struct fftwf_shim_plan;
typedef struct fftwf_shim_plan *fftwf_plan;

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);

fftwf_plan fftwf_plan_many_r2r(int rank, const int *n, int howmany,
                               float *in, const int *inembed, int istride, int idist,
                               float *out, const int *onembed, int ostride, int odist,
                               const fftwf_r2r_kind *kind, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

#ifdef __cplusplus
}
#endif
#endif
