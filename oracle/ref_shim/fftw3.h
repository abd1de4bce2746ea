/* Stand-in for <fftw3.h>, TEST INFRASTRUCTURE ONLY.
 *
 * libfftw3 is not installed in this image and there is no network, but the
 * reference's input_sdr.c / sdr_sync.c only use five FFTW entry points
 * (fftw_plan_dft_1d, fftw_execute, fftw_destroy_plan, fftw_malloc, fftw_free)
 * at sizes 2048 (forward), 1536 and 128 (backward), unnormalised.  The DFT is
 * mathematically defined, so a float64 mixed-radix DFT is an exact stand-in to
 * ~1e-13 relative.  This header is only put on the include path when the
 * UNMODIFIED reference sources are compiled into oracle/_ref (see Makefile).
 * Nothing in the product (dabtools_b200/, include/) includes it.
 */
#ifndef DABGPU_ORACLE_FFTW3_SHIM_H
#define DABGPU_ORACLE_FFTW3_SHIM_H

#include <stddef.h>

typedef double fftw_complex[2];

struct dabshim_plan;
typedef struct dabshim_plan *fftw_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out,
                           int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);

#endif
