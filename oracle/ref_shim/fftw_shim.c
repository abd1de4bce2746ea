/* Double-precision DFT behind the fftw3.h stand-in (TEST INFRASTRUCTURE ONLY).
 *
 * Sizes the reference asks for: 2048 and 128 (powers of two) and 1536 = 3*512.
 * Powers of two run an iterative radix-2 FFT; n = 3*2^k runs three 2^k FFTs
 * over the residues mod 3 followed by a radix-3 combine; anything else falls
 * back to the O(n^2) definition.  Twiddles are a cos/sin table in double, so
 * the output equals the mathematical (unnormalised) DFT to ~1e-13 relative,
 * which is what FFTW would return up to rounding.
 */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct dabshim_plan {
  int n;
  int sign;
  fftw_complex *in;
  fftw_complex *out;
  double *tw; /* tw[2k], tw[2k+1] = cos, sin of sign*2*pi*k/n */
};

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* in-place radix-2 DIT on interleaved re/im; w is the table of the *full* plan
 * size N with N % n == 0, so W_n^k = w[k * (N/n)] */
static void fft_pow2(double *x, int n, const double *w, int N) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double tr = x[2 * i], ti = x[2 * i + 1];
      x[2 * i] = x[2 * j];
      x[2 * i + 1] = x[2 * j + 1];
      x[2 * j] = tr;
      x[2 * j + 1] = ti;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    int step = N / len;
    for (int base = 0; base < n; base += len) {
      for (int k = 0; k < half; k++) {
        double wr = w[2 * k * step], wi = w[2 * k * step + 1];
        double *a = x + 2 * (base + k), *b = x + 2 * (base + k + half);
        double br = b[0] * wr - b[1] * wi, bi = b[0] * wi + b[1] * wr;
        b[0] = a[0] - br;
        b[1] = a[1] - bi;
        a[0] += br;
        a[1] += bi;
      }
    }
  }
}

/* The reference creates and destroys a plan around every transform, so the
 * cos/sin table is cached per (n, sign) and per thread instead of rebuilt. */
#define TW_CACHE 8
static __thread struct { int n, sign; double *tw; } tw_cache[TW_CACHE];

static double *twiddles_for(int n, int sign) {
  int slot = -1;
  for (int i = 0; i < TW_CACHE; i++) {
    if (tw_cache[i].tw && tw_cache[i].n == n && tw_cache[i].sign == sign) return tw_cache[i].tw;
    if (!tw_cache[i].tw && slot < 0) slot = i;
  }
  double *tw = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  for (int k = 0; k < n; k++) {
    double a = (double)sign * 2.0 * M_PI * (double)k / (double)n;
    tw[2 * k] = cos(a);
    tw[2 * k + 1] = sin(a);
  }
  if (slot >= 0) {
    tw_cache[slot].n = n;
    tw_cache[slot].sign = sign;
    tw_cache[slot].tw = tw;
  }
  return tw; /* leaked only if more than TW_CACHE distinct sizes are used */
}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out,
                           int sign, unsigned flags) {
  (void)flags;
  struct dabshim_plan *p = (struct dabshim_plan *)malloc(sizeof *p);
  p->n = n;
  p->sign = sign;
  p->in = in;
  p->out = out;
  p->tw = twiddles_for(n, sign);
  return p;
}

void fftw_execute(const fftw_plan p) {
  const int n = p->n;
  const double *w = p->tw;
  double *src = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  double *dst = (double *)p->out;
  memcpy(src, p->in, sizeof(double) * 2 * (size_t)n);

  if (is_pow2(n)) {
    fft_pow2(src, n, w, n);
    memcpy(dst, src, sizeof(double) * 2 * (size_t)n);
  } else if (n % 3 == 0 && is_pow2(n / 3)) {
    const int m = n / 3;
    double *sub = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (int q = 0; q < 3; q++) {
      for (int k = 0; k < m; k++) {
        sub[2 * (q * m + k)] = src[2 * (3 * k + q)];
        sub[2 * (q * m + k) + 1] = src[2 * (3 * k + q) + 1];
      }
      /* W_m^k = W_n^{3k}: table of size n, stride 3 == N/m */
      fft_pow2(sub + 2 * (size_t)q * m, m, w, n);
    }
    for (int k = 0; k < m; k++) {
      for (int r = 0; r < 3; r++) {
        int kk = k + r * m; /* output bin */
        double sr = sub[2 * k], si = sub[2 * k + 1];
        for (int q = 1; q < 3; q++) {
          int idx = (int)(((long)q * kk) % n);
          double wr = w[2 * idx], wi = w[2 * idx + 1];
          double xr = sub[2 * (q * m + k)], xi = sub[2 * (q * m + k) + 1];
          sr += xr * wr - xi * wi;
          si += xr * wi + xi * wr;
        }
        dst[2 * kk] = sr;
        dst[2 * kk + 1] = si;
      }
    }
    free(sub);
  } else {
    for (int k = 0; k < n; k++) {
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < n; j++) {
        int idx = (int)(((long)j * k) % n);
        sr += src[2 * j] * w[2 * idx] - src[2 * j + 1] * w[2 * idx + 1];
        si += src[2 * j] * w[2 * idx + 1] + src[2 * j + 1] * w[2 * idx];
      }
      dst[2 * k] = sr;
      dst[2 * k + 1] = si;
    }
  }
  free(src);
}

void fftw_destroy_plan(fftw_plan p) {
  free(p); /* the twiddle table belongs to the per-thread cache */
}

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }
