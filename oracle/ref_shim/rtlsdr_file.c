/* rtlsdr_file.c -- TEST INFRASTRUCTURE ONLY: a file-backed librtlsdr for the unmodified dab2eti.c.
 *
 *   RTLSDR_FILE=<uint8 IQ capture> ./dab2eti <frequency>  >  out.eti
 *
 * rtlsdr_read_async() delivers the capture in buf_len-byte callbacks and returns at its end.
 * dab2eti.c hands a buffer to its demodulator thread without waiting for it (dab2eti.c:117-130:
 * memcpy + sem_post), which a real dongle paces at 64 ms per buffer; a replay has to pace itself
 * instead.  The shim therefore interposes sem_wait(): the demodulator thread calls it when it is
 * done with a buffer (dab2eti.c:67), and the producer waits for that before the next callback, so
 * the run is deterministic and as fast as the receiver.
 * rtlsdr_set_center_freq() after the first one is the tuner feedback (dab2eti.c:76-103): the
 * samples that follow are re-mixed by the accumulated offset and re-quantised, the same virtual
 * tuner as oracle/ref_harness.c.  The Wavefinder entry points only exist so that the file links;
 * wf_open() fails like on a machine without the device (dab2eti.c:293). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <semaphore.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtl-sdr.h"

struct rtlsdr_dev {
  FILE *f;
  uint32_t f0, freq;
  int have_f0, cancel;
  long pos;
};
static struct rtlsdr_dev g_dev;

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
/* sem_wait() entries per semaphore: other threads of the process (the CUDA driver's, a profiler's) use
 * semaphores too, so only the ones on dab2eti's `data_ready` may count.  Its address is learnt from the
 * sem_post() the callback makes on the producer thread (dab2eti.c:127-129). */
static struct { sem_t *s; long waits; } g_sems[1024];
static sem_t *g_data_ready;
static pthread_t g_producer;
static int g_have_producer;

static long waits_on(sem_t *s) {
  for (int i = 0; i < 1024 && g_sems[i].s; i++)
    if (g_sems[i].s == s) return g_sems[i].waits;
  return 0;
}

int sem_wait(sem_t *s) {
  static int (*real)(sem_t *);
  if (!real) real = (int (*)(sem_t *))dlsym(RTLD_NEXT, "sem_wait");
  pthread_mutex_lock(&g_mu);
  for (int i = 0; i < 1024; i++) {
    if (g_data_ready && s != g_data_ready) break;  /* once known, only dab2eti's semaphore is tracked */
    if (g_sems[i].s == s || !g_sems[i].s) {
      g_sems[i].s = s;
      g_sems[i].waits++;
      break;
    }
  }
  pthread_cond_broadcast(&g_cv);
  pthread_mutex_unlock(&g_mu);
  return real(s);
}

int sem_post(sem_t *s) {
  static int (*real)(sem_t *);
  if (!real) real = (int (*)(sem_t *))dlsym(RTLD_NEXT, "sem_post");
  if (g_have_producer && pthread_equal(pthread_self(), g_producer)) {
    pthread_mutex_lock(&g_mu);
    g_data_ready = s;
    pthread_mutex_unlock(&g_mu);
  }
  return real(s);
}

uint32_t rtlsdr_get_device_count(void) { return getenv("RTLSDR_FILE") ? 1u : 0u; }
const char *rtlsdr_get_device_name(uint32_t i) { (void)i; return "IQ file replay"; }
int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s) {
  (void)i;
  strcpy(m, "oracle");
  strcpy(p, "rtlsdr_file");
  strcpy(s, "0");
  return 0;
}
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index) {
  (void)index;
  const char *path = getenv("RTLSDR_FILE");
  memset(&g_dev, 0, sizeof g_dev);
  g_dev.f = path ? fopen(path, "rb") : NULL;
  if (!g_dev.f) return -1;
  *dev = &g_dev;
  return 0;
}
int rtlsdr_close(rtlsdr_dev_t *dev) {
  if (dev && dev->f) fclose(dev->f);
  if (dev) dev->f = NULL;
  return 0;
}
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq) {
  if (!dev->have_f0) {
    dev->f0 = freq;
    dev->have_f0 = 1;
  }
  dev->freq = freq;
  return 0;
}
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *dev, int *gains) { (void)dev; gains[0] = 0; return 1; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *dev, int gain) { (void)dev; (void)gain; return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual) { (void)dev; (void)manual; return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate) { (void)dev; return rate == 2048000 ? 0 : -1; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev) { (void)dev; return 0; }
int rtlsdr_cancel_async(rtlsdr_dev_t *dev) { if (dev) dev->cancel = 1; return 0; }

int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len) {
  (void)buf_num;
  unsigned char *buf = malloc(buf_len);
  g_producer = pthread_self();
  g_have_producer = 1;
  long delivered = 0;
  while (!dev->cancel && fread(buf, 1, buf_len, dev->f) == buf_len) {
    const double df = (double)dev->freq - (double)dev->f0;
    if (df != 0.0) {
      for (uint32_t b = 0; b < buf_len; b += 2) {
        const double n = (double)((dev->pos + b) / 2);
        const double ph = -2.0 * M_PI * df * n / 2048000.0;
        const double c = cos(ph), s = sin(ph);
        const double xr = (double)buf[b] - 127.0, xi = (double)buf[b + 1] - 127.0;
        const double qr = floor(xr * c - xi * s + 0.5) + 127.0, qi = floor(xr * s + xi * c + 0.5) + 127.0;
        buf[b] = (unsigned char)(qr < 0 ? 0 : qr > 255 ? 255 : qr);
        buf[b + 1] = (unsigned char)(qi < 0 ? 0 : qi > 255 ? 255 : qi);
      }
    }
    dev->pos += buf_len;
    cb(buf, buf_len, ctx);   /* memcpy + sem_post(&data_ready): tells us which semaphore it is */
    delivered++;
    /* The consumer enters sem_wait(&data_ready) once before its first buffer and once after each
     * buffer it has processed: buffer k is done when it has entered delivered + 1 times. */
    pthread_mutex_lock(&g_mu);
    while (!g_data_ready || waits_on(g_data_ready) < 1 + delivered) pthread_cond_wait(&g_cv, &g_mu);
    pthread_mutex_unlock(&g_mu);
  }
  free(buf);
  fflush(stdout);
  return 0;
}

/* ---- Wavefinder entry points referenced by dab2eti.c (input_wf.h:15-18): device absent ---- */
struct wavefinder_t;
struct demapped_transmission_frame_t;
int wf_open(struct wavefinder_t *wf, char *filename) { (void)wf; (void)filename; return -1; }
int wf_init(struct wavefinder_t *wf) { (void)wf; return -1; }
void wf_tune(struct wavefinder_t *wf, int freq_khz) { (void)wf; (void)freq_khz; }
int wf_read_frame(struct wavefinder_t *wf, struct demapped_transmission_frame_t *tf) { (void)wf; (void)tf; return -1; }
