/* ref_harness.c -- TEST INFRASTRUCTURE ONLY (part of the oracle, never shipped).
 *
 * File/buffer-fed driver for the UNMODIFIED reference sources, compiled in place
 * from /root/reference/src into oracle/_ref/libdabref.so (see oracle/Makefile).
 * It replaces the librtlsdr half of the reference's dab2eti.c, which cannot be
 * built here (no <rtl-sdr.h>):
 *
 *   - ref_run_iq()        replays rtlsdr_callback (dab2eti.c:117-130) and the body
 *                         of demod_thread_fn (dab2eti.c:60-115) over an in-memory
 *                         uint8 IQ capture in 262144-byte chunks, including the
 *                         tuner feedback, which here steers a *virtual tuner*
 *                         (the capture is re-mixed by the accumulated retune
 *                         offset and re-quantised to uint8, because the reference
 *                         corrects frequency only by retuning the hardware).
 *   - ref_run_backend()   feeds already-demapped transmission frames to
 *                         dab_process_frame (dab.c:35-99) and collects the ETI.
 *
 * Every reference function (viterbi, encode, *_depuncture, fic_decode, ...) is
 * also exported by the shared object as-is for unit-level comparisons.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dab.h"
#include "fic.h"
#include "input_sdr.h"
#include "misc.h"

struct ref_call_trace {
  int32_t ok;                /* return value of sdr_demod */
  int32_t coarse_timeshift;  /* sdr state after the call */
  int32_t fine_timeshift;
  int32_t coarse_freq_shift;
  double fine_freq_shift;
  uint32_t frequency;        /* virtual tuner after feedback */
  int32_t locked;            /* dab->locked after dab_process_frame */
  int32_t eti_frames;        /* cumulative ETI frames emitted */
};

/* eti_callback has no context argument (dab.h:88), so the sink is file-static. */
static uint8_t *g_eti_out;
static long g_eti_cap, g_eti_n;

static void collect_eti(uint8_t *eti) {
  if (g_eti_out && (g_eti_n + 1) * 6144 <= g_eti_cap)
    memcpy(g_eti_out + g_eti_n * 6144, eti, 6144);
  g_eti_n++;
}

/* Tuner feedback, transcribed in behaviour from dab2eti.c:75-103.  Returns the new
 * frequency; `frequency` is uint32_t in sdr_state_t so the arithmetic is done on
 * that type exactly as the reference does. */
static void tuner_feedback(struct sdr_state_t *sdr) {
  if (abs(sdr->coarse_freq_shift) > 1) {
    if (sdr->coarse_freq_shift < 0)
      sdr->frequency = sdr->frequency - 1000;
    else
      sdr->frequency = sdr->frequency + 1000;
  }
  if (abs(sdr->coarse_freq_shift) == 1) {
    if (sdr->coarse_freq_shift < 0)
      sdr->frequency = sdr->frequency - rand() % 1000;
    else
      sdr->frequency = sdr->frequency + rand() % 1000;
  }
  if (abs(sdr->coarse_freq_shift) < 1 && (abs(sdr->fine_freq_shift) > 50)) {
    sdr->frequency = sdr->frequency + (sdr->fine_freq_shift / 3);
  }
}

/* Run one capture through the reference receiver.
 *   iq, nbytes    interleaved uint8 I,Q at 2.048 Msps
 *   chunk         bytes per simulated rtlsdr callback (<= 262144)
 *   f0            nominal tuner frequency in Hz (the capture is centred on it)
 *   rand_seed     srand() value for the |cfs|==1 dither (dab2eti.c:88-96)
 *   eti_out/cap   receives 6144-byte ETI frames (may be NULL to only count)
 *   trace/cap     one record per callback (may be NULL)
 *   tf_out/cap    if non-NULL, the demapped fic(9216)+msc(221184) bytes of every
 *                 ok frame are appended (230400 bytes each) up to tf_cap frames
 * Returns the number of ETI frames produced; *n_calls and *n_tfs are set.
 */
long ref_run_iq(const uint8_t *iq, long nbytes, int chunk, uint32_t f0,
                unsigned rand_seed, uint8_t *eti_out, long eti_cap,
                struct ref_call_trace *trace, long trace_cap, long *n_calls,
                uint8_t *tf_out, long tf_cap, long *n_tfs) {
  struct dab_state_t *dab = NULL;
  struct sdr_state_t *sdr = calloc(1, sizeof *sdr);
  long calls = 0, tfs = 0;

  g_eti_out = eti_out;
  g_eti_cap = eti_cap;
  g_eti_n = 0;
  srand(rand_seed);

  init_dab_state(&dab, sdr, collect_eti);
  dab->device_type = DAB_DEVICE_RTLSDR;
  sdr->frequency = f0;
  sdr_init(sdr);

  if (chunk <= 0 || chunk > DEFAULT_BUF_LENGTH) chunk = DEFAULT_BUF_LENGTH;
  for (long pos = 0; pos + chunk <= nbytes; pos += chunk) {
    /* rtlsdr_callback: memcpy + length (dab2eti.c:125-126) through the virtual tuner */
    double df = (double)sdr->frequency - (double)f0;
    if (df == 0.0) {
      memcpy(sdr->input_buffer, iq + pos, chunk);
    } else {
      for (long b = 0; b < chunk; b += 2) {
        double n = (double)((pos + b) / 2);
        double ph = -2.0 * M_PI * df * n / 2048000.0;
        double c = cos(ph), s = sin(ph);
        double xr = (double)iq[pos + b] - 127.0, xi = (double)iq[pos + b + 1] - 127.0;
        double yr = xr * c - xi * s, yi = xr * s + xi * c;
        double qr = floor(yr + 0.5) + 127.0, qi = floor(yi + 0.5) + 127.0;
        sdr->input_buffer[b] = (uint8_t)(qr < 0 ? 0 : qr > 255 ? 255 : qr);
        sdr->input_buffer[b + 1] = (uint8_t)(qi < 0 ? 0 : qi > 255 ? 255 : qi);
      }
    }
    sdr->input_buffer_len = chunk;

    /* demod_thread_fn body (dab2eti.c:68-103) */
    int ok = sdr_demod(&dab->tfs[dab->tfidx], sdr);
    if (ok) {
      if (tf_out && tfs < tf_cap) {
        uint8_t *dst = tf_out + tfs * 230400L;
        memcpy(dst, dab->tfs[dab->tfidx].fic_symbols_demapped, 9216);
        memcpy(dst + 9216, dab->tfs[dab->tfidx].msc_symbols_demapped, 221184);
      }
      tfs++;
      dab_process_frame(dab);
    }
    tuner_feedback(sdr);

    if (trace && calls < trace_cap) {
      struct ref_call_trace *t = &trace[calls];
      t->ok = ok;
      t->coarse_timeshift = sdr->coarse_timeshift;
      t->fine_timeshift = sdr->fine_timeshift;
      t->coarse_freq_shift = sdr->coarse_freq_shift;
      t->fine_freq_shift = sdr->fine_freq_shift;
      t->frequency = sdr->frequency;
      t->locked = dab->locked;
      t->eti_frames = (int32_t)g_eti_n;
    }
    calls++;
  }
  if (n_calls) *n_calls = calls;
  if (n_tfs) *n_tfs = tfs;

  /* the reference never frees these; the harness does so it can be called repeatedly */
  free(sdr->fifo.elems);
  fftw_free(sdr->dab_frame);
  fftw_free(sdr->prs_ifft);
  fftw_free(sdr->prs_conj_ifft);
  fftw_free(sdr->prs_syms);
  fftw_free(sdr->symbols_d);
  free(sdr);
  free(dab);
  g_eti_out = NULL;
  return g_eti_n;
}

/* Streaming form of ref_run_iq for per-step timing (bench.py --impl reference): the receiver
 * state lives across calls, so a warmed-up (locked, window full) stream can be fed one bounded
 * sample per step.  The reference is not re-entrant (statics in misc.c / viterbi.c): one stream
 * per process. */
struct ref_stream {
  struct dab_state_t *dab;
  struct sdr_state_t *sdr;
  uint32_t f0;
  long pos;  /* bytes fed so far (phase continuity of the virtual tuner) */
};

void *ref_stream_open(uint32_t f0, unsigned rand_seed) {
  struct ref_stream *h = calloc(1, sizeof *h);
  h->sdr = calloc(1, sizeof *h->sdr);
  h->f0 = f0;
  srand(rand_seed);
  init_dab_state(&h->dab, h->sdr, collect_eti);
  h->dab->device_type = DAB_DEVICE_RTLSDR;
  h->sdr->frequency = f0;
  sdr_init(h->sdr);
  return h;
}

/* feed nbytes (a multiple of `chunk`) as rtlsdr callbacks; returns the ETI frames produced by
 * this call, the first eti_cap/6144 of which are stored */
long ref_stream_feed(void *hv, const uint8_t *iq, long nbytes, int chunk, uint8_t *eti_out, long eti_cap) {
  struct ref_stream *h = hv;
  struct sdr_state_t *sdr = h->sdr;
  struct dab_state_t *dab = h->dab;
  g_eti_out = eti_out;
  g_eti_cap = eti_cap;
  g_eti_n = 0;
  if (chunk <= 0 || chunk > DEFAULT_BUF_LENGTH) chunk = DEFAULT_BUF_LENGTH;
  for (long pos = 0; pos + chunk <= nbytes; pos += chunk, h->pos += chunk) {
    double df = (double)sdr->frequency - (double)h->f0;
    if (df == 0.0) {
      memcpy(sdr->input_buffer, iq + pos, chunk);
    } else {
      for (long b = 0; b < chunk; b += 2) {
        double n = (double)((h->pos + b) / 2);
        double ph = -2.0 * M_PI * df * n / 2048000.0;
        double c = cos(ph), s = sin(ph);
        double xr = (double)iq[pos + b] - 127.0, xi = (double)iq[pos + b + 1] - 127.0;
        double yr = xr * c - xi * s, yi = xr * s + xi * c;
        double qr = floor(yr + 0.5) + 127.0, qi = floor(yi + 0.5) + 127.0;
        sdr->input_buffer[b] = (uint8_t)(qr < 0 ? 0 : qr > 255 ? 255 : qr);
        sdr->input_buffer[b + 1] = (uint8_t)(qi < 0 ? 0 : qi > 255 ? 255 : qi);
      }
    }
    sdr->input_buffer_len = chunk;
    if (sdr_demod(&dab->tfs[dab->tfidx], sdr)) dab_process_frame(dab);
    tuner_feedback(sdr);
  }
  g_eti_out = NULL;
  return g_eti_n;
}

int ref_stream_locked(void *hv) { return ((struct ref_stream *)hv)->dab->locked; }

void ref_stream_close(void *hv) {
  struct ref_stream *h = hv;
  struct sdr_state_t *sdr = h->sdr;
  free(sdr->fifo.elems);
  fftw_free(sdr->dab_frame);
  fftw_free(sdr->prs_ifft);
  fftw_free(sdr->prs_conj_ifft);
  fftw_free(sdr->prs_syms);
  fftw_free(sdr->symbols_d);
  free(sdr);
  free(h->dab);
  free(h);
}

/* Feed n_tf already-demapped transmission frames (fic 9216 bytes + msc 221184
 * bytes each, values 0/1) to dab_process_frame and collect the ETI frames.
 * Optionally returns the decoded FIBs (12*32 per TF) and CRC flags (12 per TF). */
long ref_run_backend(const uint8_t *tfs, long n_tf, uint8_t *eti_out, long eti_cap,
                     uint8_t *fibs_out, uint8_t *crc_out) {
  struct dab_state_t *dab = NULL;
  g_eti_out = eti_out;
  g_eti_cap = eti_cap;
  g_eti_n = 0;
  init_dab_state(&dab, NULL, collect_eti);
  dab->device_type = DAB_DEVICE_RTLSDR;
  for (long t = 0; t < n_tf; t++) {
    struct demapped_transmission_frame_t *tf = &dab->tfs[dab->tfidx];
    const uint8_t *src = tfs + t * 230400L;
    tf->has_fic = 1;
    memcpy(tf->fic_symbols_demapped, src, 9216);
    memcpy(tf->msc_symbols_demapped, src + 9216, 221184);
    dab_process_frame(dab);
    /* dab_process_frame may have advanced tfidx; the frame just decoded is `tf` */
    if (fibs_out) memcpy(fibs_out + t * 384, tf->fibs.FIB, 384);
    if (crc_out) memcpy(crc_out + t * 12, tf->fibs.FIB_CRC_OK, 12);
  }
  free(dab);
  g_eti_out = NULL;
  return g_eti_n;
}

/* One sdr_demod-shaped call on a caller-assembled frame, for unit-level parity of
 * the synchronisers and the OFDM demodulator: `frame` is 393216 bytes of uint8 IQ
 * that the reference would have in sdr->buffer after sdr_read_fifo.  Outputs the
 * raw spectra (76*2048 complex doubles), the DQPSK products (76*2048; row 0
 * unused) and the demapped bytes. */
int ref_demod_frame(const uint8_t *frame, int force_timesync, int32_t *coarse_timeshift,
                    int32_t *fine_timeshift, int32_t *coarse_freq_shift,
                    double *fine_freq_shift, double *symbols, double *symbols_d,
                    uint8_t *demapped) {
  struct sdr_state_t *sdr = calloc(1, sizeof *sdr);
  struct demapped_transmission_frame_t *tf = calloc(1, sizeof *tf);
  sdr_init(sdr);
  /* pre-load the FIFO so that exactly this frame is read with shift 0 and the
   * start-up discard (input_sdr.c:51-55) has already happened */
  sdr->startup_delay = GAIN_SETTLE_TIME + 1;
  sdr->force_timesync = (uint8_t)force_timesync;
  for (long i = 0; i < 393216; i++) cbWrite(&sdr->fifo, (uint8_t *)&frame[i]);
  /* need count >= 589824 to pass the gate: pad with a copy of the head (never read) */
  for (long i = 0; i < 196608; i++) cbWrite(&sdr->fifo, (uint8_t *)&frame[i]);
  sdr->input_buffer_len = 0;
  int ok = sdr_demod(tf, sdr);
  *coarse_timeshift = sdr->coarse_timeshift;
  *fine_timeshift = sdr->fine_timeshift;
  *coarse_freq_shift = sdr->coarse_freq_shift;
  *fine_freq_shift = sdr->fine_freq_shift;
  if (symbols) memcpy(symbols, sdr->symbols, sizeof(double) * 2 * 76 * 2048);
  if (symbols_d) memcpy(symbols_d, sdr->symbols_d, sizeof(double) * 2 * 76 * 2048);
  if (demapped) {
    memcpy(demapped, tf->fic_symbols_demapped, 9216);
    memcpy(demapped + 9216, tf->msc_symbols_demapped, 221184);
  }
  free(sdr->fifo.elems);
  fftw_free(sdr->dab_frame);
  fftw_free(sdr->prs_ifft);
  fftw_free(sdr->prs_conj_ifft);
  fftw_free(sdr->prs_syms);
  fftw_free(sdr->symbols_d);
  free(sdr);
  free(tf);
  return ok;
}

/* ---- Wavefinder producer: the reference's input_wf.c, unmodified, fed from a file of USB packets ----
 * wf_read_frame() (input_wf.c:65-115) reads 524-byte packets from wf->fd and returns a frame at every
 * NULL-symbol packet; do_wf_decode (dab2eti.c:251-272) discards the first frame and hands every
 * further one to dab_process_frame.  The device's timing/AFC loop (wf_sync.c: wf_prs_assemble talks
 * to the hardware through ioctls) cannot run without the device: it is stubbed and sync_locked is
 * set, i.e. the stream is taken as already synchronised. */
#include <unistd.h>
#include "input_wf.h"
int wf_prs_assemble(struct wavefinder_t *wf, unsigned char *buf) { (void)wf; (void)buf; return 0; }
int wfsyncinit(void) { return 0; }

long ref_run_wf(const uint8_t *packets, long n_packets, uint8_t *eti_out, long eti_cap) {
  struct dab_state_t *dab = NULL;
  struct wavefinder_t wf;
  char path[] = "/tmp/ref_wf_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0) return -1;
  unlink(path);
  if (write(fd, packets, n_packets * 524) != n_packets * 524) return -1;
  lseek(fd, 0, SEEK_SET);
  memset(&wf, 0, sizeof wf);
  wf.fd = fd;
  g_eti_out = eti_out;
  g_eti_cap = eti_cap;
  g_eti_n = 0;
  init_dab_state(&dab, &wf, collect_eti);
  dab->device_type = DAB_DEVICE_WAVEFINDER;
  wf_init(&wf);
  wf.sync_locked = 1;
  int stderr_copy = dup(2);  /* wf_read_frame reports the end of the file as a read error */
  if (wf_read_frame(&wf, &dab->tfs[0]) == 0) {
    while (wf_read_frame(&wf, &dab->tfs[dab->tfidx]) == 0) dab_process_frame(dab);
  }
  (void)stderr_copy;
  close(fd);
  free(dab);
  g_eti_out = NULL;
  return g_eti_n;
}

/* Accessors for file-static / extern tables the tests want to pin. */
extern const uint16_t rev_freq_deint_tab[1536];
extern fftw_complex prs_static[1536];
extern int Syms[];
int ref_sizeof_dab_state(void) { return (int)sizeof(struct dab_state_t); }
int ref_sizeof_sdr_state(void) { return (int)sizeof(struct sdr_state_t); }
int ref_sizeof_tf(void) { return (int)sizeof(struct demapped_transmission_frame_t); }

#include <stddef.h>
void ref_abi_offsets(int32_t *out) {
  int i = 0;
  out[i++] = (int)offsetof(struct dab_state_t, tfs);
  out[i++] = (int)offsetof(struct dab_state_t, tf_info);
  out[i++] = (int)offsetof(struct dab_state_t, ens_info);
  out[i++] = (int)offsetof(struct dab_state_t, cifs_msc);
  out[i++] = (int)offsetof(struct dab_state_t, ncifs);
  out[i++] = (int)offsetof(struct dab_state_t, eti_callback);
  out[i++] = (int)offsetof(struct sdr_state_t, input_buffer_len);
  out[i++] = (int)offsetof(struct sdr_state_t, buffer);
  out[i++] = (int)offsetof(struct sdr_state_t, fine_freq_shift);
  out[i++] = (int)offsetof(struct sdr_state_t, fifo);
  out[i++] = (int)offsetof(struct sdr_state_t, symbols);
  out[i++] = (int)offsetof(struct sdr_state_t, startup_delay);
  out[i++] = (int)offsetof(struct sdr_state_t, p_e_after_vitdec);
  out[i++] = (int)offsetof(struct demapped_transmission_frame_t, fibs);
  out[i++] = (int)offsetof(struct demapped_transmission_frame_t, msc_symbols_demapped);
}
