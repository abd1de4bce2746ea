/* dab_oracle.h -- CPU restatement of the dabtools receive hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dabtools_b200/ or include/ may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Every function restates (not copies) the algorithm of the reference function
 * named in its comment, file:line relative to /root/reference.  Parity pinning:
 * the reference ships no tests or golden vectors, so this port is
 * pinned against the reference itself, compiled unmodified into oracle/_ref
 * (tests/test_oracle_vs_ref.py, run in the build container) and against the
 * fixtures that script family commits under tests/golden/.
 */
#ifndef DAB_ORACLE_H
#define DAB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- channel decoding ---------------------------------------------------------- */
void orc_gen_metrics(int mettab[2][256], int amp, double noise, double bias, int scale);
void orc_encode(uint8_t *symbols, const uint8_t *data, unsigned nbytes);
int orc_viterbi(const uint8_t *symbols, uint8_t *data, unsigned nbits);

void orc_fic_depuncture(uint8_t *out3096, const uint8_t *in2304);
int orc_uep_depuncture(uint8_t *out, const uint8_t *in, int uep_index);
int orc_eep_depuncture(uint8_t *out, const uint8_t *in, int protlev, int size_cu, int bitrate);

void orc_descramble(uint8_t *buf, int nbytes);
uint16_t orc_crc16(const uint8_t *p, int n, uint16_t init);
int orc_check_fib_crc(const uint8_t *fib32);
void orc_time_deinterleave(uint8_t *dst, const uint8_t *const cifs[16]);
int orc_fic_decode(const uint8_t *fic_bits9216, uint8_t *fibs384, uint8_t *crc_ok12);

/* ---- FIG parsing / ensemble bookkeeping ---------------------------------------- */
struct orc_subch {
  int id; /* -1 = inactive */
  int eepprot, slForm, uep_index, eep_option, start_cu, size, bitrate, eep_protlev, protlev, ASCTy;
};
struct orc_tf_info {
  uint16_t EId;
  uint8_t CIFCount_hi, CIFCount_lo;
  struct orc_subch subchans[64];
};
struct orc_ens_info {
  uint16_t EId;
  uint8_t CIFCount_hi, CIFCount_lo;
  struct orc_subch subchans[64];
};
void orc_fib_decode(struct orc_tf_info *info, const uint8_t *fibs384, const uint8_t *crc_ok12, int nfibs);
void orc_merge_info(struct orc_ens_info *ei, const struct orc_tf_info *info);
int orc_init_eti(uint8_t *eti, const struct orc_ens_info *info);

/* ---- receiver back-end (demapped TF -> ETI) ------------------------------------- */
struct orc_rx; /* opaque */
typedef void (*orc_eti_cb)(void *ctx, const uint8_t *eti6144);
struct orc_rx *orc_rx_new(orc_eti_cb cb, void *ctx);
void orc_rx_free(struct orc_rx *rx);
/* buffers of the TF slot the next frame must be written to (fic 9216, msc 221184) */
uint8_t *orc_rx_fic_slot(struct orc_rx *rx);
uint8_t *orc_rx_msc_slot(struct orc_rx *rx);
void orc_rx_process_frame(struct orc_rx *rx);
int orc_rx_locked(const struct orc_rx *rx);
const uint8_t *orc_rx_last_fibs(const struct orc_rx *rx);   /* 384 bytes */
const uint8_t *orc_rx_last_crc(const struct orc_rx *rx);    /* 12 bytes  */
long orc_run_backend(const uint8_t *tfs, long n_tf, uint8_t *eti_out, long eti_cap,
                     uint8_t *fibs_out, uint8_t *crc_out);

/* ---- RTL-SDR front-end ---------------------------------------------------------- */
uint32_t orc_coarse_time_sync(const int8_t *real, int force);
int32_t orc_fine_time_sync(const double *frame /* 196608 x (re,im) */);
int32_t orc_coarse_freq_sync(const double *shifted_spectrum /* 2048 x (re,im), fftshifted */);
double orc_fine_freq_corr(const double *frame);

struct orc_call_trace {
  int32_t ok, coarse_timeshift, fine_timeshift, coarse_freq_shift;
  double fine_freq_shift;
  uint32_t frequency;
  int32_t locked, eti_frames;
};
int orc_demod_frame(const uint8_t *frame393216, int force_timesync, int32_t *cts, int32_t *fts,
                    int32_t *cfs, double *ffs, double *symbols, double *symbols_d,
                    uint8_t *demapped230400);
long orc_run_iq(const uint8_t *iq, long nbytes, int chunk, uint32_t f0, unsigned rand_seed,
                uint8_t *eti_out, long eti_cap, struct orc_call_trace *trace, long trace_cap,
                long *n_calls, uint8_t *tf_out, long tf_cap, long *n_tfs);

/* glibc-compatible rand() (TYPE_3 additive feedback), so the |cfs|==1 dither of the
 * tuner feedback (dab2eti.c:88-96) is reproducible per stream */
struct orc_rand { int32_t r[34]; int idx; };
void orc_srand(struct orc_rand *s, unsigned seed);
int orc_rand(struct orc_rand *s);

/* ---- table accessors (from include/dabgpu_tables.h) for the tests --------------- */
void orc_tab_freq_deint(uint16_t rev[1536]);
void orc_tab_prs(uint8_t q[1536]);
uint32_t orc_tab_puncture_mask(int pi);
int orc_tab_shape(int kind /*0 fic,1 uep,2 eep*/, int a, int b, int32_t *out /* 2+1+20 ints */);

#ifdef __cplusplus
}
#endif
#endif
