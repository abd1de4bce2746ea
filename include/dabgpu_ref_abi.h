/* dabgpu_ref_abi.h -- the reference's own C interface for the receive hot path, as exported
 * by libdabgpu.
 *
 * These are the functions and the in-memory struct layouts that dab2eti.c (and the
 * reference's internal callers) use for this path; libdabgpu exports them with identical
 * names, argument meaning and error behaviour, so an unmodified dab2eti.o can be linked
 * against libdabgpu.so instead of input_sdr.o sdr_sync.o sdr_fifo.o dab.o fic.o misc.o
 * depuncture.o viterbi.o dab_tables.o.  Every compute step behind them runs as CUDA kernels
 * on the current device (batch of one); control logic (lock state machine, FIG parsing,
 * window bookkeeping) stays on the host as in the reference.
 *
 * Layout contract: field order, types and array sizes below reproduce
 *   src/dab.h:21-89, src/input_sdr.h:12-41, src/sdr_fifo.h:27-33
 * byte for byte (tests/test_abi_layout.py compares sizeof/offsetof with the compiled
 * reference).  Each prototype cites the reference declaration it replaces.
 */
#ifndef DABGPU_REF_ABI_H
#define DABGPU_REF_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* <fftw3.h> is not required: the only thing the ABI needs from it is this typedef. */
#ifndef DABGPU_HAVE_FFTW_COMPLEX
#define DABGPU_HAVE_FFTW_COMPLEX
typedef double fftw_complex[2];
#endif

/* ---- src/dab.h ------------------------------------------------------------------------- */
enum device_type_t { DAB_DEVICE_WAVEFINDER, DAB_DEVICE_RTLSDR };

struct tf_fibs_t {                 /* dab.h:21-25 */
  uint8_t ok_count;
  uint8_t FIB[12][32];
  uint8_t FIB_CRC_OK[12];
};

struct demapped_transmission_frame_t { /* dab.h:27-33: hand-off front-end -> back-end */
  uint8_t has_fic;
  uint8_t fic_symbols_demapped[3][3072];
  struct tf_fibs_t fibs;
  uint8_t msc_filter[72];
  uint8_t msc_symbols_demapped[72][3072];
};

struct subchannel_info_t {         /* dab.h:35-47 */
  int id;
  int eepprot;
  int slForm;
  int uep_index;
  int eep_option;
  int start_cu;
  int size;
  int bitrate;
  int eep_protlev;
  int protlev;
  int ASCTy;
};

struct tf_info_t {                 /* dab.h:50-61 */
  uint16_t EId;
  uint8_t CIFCount_hi;
  uint8_t CIFCount_lo;
  struct subchannel_info_t subchans[64];
};

struct ens_info_t {                /* dab.h:63-68 */
  uint16_t EId;
  uint8_t CIFCount_hi;
  uint8_t CIFCount_lo;
  struct subchannel_info_t subchans[64];
};

struct dab_state_t {               /* dab.h:70-89 */
  enum device_type_t device_type;
  void *device_state;
  struct demapped_transmission_frame_t tfs[5];
  struct tf_info_t tf_info;
  struct ens_info_t ens_info;
  void *v;
  unsigned char *cifs_msc[16];
  unsigned char *cifs_fibs[16];
  int ncifs;
  int tfidx;
  int locked;
  int ens_info_shown;
  int okcount;
  void (*eti_callback)(uint8_t *eti);
};

/* ---- src/sdr_fifo.h ---------------------------------------------------------------------- */
typedef struct {                   /* sdr_fifo.h:27-33 */
  uint32_t size;
  uint32_t start;
  uint32_t count;
  uint8_t *elems;
} CircularBuffer;

/* ---- src/input_sdr.h --------------------------------------------------------------------- */
#define DEFAULT_BUF_LENGTH (16 * 16384)
#define GAIN_SETTLE_TIME 0

struct sdr_state_t {               /* input_sdr.h:12-41 */
  uint32_t frequency;
  uint8_t input_buffer[DEFAULT_BUF_LENGTH];
  int input_buffer_len;
  uint8_t buffer[196608 * 2];
  int32_t coarse_timeshift;
  int32_t fine_timeshift;
  int32_t coarse_freq_shift;
  double fine_freq_shift;
  CircularBuffer fifo;
  int8_t real[196608];
  int8_t imag[196608];
  float filt[196608 - 2662];
  fftw_complex *dab_frame;
  fftw_complex *prs_ifft;
  fftw_complex *prs_conj_ifft;
  fftw_complex *prs_syms;
  fftw_complex symbols[76][2048];
  fftw_complex *symbols_d;
  int32_t startup_delay;
  uint8_t force_timesync;
  double p_e_prior_dep;
  double p_e_prior_vitdec;
  double p_e_after_vitdec;
};

/* ---- front-end: src/input_sdr.h:43-44, src/sdr_sync.h:28-31, src/sdr_fifo.h:36-43 --------- */
void sdr_init(struct sdr_state_t *sdr);
int sdr_demod(struct demapped_transmission_frame_t *tf, struct sdr_state_t *sdr);
uint32_t dab_coarse_time_sync(int8_t *real, float *filt, uint8_t force_timesync);
int32_t dab_fine_time_sync(fftw_complex *frame);
int32_t dab_coarse_freq_sync_2(fftw_complex *symbols);
double dab_fine_freq_corr(fftw_complex *dab_frame, int32_t fine_timeshift);
void cbInit(CircularBuffer *cb, uint32_t size);
void cbFree(CircularBuffer *cb);
int cbIsFull(CircularBuffer *cb);
int cbIsEmpty(CircularBuffer *cb);
void cbWrite(CircularBuffer *cb, uint8_t *elem);
void cbRead(CircularBuffer *cb, uint8_t *elem);
int32_t sdr_read_fifo(CircularBuffer *fifo, uint32_t bytes, int32_t shift, uint8_t *buffer);

/* ---- channel decoding: src/depuncture.h:7-9, src/viterbi.h:6-8, src/misc.h:6-10 ------------- */
void fic_depuncture(uint8_t *obuf, uint8_t *inbuf);
void uep_depuncture(uint8_t *obuf, uint8_t *inbuf, struct subchannel_info_t *s, int *len);
void eep_depuncture(uint8_t *obuf, uint8_t *inbuf, struct subchannel_info_t *s, int *len);
int init_viterbi(void);
/* viterbi_spiral.h:22: what dab.c:27-30 calls instead of init_viterbi() when the reference is built
 * with -DENABLE_SPIRAL_VITERBI.  Returns an opaque non-null handle; the decoder behind viterbi() is
 * the same (plain viterbi.c tie-breaking, the alphabet of either to_viterbi() variant is accepted) */
void *create_viterbi(int len);
/* viterbi.h:8 declares void, viterbi.c:352 defines int(...unsigned); int is ABI-compatible.
 * Symbols: to_viterbi()'s alphabet (127 / 128 / 129, or 0 / 128 / 255 of the Spiral build) goes to the
 * hard-decision kernel; any other value makes the call a soft-decision decode with the reference's
 * metric table gen_met(1, 1.0, 0, 4), exact for symbols in 121..135 and saturating beyond (the
 * reference's table overflows to INT_MIN there, viterbi.c:126-191). */
int viterbi(void *p, unsigned char *symbols, unsigned char *data, unsigned int framebits);
void dab_descramble_bytes(uint8_t *buf, int32_t nbytes);
int check_fib_crc(uint8_t *data);
void time_deinterleave(uint8_t *dst, uint8_t *cifs[]);  /* misc.c:29 (not in a header) */

/* ---- FIC / ensemble / ETI: src/fic.h:5-7, src/misc.h:6-8, src/dab.h:91-92 -------------------- */
void fic_decode(struct dab_state_t *dab, struct demapped_transmission_frame_t *tf);
void fib_decode(struct tf_info_t *info, struct tf_fibs_t *fibs, int nfibs);
void dump_tf_info(struct tf_info_t *info);
void merge_info(struct ens_info_t *ei, struct tf_info_t *info);
void dump_ens_info(struct ens_info_t *info);
int init_eti(uint8_t *eti, struct ens_info_t *info);    /* misc.c:153 (not in a header) */
void create_eti(struct dab_state_t *dab);
void init_dab_state(struct dab_state_t **dab, void *device_state, void (*eti_callback)(uint8_t *eti));
void dab_process_frame(struct dab_state_t *dab);

#ifdef __cplusplus
}
#endif
#endif /* DABGPU_REF_ABI_H */
