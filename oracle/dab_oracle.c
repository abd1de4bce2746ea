/* dab_oracle.c -- CPU restatement of the dabtools receive hot path.
 * TEST INFRASTRUCTURE ONLY -- see dab_oracle.h for the rules and the pinning story.
 *
 * All file:line citations are relative to /root/reference.
 */
#include "dab_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dabgpu_tables.h"
#include "ref_shim/fftw3.h" /* our own double-precision DFT (ref_shim/fftw_shim.c) */

/* ==================================================================================
 * 1. Viterbi metrics, encoder, decoder             (src/viterbi.c)
 * ================================================================================== */

static double phi(double x) { return 0.5 + 0.5 * erf(x / M_SQRT2); }
static double lg2(double x) { return log(x) * M_LOG2E; } /* viterbi.c:120 (sic: ln(x)*log2(e)) */

static int round_to_int(double v) {
  double f = floor(v + 0.5);
  /* (int) of NaN / +-inf / out-of-range is INT_MIN on x86-64 (cvttsd2si), which is
   * what the reference's table ends up holding for the far-out symbol values */
  if (!(f > -2147483648.0 && f < 2147483648.0)) return INT_MIN;
  return (int)f;
}

/* viterbi.c:126-191 gen_met(): log-likelihood metrics for an 8-bit soft symbol
 * centred on 128 under BPSK+AWGN; rx value 0 and 255 absorb the tails. */
void orc_gen_metrics(int mettab[2][256], int amp, double noise, double bias, int scale) {
  for (int s = 0; s < 256; s++) {
    double p0, p1;
    double lo = (s - 128 - 0.5) / amp, hi = (s - 128 + 0.5) / amp;
    if (s == 0) {
      p1 = phi((hi - 1) / noise);
      p0 = phi((hi + 1) / noise);
    } else if (s == 255) {
      p1 = 1 - phi((lo - 1) / noise);
      p0 = 1 - phi((lo + 1) / noise);
    } else {
      p1 = phi((hi - 1) / noise) - phi((lo - 1) / noise);
      p0 = phi((hi + 1) / noise) - phi((lo + 1) / noise);
    }
    double m0 = lg2(2 * p0 / (p1 + p0)) - bias;
    double m1 = lg2(2 * p1 / (p1 + p0)) - bias;
    mettab[0][s] = round_to_int(m0 * scale);
    mettab[1][s] = round_to_int(m1 * scale);
  }
}

static int parity8(unsigned x) {
  x ^= x >> 4;
  x ^= x >> 2;
  x ^= x >> 1;
  return (int)(x & 1u);
}

/* viterbi.c:322-347 encode(): MSB-first data bits into a left-shifting register,
 * 4 output symbols (0/1) per bit, then 6 zero tail bits. */
void orc_encode(uint8_t *symbols, const uint8_t *data, unsigned nbytes) {
  unsigned reg = 0;
  unsigned total = nbytes * 8 + 6;
  for (unsigned t = 0; t < total; t++) {
    unsigned bit = t < nbytes * 8 ? (data[t >> 3] >> (7 - (t & 7))) & 1u : 0u;
    reg = (reg << 1) | bit;
    for (int j = 0; j < 4; j++) *symbols++ = (uint8_t)parity8(reg & DABGPU_POLYS[j]);
  }
}

static int g_mettab[2][256];
static int g_branch_sym[128]; /* expected 4-bit symbol for 7-bit register value */
static int g_vit_ready;

static void viterbi_setup(void) {
  if (g_vit_ready) return;
  orc_gen_metrics(g_mettab, 1, 1.0, 0.0, 4); /* viterbi.c:455-462 init_viterbi() */
  for (int r = 0; r < 128; r++) {
    int s = 0;
    for (int j = 0; j < 4; j++) s = (s << 1) | parity8((unsigned)r & DABGPU_POLYS[j]);
    g_branch_sym[r] = s; /* viterbi.c:370-376 Syms[] */
  }
  g_vit_ready = 1;
}

/* viterbi.c:352-452 viterbi(): 64-state maximum-metric decoder, start state 0
 * (others at -999999), one 64-bit decision word per step, full traceback from
 * state 0.  Predecessors of state i are i>>1 (decision 0) and (i>>1)+32
 * (decision 1); the second wins only on a strictly larger metric. */
int orc_viterbi(const uint8_t *symbols, uint8_t *data, unsigned nbits) {
  viterbi_setup();
  unsigned steps = nbits + 6;
  uint64_t *dec = (uint64_t *)calloc(steps, sizeof(uint64_t));
  long cur[64], nxt[64];
  for (int i = 0; i < 64; i++) cur[i] = -999999;
  cur[0] = 0;

  for (unsigned t = 0; t < steps; t++) {
    int bm[16];
    for (int c = 0; c < 16; c++) {
      int m = 0;
      for (int j = 0; j < 4; j++) m += g_mettab[(c >> (3 - j)) & 1][symbols[4 * t + j]];
      bm[c] = m;
    }
    uint64_t d = 0;
    for (int i = 0; i < 64; i++) {
      /* register value on the branch pred -> i is (pred<<1 | (i&1)) = i or i+64 */
      long via0 = cur[i >> 1] + bm[g_branch_sym[i]];
      long via1 = cur[(i >> 1) + 32] + bm[g_branch_sym[i + 64]];
      if (via1 > via0) {
        nxt[i] = via1;
        d |= (uint64_t)1 << i;
      } else {
        nxt[i] = via0;
      }
    }
    dec[t] = d;
    memcpy(cur, nxt, sizeof cur);
  }

  if (data) {
    memset(data, 0, (nbits + 7) / 8);
    unsigned state = 0;
    for (int i = (int)nbits - 1; i >= 0; i--) {
      /* decision taken when entering `state` at step i+6 is the bit that left the
       * register there, i.e. information bit i (viterbi.c:443-450) */
      unsigned bit = (unsigned)((dec[i + 6] >> state) & 1u);
      if (bit) data[i >> 3] |= (uint8_t)(0x80 >> (i & 7));
      state = (state >> 1) | (bit << 5);
    }
  }
  free(dec);
  return 0;
}

/* ==================================================================================
 * 2. Depuncturing                                   (src/depuncture.c)
 * ================================================================================== */

/* Soft-decision extension (NOT in the reference, whose demapper slices): with
 * orc_set_soft(1) the "bits" handed to the depuncturers are received symbol values around 128 and
 * are passed through to viterbi() as they are -- saturated to 121..135, the range in which the
 * reference's metric table gen_met(1, 1.0, 0, 4) is defined (viterbi.c:126-191) -- instead of being
 * mapped by to_viterbi().  viterbi() itself is the reference's algorithm unchanged. */
static int g_soft;
void orc_set_soft(int on) { g_soft = on; }

/* depuncture.c:36-43 to_viterbi(): hard bit -> 127/129, punctured -> 128 */
static int depuncture_shape(const dabgpu_cw_shape *sh, uint8_t *out, const uint8_t *in) {
  int k = 0, j = 0;
  for (int r = 0; r < sh->n_regions; r++) {
    uint32_t mask = dabgpu_puncture_mask(sh->r[r].pi);
    int nbits = 4 * sh->r[r].steps;
    for (int i = 0; i < nbits; i++) {
      if ((mask >> (i & 31)) & 1u) {
        const int v = in[j++];
        out[k++] = g_soft ? (uint8_t)(v < 121 ? 121 : v > 135 ? 135 : v) : (uint8_t)(127 + 2 * v);
      } else
        out[k++] = 128;
    }
  }
  return k;
}

/* depuncture.c:45-82 */
void orc_fic_depuncture(uint8_t *out, const uint8_t *in) {
  dabgpu_cw_shape sh;
  dabgpu_shape_fic(&sh);
  depuncture_shape(&sh, out, in);
}

/* depuncture.c:84-105 */
int orc_uep_depuncture(uint8_t *out, const uint8_t *in, int uep_index) {
  dabgpu_cw_shape sh;
  if (dabgpu_shape_uep(&sh, uep_index)) return -1;
  return depuncture_shape(&sh, out, in);
}

/* depuncture.c:107-132 (including the 8 kbit/s 2-A special case at :113-114, which
 * only triggers when the *bitrate* field says 8) */
int orc_eep_depuncture(uint8_t *out, const uint8_t *in, int protlev, int size_cu, int bitrate) {
  dabgpu_cw_shape sh;
  if (dabgpu_shape_eep(&sh, protlev, size_cu, bitrate)) return -1;
  return depuncture_shape(&sh, out, in);
}

/* ==================================================================================
 * 3. Energy dispersal, CRC, time de-interleaving    (src/misc.c)
 * ================================================================================== */

/* misc.c:41-58 dab_descramble_bytes(): XOR with the x^9+x^5+1 PRBS, restarted (all
 * ones) on every call, MSB first */
void orc_descramble(uint8_t *buf, int nbytes) {
  unsigned reg = 0x1ff;
  for (int i = 0; i < nbytes; i++) {
    unsigned q = 0;
    for (int b = 0; b < 8; b++) {
      unsigned fb = ((reg >> 8) ^ (reg >> 4)) & 1u;
      reg = ((reg << 1) | fb) & 0x1ffu;
      q = (q << 1) | fb;
    }
    buf[i] ^= (uint8_t)q;
  }
}

/* misc.c:131-143 calc_crc() with crctab_1021: CRC-16-CCITT, MSB first */
uint16_t orc_crc16(const uint8_t *p, int n, uint16_t init) { return dabgpu_crc16(p, n, init); }

/* misc.c:145-150: CRC over all 32 bytes (data + inverted CRC) must leave 0x1d0f */
int orc_check_fib_crc(const uint8_t *fib) { return orc_crc16(fib, 32, 0xffff) == 0x1d0f; }

/* misc.c:29-39: bit i of the output comes from window slot map[i&15], slot 0 oldest */
void orc_time_deinterleave(uint8_t *dst, const uint8_t *const cifs[16]) {
  for (int i = 0; i < DABGPU_CIF_BITS; i++) dst[i] = cifs[DABGPU_TDI_DELAY[i & 15]][i];
}

/* ==================================================================================
 * 4. FIC decode and FIG parsing                     (src/fic.c)
 * ================================================================================== */

/* fic.c:160-208 (has_fic == 1 branch): 4 x { depuncture 2304->3096, viterbi 768,
 * descramble 96 bytes, CRC of the 3 FIBs } */
int orc_fic_decode(const uint8_t *fic_bits, uint8_t *fibs, uint8_t *crc_ok) {
  uint8_t soft[3096];
  int ok = 0;
  for (int g = 0; g < 4; g++) {
    orc_fic_depuncture(soft, fic_bits + g * 2304);
    orc_viterbi(soft, fibs + g * 96, 768);
    orc_descramble(fibs + g * 96, 96);
    for (int j = 0; j < 3; j++) {
      int good = orc_check_fib_crc(fibs + g * 96 + j * 32);
      crc_ok[g * 3 + j] = (uint8_t)good;
      ok += good;
    }
  }
  return ok;
}

/* fic.c:47-130 fib_parse(): FIG 0/0 (EId, CIF count), FIG 0/1 (sub-channel
 * organisation), FIG 0/2 (only ASCTy of stream-mode audio components) */
static void fib_parse(struct orc_tf_info *info, const uint8_t *fib) {
  int i = 0;
  while (fib[i] != 0xff && i < 30) {
    int type = fib[i] >> 5;
    int len = fib[i] & 0x1f;
    i++;
    if (type == 0) {
      int ext = fib[i] & 0x1f;
      int pd = (fib[i] >> 5) & 1;
      if (ext == 0) {
        info->EId = (uint16_t)((fib[i + 1] << 8) | fib[i + 2]);
        info->CIFCount_hi = fib[i + 3] & 0x1f;
        info->CIFCount_lo = fib[i + 4];
      } else if (ext == 1) {
        int j = i + 1;
        while (j < i + len) {
          int id = fib[j] >> 2;
          struct orc_subch *sc = &info->subchans[id];
          sc->id = id;
          sc->start_cu = ((fib[j] & 3) << 8) | fib[j + 1];
          sc->slForm = fib[j + 2] >> 7;
          sc->eepprot = sc->slForm;
          if (!sc->slForm) {
            sc->uep_index = fib[j + 2] & 0x3f;
            sc->size = DABGPU_UEP[sc->uep_index].size_cu;
            sc->bitrate = DABGPU_UEP[sc->uep_index].bitrate;
            sc->protlev = DABGPU_UEP[sc->uep_index].prot_level;
            j += 3;
          } else {
            int option = (fib[j + 2] >> 4) & 7;
            sc->protlev = ((fib[j + 2] >> 2) & 3) | (option << 2);
            sc->size = ((fib[j + 2] & 3) << 8) | fib[j + 3];
            /* the reference indexes eeptable[protlev] unguarded; options >1 are
             * reserved in the standard, so only 0..7 is defined behaviour */
            int lv = sc->protlev & 7;
            sc->bitrate = (sc->size / dabgpu_eep_size_mul(lv)) * (lv < 4 ? 8 : 32);
            j += 4;
          }
        }
      } else if (ext == 2) {
        int j = i + 1;
        while (j < i + len) {
          j += pd ? 4 : 2; /* SId */
          int ncomp = fib[j++] & 0x0f;
          for (int k = 0; k < ncomp; k++) {
            int tmid = fib[j] >> 6;
            if (tmid == 0) {
              int id = fib[j + 1] >> 2;
              info->subchans[id].ASCTy = fib[j] & 0x3f;
            }
            j += 2;
          }
        }
      }
    }
    i += len;
  }
}

/* fic.c:132-147 */
void orc_fib_decode(struct orc_tf_info *info, const uint8_t *fibs, const uint8_t *crc_ok, int nfibs) {
  memset(info, 0, sizeof *info);
  for (int i = 0; i < 64; i++) {
    info->subchans[i].id = -1;
    info->subchans[i].ASCTy = -1;
  }
  for (int i = 0; i < nfibs; i++)
    if (crc_ok[i]) fib_parse(info, fibs + 32 * i);
}

/* misc.c:14-27 */
void orc_merge_info(struct orc_ens_info *ei, const struct orc_tf_info *info) {
  for (int i = 0; i < 64; i++)
    if (info->subchans[i].id >= 0) ei->subchans[i] = info->subchans[i];
  ei->EId = info->EId;
  if (ei->CIFCount_hi == 0xff) { /* latched once, free-running afterwards */
    ei->CIFCount_hi = info->CIFCount_hi;
    ei->CIFCount_lo = info->CIFCount_lo;
  }
}

/* ==================================================================================
 * 5. ETI(NI) frame assembly                         (src/misc.c:153-314)
 * ================================================================================== */

/* misc.c:153-213 init_eti(): SYNC, FC, STC, EOH. Returns the header length. */
int orc_init_eti(uint8_t *eti, const struct orc_ens_info *info) {
  int n = 0, nst = 0, fl = 0;
  int odd = info->CIFCount_lo & 1;
  eti[n++] = 0xff; /* ERR */
  eti[n++] = odd ? 0xf8 : 0x07; /* FSYNC alternates */
  eti[n++] = odd ? 0xc5 : 0x3a;
  eti[n++] = odd ? 0x49 : 0xb6;
  eti[n++] = info->CIFCount_lo; /* FCT */
  for (int j = 0; j < 64; j++)
    if (info->subchans[j].id >= 0) {
      nst++;
      fl += info->subchans[j].bitrate * 3 / 4;
    }
  fl += nst + 1 + 24; /* STC words + EOH + 96 FIC bytes, in 32-bit words */
  eti[n++] = (uint8_t)(0x80 | nst); /* FICF=1 */
  int fp = (info->CIFCount_hi * 250 + info->CIFCount_lo) % 8;
  eti[n++] = (uint8_t)((fp << 5) | (1 << 3) | ((fl >> 8) & 7)); /* MID=1 */
  eti[n++] = (uint8_t)(fl & 0xff);
  for (int j = 0; j < 64; j++) {
    const struct orc_subch *sc = &info->subchans[j];
    if (sc->id < 0) continue;
    int tpl = sc->slForm == 0 ? (0x10 | (sc->protlev - 1)) : (0x20 | sc->protlev);
    int stl = sc->bitrate * 3 / 8;
    eti[n++] = (uint8_t)((sc->id << 2) | ((sc->start_cu >> 8) & 3));
    eti[n++] = (uint8_t)(sc->start_cu & 0xff);
    eti[n++] = (uint8_t)((tpl << 2) | ((stl >> 8) & 3));
    eti[n++] = (uint8_t)(stl & 0xff);
  }
  eti[n++] = 0xff; /* MNSC */
  eti[n++] = 0xff;
  uint16_t hcrc = (uint16_t)~orc_crc16(eti + 4, n - 4, 0xffff);
  eti[n++] = (uint8_t)(hcrc >> 8);
  eti[n++] = (uint8_t)(hcrc & 0xff);
  return n;
}

/* ==================================================================================
 * 6. Receiver back-end state machine                (src/dab.c)
 * ================================================================================== */

struct orc_tf {
  uint8_t fic[9216];
  uint8_t msc[72 * 3072];
  uint8_t fibs[384];
  uint8_t crc_ok[12];
  int ok_count;
  int no_fic; /* Wavefinder frame without all three FIC symbols (tf->has_fic == 0) */
};

struct orc_rx {
  struct orc_tf tfs[5];
  struct orc_tf_info tf_info;
  struct orc_ens_info ens;
  const uint8_t *cif_msc[16];
  const uint8_t *cif_fib[16];
  int ncifs, tfidx, locked, okcount;
  const struct orc_tf *last;
  orc_eti_cb cb;
  void *ctx;
};

/* dab.c:14-33 init_dab_state() */
struct orc_rx *orc_rx_new(orc_eti_cb cb, void *ctx) {
  struct orc_rx *rx = (struct orc_rx *)calloc(1, sizeof *rx);
  for (int i = 0; i < 64; i++) {
    rx->ens.subchans[i].id = -1;
    rx->ens.subchans[i].ASCTy = -1;
  }
  rx->ens.CIFCount_hi = 0xff;
  rx->ens.CIFCount_lo = 0xff;
  rx->cb = cb;
  rx->ctx = ctx;
  return rx;
}
void orc_rx_free(struct orc_rx *rx) { free(rx); }
uint8_t *orc_rx_fic_slot(struct orc_rx *rx) { return rx->tfs[rx->tfidx].fic; }
uint8_t *orc_rx_msc_slot(struct orc_rx *rx) { return rx->tfs[rx->tfidx].msc; }
int orc_rx_locked(const struct orc_rx *rx) { return rx->locked; }
const uint8_t *orc_rx_last_fibs(const struct orc_rx *rx) { return rx->last ? rx->last->fibs : NULL; }
const uint8_t *orc_rx_last_crc(const struct orc_rx *rx) { return rx->last ? rx->last->crc_ok : NULL; }

/* misc.c:218-314 create_eti(): header, FIBs of the oldest CIF, then every active
 * sub-channel in SubChId order: depuncture -> viterbi -> descramble; EOF, TIST, pad */
static void create_eti(struct orc_rx *rx) {
  static uint8_t logical[DABGPU_CIF_BITS];
  static uint8_t soft[4 * (9216 + 6)];
  uint8_t eti[6144];
  memset(eti, 0, sizeof eti);
  int e1 = orc_init_eti(eti, &rx->ens);
  memcpy(eti + e1, rx->cif_fib[0], 96);
  int e = e1 + 96;
  orc_time_deinterleave(logical, rx->cif_msc);
  for (int i = 0; i < 64; i++) {
    const struct orc_subch *sc = &rx->ens.subchans[i];
    if (sc->id < 0) continue;
    const uint8_t *src = logical + sc->start_cu * 64;
    int len = sc->eepprot ? orc_eep_depuncture(soft, src, sc->protlev, sc->size, sc->bitrate)
                          : orc_uep_depuncture(soft, src, sc->uep_index);
    int bits = len / 4 - 6;
    int obytes = ((bits / 8) + 7) & 0xfff8;
    orc_viterbi(soft, eti + e, (unsigned)bits);
    orc_descramble(eti + e, obytes);
    e += obytes;
  }
  uint16_t crc = (uint16_t)~orc_crc16(eti + e1, e - e1, 0xffff);
  eti[e++] = (uint8_t)(crc >> 8);
  eti[e++] = (uint8_t)(crc & 0xff);
  eti[e++] = 0xff; /* RFU */
  eti[e++] = 0xff;
  for (int k = 0; k < 4; k++) eti[e++] = 0xff; /* TIST unused */
  memset(eti + e, 0x55, 6144 - e);
  if (rx->cb) rx->cb(rx->ctx, eti);
  if (++rx->ens.CIFCount_lo == 250) {
    rx->ens.CIFCount_lo = 0;
    if (++rx->ens.CIFCount_hi == 20) rx->ens.CIFCount_hi = 0;
  }
}

/* dab.c:35-99 dab_process_frame() */
void orc_rx_process_frame(struct orc_rx *rx) {
  struct orc_tf *tf = &rx->tfs[rx->tfidx];
  rx->last = tf;
  if (tf->no_fic) { /* fic.c:167-175: no FIC in the received data, replace with NULL FIBs */
    static const uint8_t null_fib[32] = {0xff, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xa8, 0xa8};
    for (int i = 0; i < 12; i++) {
      memcpy(tf->fibs + 32 * i, null_fib, 32);
      tf->crc_ok[i] = 1;
    }
    tf->ok_count = 12;
  } else {
    tf->ok_count = orc_fic_decode(tf->fic, tf->fibs, tf->crc_ok);
  }
  if (tf->ok_count > 0) orc_fib_decode(&rx->tf_info, tf->fibs, tf->crc_ok, 12);

  if (tf->ok_count == 12) {
    rx->okcount++;
    if (rx->okcount >= 10 && !rx->locked) rx->locked = 1;
  } else {
    rx->okcount = 0;
    if (rx->locked) {
      rx->locked = 0;
      rx->ncifs = 0;
      rx->tfidx = 0;
      return;
    }
  }
  if (!rx->locked) return;

  orc_merge_info(&rx->ens, &rx->tf_info);
  if (rx->ncifs < 16) {
    for (int k = 0; k < 4; k++) {
      rx->cif_fib[rx->ncifs] = tf->fibs + 96 * k;
      rx->cif_msc[rx->ncifs] = tf->msc + (size_t)DABGPU_CIF_BITS * k;
      rx->ncifs++;
    }
  } else {
    for (int k = 0; k < 4; k++) {
      create_eti(rx);
      memmove(rx->cif_fib, rx->cif_fib + 1, 15 * sizeof rx->cif_fib[0]);
      memmove(rx->cif_msc, rx->cif_msc + 1, 15 * sizeof rx->cif_msc[0]);
      rx->cif_fib[15] = tf->fibs + 96 * k;
      rx->cif_msc[15] = tf->msc + (size_t)DABGPU_CIF_BITS * k;
    }
  }
  rx->tfidx = (rx->tfidx + 1) % 5;
}

struct eti_sink {
  uint8_t *out;
  long cap, n;
};
static void sink_cb(void *ctx, const uint8_t *eti) {
  struct eti_sink *s = (struct eti_sink *)ctx;
  if (s->out && (s->n + 1) * 6144 <= s->cap) memcpy(s->out + s->n * 6144, eti, 6144);
  s->n++;
}

long orc_run_backend(const uint8_t *tfs, long n_tf, uint8_t *eti_out, long eti_cap,
                     uint8_t *fibs_out, uint8_t *crc_out) {
  struct eti_sink sink = {eti_out, eti_cap, 0};
  struct orc_rx *rx = orc_rx_new(sink_cb, &sink);
  for (long t = 0; t < n_tf; t++) {
    memcpy(orc_rx_fic_slot(rx), tfs + t * 230400L, 9216);
    memcpy(orc_rx_msc_slot(rx), tfs + t * 230400L + 9216, 221184);
    orc_rx_process_frame(rx);
    if (fibs_out) memcpy(fibs_out + t * 384, rx->last->fibs, 384);
    if (crc_out) memcpy(crc_out + t * 12, rx->last->crc_ok, 12);
  }
  orc_rx_free(rx);
  return sink.n;
}

/* ==================================================================================
 * 7. RTL-SDR front-end: synchronisers               (src/sdr_sync.c)
 * ================================================================================== */

static void dft(int n, int sign, const double *in, double *out) {
  fftw_plan p = fftw_plan_dft_1d(n, (fftw_complex *)in, (fftw_complex *)out, sign, FFTW_ESTIMATE);
  fftw_execute(p);
  fftw_destroy_plan(p);
}

static void prs_value(int c, double *re, double *im) {
  static uint8_t q[1536];
  static int ready;
  if (!ready) {
    dabgpu_build_prs(q);
    ready = 1;
  }
  static const double cr[4] = {1, 0, -1, 0}, ci[4] = {0, 1, 0, -1};
  *re = cr[q[c]];
  *im = ci[q[c]];
}

/* sdr_sync.c:34-68: null-symbol energy gate on every 10th sample of the first 2656;
 * on a miss, the minimum of a 266-tap sliding sum over the 10x-decimated |real|
 * gives the null position.  Returns a shift in BYTES (2 per sample). */
uint32_t orc_coarse_time_sync(const int8_t *real, int force) {
  float e = 0;
  for (int k = 0; k < 2656; k += 10) e += (float)abs(real[k]);
  if (e < 5000.0f && !force) return 0;
  const int nwin = (196608 - 2656) / 10; /* 19395 windows are searched */
  float best = 9999999.0f;
  uint32_t best_pos = 0;
  for (int w = 0; w < nwin; w++) {
    float s = 0;
    for (int k = 0; k < 2656; k += 10) s += (float)abs(real[10 * w + k]);
    if (s < best) {
      best = s;
      best_pos = (uint32_t)(10 * w);
    }
  }
  return best_pos * 2;
}

/* sdr_sync.c:71-202: FFT of the PRS useful part, gather 1536 carriers (upper half
 * taken from bins i-765, sic), multiply by conj(PRS), 1536-point inverse DFT, argmax
 * of the magnitude (compared in float).  Returns a shift in bytes. */
int32_t orc_fine_time_sync(const double *frame) {
  static double spec[2 * 2048], prod[2 * 1536], corr[2 * 1536];
  dft(2048, FFTW_FORWARD, frame + 2 * (2656 + 504), spec);
  for (int i = 0; i < 1536; i++) {
    int bin = i < 768 ? i + 1280 : i - 765;
    double pr, pi;
    prs_value(i, &pr, &pi);
    double a = spec[2 * bin], b = spec[2 * bin + 1];
    prod[2 * i] = a * pr + b * pi;     /* (a+jb) * (pr - j pi) */
    prod[2 * i + 1] = b * pr - a * pi;
  }
  dft(1536, FFTW_BACKWARD, prod, corr);
  uint32_t pos = 0;
  float best = -99999.0f;
  for (int i = 0; i < 1536; i++) {
    float m = (float)sqrt(corr[2 * i] * corr[2 * i] + corr[2 * i + 1] * corr[2 * i + 1]);
    if (m > best) {
      best = m;
      pos = (uint32_t)i;
    }
  }
  return pos < 768 ? (int32_t)(pos * 2 + 16) : (int32_t)((pos - 1536u) * 2u);
}

/* sdr_sync.c:205-258: for 29 carrier offsets k, correlate 128 carriers of the
 * fftshifted spectrum with conj(PRS[14..141]), 128-point inverse DFT, keep the
 * offset whose peak magnitude is largest (first one on ties). */
int32_t orc_coarse_freq_sync(const double *sh) {
  double prod[2 * 128], corr[2 * 128];
  float gbest = -99999.0f;
  int gk = 0;
  for (int k = -14; k <= 14; k++) {
    for (int s = 0; s < 128; s++) {
      double pr, pi;
      prs_value(14 + s, &pr, &pi);
      double a = sh[2 * (14 + k + 256 + s)], b = sh[2 * (14 + k + 256 + s) + 1];
      prod[2 * s] = pr * a + pi * b;
      prod[2 * s + 1] = pr * b - pi * a;
    }
    dft(128, FFTW_BACKWARD, prod, corr);
    float best = -99999.0f;
    for (int s = 0; s < 128; s++) {
      float m = (float)sqrt(corr[2 * s] * corr[2 * s] + corr[2 * s + 1] * corr[2 * s + 1]);
      if (m > best) best = m;
    }
    if (best > gbest) {
      gbest = best;
      gk = k;
    }
  }
  return gk;
}

/* sdr_sync.c:259-302: mean phase of x[n+2048]*conj(x[n]) over the 504 guard samples
 * of the PRS (the fine_timeshift argument is overwritten with 0 at :270), in Hz */
double orc_fine_freq_corr(const double *frame) {
  double mean = 0;
  for (int i = 0; i < 504; i++) {
    const double *l = frame + 2 * (2656 + 2048 + i), *r = frame + 2 * (2656 + i);
    double re = l[0] * r[0] + l[1] * r[1];
    double im = l[1] * r[0] - l[0] * r[1];
    mean += atan2(im, re);
  }
  mean /= 504;
  return mean / (2 * M_PI) * 1000;
}

/* ==================================================================================
 * 8. RTL-SDR front-end: FIFO and sdr_demod           (src/sdr_fifo.c, src/input_sdr.c)
 * ================================================================================== */

struct orc_fifo {
  uint32_t size, start, count;
  uint8_t *elems;
};

/* sdr_fifo.c:26-35 cbWrite(): overwrite-oldest on overflow */
static void fifo_put(struct orc_fifo *f, uint8_t v) {
  f->elems[(f->start + f->count) % f->size] = v;
  if (f->count == f->size)
    f->start = (f->start + 1) % f->size;
  else
    f->count++;
}
static uint8_t fifo_get(struct orc_fifo *f) {
  uint8_t v = f->elems[f->start];
  f->start = (f->start + 1) % f->size;
  f->count--;
  return v;
}

/* sdr_fifo.c:43-61 sdr_read_fifo(): shift>0 skips `shift` bytes first; shift<=0
 * reads only bytes+shift and leaves the tail of `buf` untouched (stale) */
static void fifo_read_frame(struct orc_fifo *f, uint32_t bytes, int32_t shift, uint8_t *buf) {
  if (shift > 0) {
    for (int32_t i = 0; i < shift; i++)
      if (f->count) buf[i] = fifo_get(f);
    for (uint32_t j = 0; j < bytes; j++)
      if (f->count) buf[j] = fifo_get(f);
  } else {
    for (uint32_t j = 0; j < bytes + (uint32_t)shift; j++) buf[j] = fifo_get(f);
  }
}

struct orc_sdr {
  uint32_t frequency;
  struct orc_fifo fifo;
  uint8_t buffer[393216];
  int8_t real[196608], imag[196608];
  double *frame;     /* 196608 x 2 */
  double *symbols;   /* 76 x 2048 x 2, fftshifted */
  double *symbols_d; /* 76 x 2048 x 2 */
  int32_t coarse_timeshift, fine_timeshift, coarse_freq_shift;
  double fine_freq_shift;
  int32_t startup_delay;
  uint8_t force_timesync;
};

static struct orc_sdr *sdr_new(void) {
  struct orc_sdr *s = (struct orc_sdr *)calloc(1, sizeof *s);
  s->fifo.size = 196608 * 2 * 4; /* input_sdr.c:170: four frames */
  s->fifo.elems = (uint8_t *)calloc(s->fifo.size, 1);
  s->frame = (double *)malloc(sizeof(double) * 2 * 196608);
  s->symbols = (double *)calloc(2 * 76 * 2048, sizeof(double));
  s->symbols_d = (double *)calloc(2 * 76 * 2048, sizeof(double));
  return s;
}
static void sdr_free(struct orc_sdr *s) {
  free(s->fifo.elems);
  free(s->frame);
  free(s->symbols);
  free(s->symbols_d);
  free(s);
}

static void fft_shifted(const double *in, double *out) {
  double tmp[2 * 2048];
  dft(2048, FFTW_FORWARD, in, tmp);
  memcpy(out, tmp + 2 * 1024, sizeof(double) * 2 * 1024); /* input_sdr.c:95-104 */
  memcpy(out + 2 * 1024, tmp, sizeof(double) * 2 * 1024);
}

/* input_sdr.c:27-165 sdr_demod(); fic_out 9216 bytes, msc_out 221184 bytes */
static int sdr_demod(struct orc_sdr *s, const uint8_t *chunk, int chunk_len, uint8_t *fic_out,
                     uint8_t *msc_out) {
  static uint16_t rev[1536];
  static int rev_ready;
  if (!rev_ready) {
    dabgpu_build_freq_deint(rev);
    rev_ready = 1;
  }
  s->coarse_freq_shift = 0;
  for (int i = 0; i < chunk_len; i++) fifo_put(&s->fifo, chunk[i]);
  if (s->fifo.count < 196608 * 3) return 0;
  fifo_read_frame(&s->fifo, 196608 * 2, s->coarse_timeshift + s->fine_timeshift, s->buffer);
  if (s->startup_delay <= 0) { /* GAIN_SETTLE_TIME == 0: first frame is discarded */
    s->startup_delay++;
    return 0;
  }
  for (int j = 0; j < 196608; j++) {
    s->real[j] = (int8_t)(s->buffer[2 * j] - 127); /* 255 wraps to -128 */
    s->imag[j] = (int8_t)(s->buffer[2 * j + 1] - 127);
  }
  s->coarse_timeshift = (int32_t)orc_coarse_time_sync(s->real, s->force_timesync);
  s->force_timesync = 0;
  if (s->coarse_timeshift) return 0;
  for (int j = 0; j < 196608; j++) {
    s->frame[2 * j] = s->real[j];
    s->frame[2 * j + 1] = s->imag[j];
  }
  s->fine_timeshift = orc_fine_time_sync(s->frame);
  /* input_sdr.c:86-88 is dead code: coarse_freq_shift was zeroed above */
  fft_shifted(s->frame + 2 * (2656 + 505 + s->fine_timeshift), s->symbols);
  s->coarse_freq_shift = orc_coarse_freq_sync(s->symbols);
  if (abs(s->coarse_freq_shift) > 1) {
    s->force_timesync = 1;
    return 0;
  }
  s->fine_freq_shift = orc_fine_freq_corr(s->frame);

  for (int l = 0; l < 76; l++)
    fft_shifted(s->frame + 2 * (2656 + 2552 * l + 504), s->symbols + 2 * 2048 * l);
  for (int l = 1; l < 76; l++) {
    const double *cur = s->symbols + 2 * 2048 * l, *prv = cur - 2 * 2048;
    double *d = s->symbols_d + 2 * 2048 * l;
    for (int i = 0; i < 2048; i++) {
      double den = prv[2 * i] * prv[2 * i] + prv[2 * i + 1] * prv[2 * i + 1];
      d[2 * i] = (cur[2 * i] * prv[2 * i] + cur[2 * i + 1] * prv[2 * i + 1]) / den;
      d[2 * i + 1] = (cur[2 * i] * prv[2 * i + 1] - cur[2 * i + 1] * prv[2 * i]) / den;
    }
  }
  for (int l = 1; l < 76; l++) {
    uint8_t *dst = l < 4 ? fic_out + 3072 * (l - 1) : msc_out + 3072 * (l - 4);
    const double *d = s->symbols_d + 2 * 2048 * l;
    int c = 0;
    for (int i = 256; i < 1793; i++) { /* fftshifted bins of the 1536 carriers */
      if (i == 1024) continue;
      int n = rev[c++];
      dst[n] = d[2 * i] > 0 ? 0 : 1;
      dst[1536 + n] = d[2 * i + 1] > 0 ? 1 : 0;
    }
  }
  return 1;
}

int orc_demod_frame(const uint8_t *frame, int force_timesync, int32_t *cts, int32_t *fts,
                    int32_t *cfs, double *ffs, double *symbols, double *symbols_d,
                    uint8_t *demapped) {
  struct orc_sdr *s = sdr_new();
  uint8_t *out = (uint8_t *)calloc(230400, 1);
  s->startup_delay = 1;
  s->force_timesync = (uint8_t)force_timesync;
  for (long i = 0; i < 393216; i++) fifo_put(&s->fifo, frame[i]);
  for (long i = 0; i < 196608; i++) fifo_put(&s->fifo, frame[i]);
  int ok = sdr_demod(s, NULL, 0, out, out + 9216);
  *cts = s->coarse_timeshift;
  *fts = s->fine_timeshift;
  *cfs = s->coarse_freq_shift;
  *ffs = s->fine_freq_shift;
  if (symbols) memcpy(symbols, s->symbols, sizeof(double) * 2 * 76 * 2048);
  if (symbols_d) memcpy(symbols_d, s->symbols_d, sizeof(double) * 2 * 76 * 2048);
  if (demapped) memcpy(demapped, out, 230400);
  free(out);
  sdr_free(s);
  return ok;
}

/* glibc random_r() TYPE_3 (x^31 + x^3 + 1 additive feedback), as used by rand() */
void orc_srand(struct orc_rand *s, unsigned seed) {
  int32_t st[34];
  if (seed == 0) seed = 1;
  st[0] = (int32_t)seed;
  for (int i = 1; i < 31; i++) {
    long hi = st[i - 1] / 127773, lo = st[i - 1] % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    st[i] = (int32_t)w;
  }
  memcpy(s->r, st, sizeof(int32_t) * 31);
  s->idx = 0; /* rptr = 0, fptr = 3 */
  for (int i = 0; i < 310; i++) (void)orc_rand(s);
}
int orc_rand(struct orc_rand *s) {
  int r = s->idx, f = (s->idx + 3) % 31;
  uint32_t v = (uint32_t)s->r[f] + (uint32_t)s->r[r];
  s->r[f] = (int32_t)v;
  s->idx = (s->idx + 1) % 31;
  return (int)(v >> 1);
}

/* dab2eti.c:75-103: tuner feedback after every callback, also on calls that
 * produced no frame (then coarse_freq_shift is 0 and fine_freq_shift is stale) */
static void tuner_feedback(struct orc_sdr *s, struct orc_rand *rng) {
  int cfs = s->coarse_freq_shift;
  if (abs(cfs) > 1) s->frequency = cfs < 0 ? s->frequency - 1000 : s->frequency + 1000;
  if (abs(cfs) == 1) {
    int d = orc_rand(rng) % 1000;
    s->frequency = cfs < 0 ? s->frequency - (uint32_t)d : s->frequency + (uint32_t)d;
  }
  /* abs() here is the int abs: the double is truncated toward zero first */
  if (abs(cfs) < 1 && abs((int)s->fine_freq_shift) > 50)
    s->frequency = (uint32_t)((double)s->frequency + s->fine_freq_shift / 3);
}

/* dab2eti.c:60-130 (demod_thread_fn + rtlsdr_callback) over an in-memory capture,
 * with the virtual tuner described in oracle/ref_harness.c */
long orc_run_iq(const uint8_t *iq, long nbytes, int chunk, uint32_t f0, unsigned rand_seed,
                uint8_t *eti_out, long eti_cap, struct orc_call_trace *trace, long trace_cap,
                long *n_calls, uint8_t *tf_out, long tf_cap, long *n_tfs) {
  struct eti_sink sink = {eti_out, eti_cap, 0};
  struct orc_rx *rx = orc_rx_new(sink_cb, &sink);
  struct orc_sdr *s = sdr_new();
  struct orc_rand rng;
  uint8_t *cbuf = (uint8_t *)malloc(262144);
  long calls = 0, tfs = 0;
  orc_srand(&rng, rand_seed);
  s->frequency = f0;
  if (chunk <= 0 || chunk > 262144) chunk = 262144;
  for (long pos = 0; pos + chunk <= nbytes; pos += chunk) {
    double df = (double)s->frequency - (double)f0;
    const uint8_t *src = iq + pos;
    if (df != 0.0) {
      for (long b = 0; b < chunk; b += 2) {
        double n = (double)((pos + b) / 2);
        double ph = -2.0 * M_PI * df * n / 2048000.0;
        double c = cos(ph), sn = sin(ph);
        double xr = (double)iq[pos + b] - 127.0, xi = (double)iq[pos + b + 1] - 127.0;
        double qr = floor(xr * c - xi * sn + 0.5) + 127.0, qi = floor(xr * sn + xi * c + 0.5) + 127.0;
        cbuf[b] = (uint8_t)(qr < 0 ? 0 : qr > 255 ? 255 : qr);
        cbuf[b + 1] = (uint8_t)(qi < 0 ? 0 : qi > 255 ? 255 : qi);
      }
      src = cbuf;
    }
    int ok = sdr_demod(s, src, chunk, orc_rx_fic_slot(rx), orc_rx_msc_slot(rx));
    if (ok) {
      if (tf_out && tfs < tf_cap) {
        memcpy(tf_out + tfs * 230400L, orc_rx_fic_slot(rx), 9216);
        memcpy(tf_out + tfs * 230400L + 9216, orc_rx_msc_slot(rx), 221184);
      }
      tfs++;
      orc_rx_process_frame(rx);
    }
    tuner_feedback(s, &rng);
    if (trace && calls < trace_cap) {
      struct orc_call_trace *t = &trace[calls];
      t->ok = ok;
      t->coarse_timeshift = s->coarse_timeshift;
      t->fine_timeshift = s->fine_timeshift;
      t->coarse_freq_shift = s->coarse_freq_shift;
      t->fine_freq_shift = s->fine_freq_shift;
      t->frequency = s->frequency;
      t->locked = rx->locked;
      t->eti_frames = (int32_t)sink.n;
    }
    calls++;
  }
  if (n_calls) *n_calls = calls;
  if (n_tfs) *n_tfs = tfs;
  free(cbuf);
  sdr_free(s);
  orc_rx_free(rx);
  return sink.n;
}

/* Wavefinder producer: input_wf.c:23-40 (wf_demap_symbol), :65-115 (wf_read_frame) and the loop of
 * do_wf_decode (dab2eti.c:251-272) over a buffer of 524-byte USB packets; the device's timing loop
 * (wf_sync.c) is taken as locked, as in oracle/ref_harness.c:ref_run_wf */
static void wf_demap_symbol(uint8_t *dst, const uint8_t *src, const uint16_t *rev) {
  int q = 0;
  for (int i = 0; i < 192; i++, src += 2) {
    const int k = (src[1] << 8) | src[0];
    for (int j = 15; j > 0; j -= 2, q++) {
      dst[rev[q]] = (uint8_t)((k >> j) & 1);
      dst[rev[q] + 1536] = (uint8_t)((k >> (j - 1)) & 1);
    }
  }
}
/* returns 0 with a complete frame in *tf, 1 at the end of the packets */
static int wf_read_frame(const uint8_t *packets, long n_packets, long *pos, struct orc_tf *tf, const uint16_t *rev) {
  int fic_read[3] = {0, 0, 0};
  while (*pos < n_packets) {
    const uint8_t *buf = packets + 524 * (*pos)++;
    const int symbol = buf[2];
    if (symbol == 1) continue; /* PRS: wf_prs_assemble (hardware loop) */
    if (symbol == 0) {
      tf->no_fic = !(fic_read[0] && fic_read[1] && fic_read[2]);
      return 0;
    } else if (symbol <= 4) {
      wf_demap_symbol(tf->fic + 3072 * (symbol - 2), buf + 12, rev);
      fic_read[symbol - 2] = 1;
    } else if (symbol <= 76) {
      wf_demap_symbol(tf->msc + 3072 * (symbol - 5), buf + 12, rev);
    }
  }
  return 1;
}
long orc_run_wf(const uint8_t *packets, long n_packets, uint8_t *eti_out, long eti_cap) {
  struct eti_sink sink = {eti_out, eti_cap, 0};
  struct orc_rx *rx = orc_rx_new(sink_cb, &sink);
  uint16_t rev[1536];
  long pos = 0;
  dabgpu_build_freq_deint(rev);
  if (wf_read_frame(packets, n_packets, &pos, &rx->tfs[0], rev) == 0) /* first frame: read and discarded */
    while (wf_read_frame(packets, n_packets, &pos, &rx->tfs[rx->tfidx], rev) == 0) orc_rx_process_frame(rx);
  orc_rx_free(rx);
  return sink.n;
}

/* Streaming form of orc_run_iq (same contract as ref_stream_* in oracle/ref_harness.c): the
 * receiver state lives across calls so that bench.py can time one bounded sample per step. */
struct orc_stream {
  struct eti_sink sink;
  struct orc_rx *rx;
  struct orc_sdr *s;
  struct orc_rand rng;
  uint8_t *cbuf;
  uint32_t f0;
  long pos;
};

void *orc_stream_open(uint32_t f0, unsigned rand_seed) {
  struct orc_stream *h = (struct orc_stream *)calloc(1, sizeof *h);
  h->rx = orc_rx_new(sink_cb, &h->sink);
  h->s = sdr_new();
  h->cbuf = (uint8_t *)malloc(262144);
  orc_srand(&h->rng, rand_seed);
  h->s->frequency = f0;
  h->f0 = f0;
  return h;
}

long orc_stream_feed(void *hv, const uint8_t *iq, long nbytes, int chunk, uint8_t *eti_out, long eti_cap) {
  struct orc_stream *h = (struct orc_stream *)hv;
  struct orc_sdr *s = h->s;
  h->sink.out = eti_out;
  h->sink.cap = eti_cap;
  h->sink.n = 0;
  if (chunk <= 0 || chunk > 262144) chunk = 262144;
  for (long pos = 0; pos + chunk <= nbytes; pos += chunk, h->pos += chunk) {
    double df = (double)s->frequency - (double)h->f0;
    const uint8_t *src = iq + pos;
    if (df != 0.0) {
      for (long b = 0; b < chunk; b += 2) {
        double n = (double)((h->pos + b) / 2);
        double ph = -2.0 * M_PI * df * n / 2048000.0;
        double c = cos(ph), sn = sin(ph);
        double xr = (double)iq[pos + b] - 127.0, xi = (double)iq[pos + b + 1] - 127.0;
        double qr = floor(xr * c - xi * sn + 0.5) + 127.0, qi = floor(xr * sn + xi * c + 0.5) + 127.0;
        h->cbuf[b] = (uint8_t)(qr < 0 ? 0 : qr > 255 ? 255 : qr);
        h->cbuf[b + 1] = (uint8_t)(qi < 0 ? 0 : qi > 255 ? 255 : qi);
      }
      src = h->cbuf;
    }
    if (sdr_demod(s, src, chunk, orc_rx_fic_slot(h->rx), orc_rx_msc_slot(h->rx))) orc_rx_process_frame(h->rx);
    tuner_feedback(s, &h->rng);
  }
  h->sink.out = NULL;
  return h->sink.n;
}

int orc_stream_locked(void *hv) { return ((struct orc_stream *)hv)->rx->locked; }

void orc_stream_close(void *hv) {
  struct orc_stream *h = (struct orc_stream *)hv;
  free(h->cbuf);
  sdr_free(h->s);
  orc_rx_free(h->rx);
  free(h);
}

/* ==================================================================================
 * 9. Table accessors for the tests
 * ================================================================================== */
void orc_tab_freq_deint(uint16_t rev[1536]) { dabgpu_build_freq_deint(rev); }
void orc_tab_prs(uint8_t q[1536]) { dabgpu_build_prs(q); }
uint32_t orc_tab_puncture_mask(int pi) { return dabgpu_puncture_mask(pi); }
int orc_tab_shape(int kind, int a, int b, int32_t *out) {
  dabgpu_cw_shape sh;
  int rc = 0;
  if (kind == 0)
    dabgpu_shape_fic(&sh);
  else if (kind == 1)
    rc = dabgpu_shape_uep(&sh, a);
  else
    rc = dabgpu_shape_eep(&sh, a, b, -1);
  if (rc) return rc;
  memcpy(out, &sh, sizeof sh);
  return 0;
}
