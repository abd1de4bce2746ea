/* dabgpu_tables.h -- DAB Transmission Mode I constants, derived from ETSI EN 300 401.
 *
 * Plain C99 / C++ header (static data + static inline generators), shared by the
 * libdabgpu host code (which uploads the derived tables to the GPU) and by the
 * CPU oracle.  Everything here is stated in the standard's own terms (1-based
 * puncturing indices PI, generator rules instead of expanded tables); the
 * reference keeps the same information in src/dab_tables.c and src/sdr_prstab.c:
 *
 *   DABGPU_UEP[]            <-> ueptable[]            dab_tables.c:16-81   (EN 300 401 table 7/36)
 *   dabgpu_eep_layout()     <-> eeptable[], eep2a8kbps dab_tables.c:87-100  (tables 8/9/39/41)
 *   dabgpu_puncture_mask()  <-> pvec[][32]            dab_tables.c:102-127 (table 29)
 *   dabgpu_build_freq_deint <-> rev_freq_deint_tab[]  dab_tables.c:130-357 (clause 14.6)
 *   dabgpu_build_prs()      <-> prs_static[]          sdr_prstab.c:1-1538  (clause 14.3.2)
 *   DABGPU_TDI_DELAY[]      <-> map[] in time_deinterleave, misc.c:32      (clause 12)
 *
 * tests/test_tables.py checks each of them against the compiled reference.
 */
#ifndef DABGPU_TABLES_H
#define DABGPU_TABLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Mode I geometry -------------------------------------------------------------- */
enum {
  DABGPU_FFT_N = 2048,          /* useful symbol length, samples */
  DABGPU_CP_N = 504,            /* guard interval */
  DABGPU_SYM_N = 2552,          /* CP + useful */
  DABGPU_NULL_N = 2656,         /* null symbol */
  DABGPU_SYMS_PER_TF = 76,      /* PRS + 75 data symbols */
  DABGPU_TF_SAMPLES = 196608,   /* 2656 + 76*2552 */
  DABGPU_TF_BYTES = 393216,     /* uint8 I,Q */
  DABGPU_CARRIERS = 1536,
  DABGPU_BITS_PER_SYM = 3072,
  DABGPU_FIC_SYMS = 3,
  DABGPU_MSC_SYMS = 72,
  DABGPU_CIFS_PER_TF = 4,
  DABGPU_SYMS_PER_CIF = 18,
  DABGPU_CIF_BITS = 55296,      /* 864 CU * 64 */
  DABGPU_FIC_BITS_PER_CIF = 2304,
  DABGPU_FIB_BYTES = 32,
  DABGPU_ETI_BYTES = 6144,
  DABGPU_TDI_DEPTH = 16
};

/* ---- convolutional code: K=7, rate 1/4, generators 133,171,145,133 (octal) -------- */
/* The reference shifts the register left and writes the polynomials bit-reversed
 * (viterbi.c:35): 0x6d, 0x4f, 0x53, 0x6d on encstate = (encstate<<1)|bit. */
static const uint8_t DABGPU_POLYS[4] = {0x6d, 0x4f, 0x53, 0x6d};

/* ---- puncturing vectors (EN 300 401 table 29) --------------------------------------
 * V_PI has 8+PI ones in 32 positions.  Position 4g (g = 0..7) is always kept; the
 * remaining ones are switched on in bit-reversed group order 0,4,2,6,1,5,3,7, first
 * at offset 1 of each group (PI 1..8), then offset 2 (PI 9..16), then offset 3
 * (PI 17..24).  Returned mask: bit p (LSB first) set <=> coded bit p of the 32-bit
 * period is transmitted. */
static inline uint32_t dabgpu_puncture_mask(int pi /* 1..24 */) {
  static const uint8_t order[8] = {0, 4, 2, 6, 1, 5, 3, 7};
  uint32_t m = 0x11111111u;
  for (int j = 0; j < pi; j++) m |= 1u << (4 * order[j & 7] + 1 + (j >> 3));
  return m;
}

/* ---- time de-interleaving (clause 12): bit i of a logical frame is delayed by
 * DABGPU_TDI_DELAY[i & 15] CIFs behind the oldest CIF of a 16-CIF window. */
static const uint8_t DABGPU_TDI_DELAY[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

/* ---- UEP protection profiles (table 7 / table 36), PI values 1-based as in the
 * standard, 0 where the fourth block is empty. ------------------------------------- */
typedef struct {
  uint16_t bitrate;   /* kbit/s */
  uint16_t size_cu;   /* sub-channel size in capacity units */
  uint8_t prot_level; /* 1 (strongest) .. 5 */
  uint8_t L[4];       /* blocks of 128 mother-code bits per puncturing region */
  uint8_t PI[4];      /* puncturing index per region */
  uint8_t pad_bits;   /* zero bits appended to fill the sub-channel */
} dabgpu_uep_profile;

static const dabgpu_uep_profile DABGPU_UEP[64] = {
  /* 32 kbit/s */
  { 32,  16, 5, { 3,  4,  17, 0}, { 5,  3,  2,  0}, 0},
  { 32,  21, 4, { 3,  3,  18, 0}, {11,  6,  5,  0}, 0},
  { 32,  24, 3, { 3,  4,  14, 3}, {15,  9,  6,  8}, 0},
  { 32,  29, 2, { 3,  4,  14, 3}, {22, 13,  8, 13}, 0},
  { 32,  35, 1, { 3,  5,  13, 3}, {24, 17, 12, 17}, 4},
  /* 48 kbit/s */
  { 48,  24, 5, { 4,  3,  26, 3}, { 5,  4,  2,  3}, 0},
  { 48,  29, 4, { 3,  4,  26, 3}, { 9,  6,  4,  6}, 0},
  { 48,  35, 3, { 3,  4,  26, 3}, {15, 10,  6,  9}, 4},
  { 48,  42, 2, { 3,  4,  26, 3}, {24, 14,  8, 15}, 0},
  { 48,  52, 1, { 3,  5,  25, 3}, {24, 18, 13, 18}, 0},
  /* 56 kbit/s */
  { 56,  29, 5, { 6, 10,  23, 3}, { 5,  4,  2,  3}, 0},
  { 56,  35, 4, { 6, 10,  23, 3}, { 9,  6,  4,  5}, 0},
  { 56,  42, 3, { 6, 12,  21, 3}, {16,  7,  6,  9}, 0},
  { 56,  52, 2, { 6, 10,  23, 3}, {23, 13,  8, 13}, 8},
  /* 64 kbit/s */
  { 64,  32, 5, { 6,  9,  31, 2}, { 5,  3,  2,  3}, 0},
  { 64,  42, 4, { 6,  9,  33, 0}, {11,  6,  5,  0}, 0},
  { 64,  48, 3, { 6, 12,  27, 3}, {16,  8,  6,  9}, 0},
  { 64,  58, 2, { 6, 10,  29, 3}, {23, 13,  8, 13}, 8},
  { 64,  70, 1, { 6, 11,  28, 3}, {24, 18, 12, 18}, 4},
  /* 80 kbit/s */
  { 80,  40, 5, { 6, 10,  41, 3}, { 6,  3,  2,  3}, 0},
  { 80,  52, 4, { 6, 10,  41, 3}, {11,  6,  5,  6}, 0},
  { 80,  58, 3, { 6, 11,  40, 3}, {16,  8,  6,  7}, 0},
  { 80,  70, 2, { 6, 10,  41, 3}, {23, 13,  8, 13}, 8},
  { 80,  84, 1, { 6, 10,  41, 3}, {24, 17, 12, 18}, 4},
  /* 96 kbit/s */
  { 96,  48, 5, { 7,  9,  53, 3}, { 5,  4,  2,  4}, 0},
  { 96,  58, 4, { 7, 10,  52, 3}, { 9,  6,  4,  6}, 0},
  { 96,  70, 3, { 6, 12,  51, 3}, {16,  9,  6, 10}, 4},
  { 96,  84, 2, { 6, 10,  53, 3}, {22, 12,  9, 12}, 0},
  { 96, 104, 1, { 6, 13,  50, 3}, {24, 18, 13, 19}, 0},
  /* 112 kbit/s */
  {112,  58, 5, {14, 17,  50, 3}, { 5,  4,  2,  5}, 0},
  {112,  70, 4, {11, 21,  49, 3}, { 9,  6,  4,  8}, 0},
  {112,  84, 3, {11, 23,  47, 3}, {16,  8,  6,  9}, 0},
  {112, 104, 2, {11, 21,  49, 3}, {23, 12,  9, 14}, 4},
  /* 128 kbit/s */
  {128,  64, 5, {12, 19,  62, 3}, { 5,  3,  2,  4}, 0},
  {128,  84, 4, {11, 21,  61, 3}, {11,  6,  5,  7}, 0},
  {128,  96, 3, {11, 22,  60, 3}, {16,  9,  6, 10}, 4},
  {128, 116, 2, {11, 21,  61, 3}, {22, 12,  9, 14}, 0},
  {128, 140, 1, {11, 20,  62, 3}, {24, 17, 13, 19}, 8},
  /* 160 kbit/s */
  {160,  80, 5, {11, 19,  87, 3}, { 5,  4,  2,  4}, 0},
  {160, 104, 4, {11, 23,  83, 3}, {11,  6,  5,  9}, 0},
  {160, 116, 3, {11, 24,  82, 3}, {16,  8,  6, 11}, 0},
  {160, 140, 2, {11, 21,  85, 3}, {22, 11,  9, 13}, 0},
  {160, 168, 1, {11, 22,  84, 3}, {24, 18, 12, 19}, 0},
  /* 192 kbit/s */
  {192,  96, 5, {11, 20, 110, 3}, { 6,  4,  2,  5}, 0},
  {192, 116, 4, {11, 22, 108, 3}, {10,  6,  4,  9}, 0},
  {192, 140, 3, {11, 24, 106, 3}, {16, 10,  6, 11}, 0},
  {192, 168, 2, {11, 20, 110, 3}, {22, 13,  9, 13}, 8},
  {192, 208, 1, {11, 21, 109, 3}, {24, 20, 13, 24}, 0},
  /* 224 kbit/s */
  {224, 116, 5, {12, 22, 131, 3}, { 8,  6,  2,  6}, 4},
  {224, 140, 4, {12, 26, 127, 3}, {12,  8,  4, 11}, 0},
  {224, 168, 3, {11, 20, 134, 3}, {16, 10,  7,  9}, 0},
  {224, 208, 2, {11, 22, 132, 3}, {24, 16, 10, 15}, 0},
  {224, 232, 1, {11, 24, 130, 3}, {24, 20, 12, 20}, 4},
  /* 256 kbit/s */
  {256, 128, 5, {11, 24, 154, 3}, { 6,  5,  2,  5}, 0},
  {256, 168, 4, {11, 24, 154, 3}, {12,  9,  5, 10}, 4},
  {256, 192, 3, {11, 27, 151, 3}, {16, 10,  7, 10}, 0},
  {256, 232, 2, {11, 22, 156, 3}, {24, 14, 10, 13}, 8},
  {256, 280, 1, {11, 26, 152, 3}, {24, 19, 14, 18}, 4},
  /* 320 kbit/s */
  {320, 160, 5, {11, 26, 200, 3}, { 8,  5,  2,  6}, 4},
  {320, 208, 4, {11, 25, 201, 3}, {13,  9,  5, 10}, 8},
  {320, 280, 2, {11, 26, 200, 3}, {24, 17,  9, 17}, 0},
  /* 384 kbit/s */
  {384, 192, 5, {11, 27, 247, 3}, { 8,  6,  2,  7}, 0},
  {384, 280, 3, {11, 24, 250, 3}, {16,  9,  7, 10}, 4},
  {384, 416, 1, {12, 28, 245, 3}, {24, 20, 14, 23}, 8},
};

/* ---- EEP profiles (tables 8, 9, 39, 41) ---------------------------------------------
 * level 0..3 = 1-A..4-A (bitrate 8n), 4..7 = 1-B..4-B (bitrate 32n).  Returns n, or
 * <=0 when size_cu is not a valid multiple. */
typedef struct {
  int n;         /* bitrate multiplier */
  int bitrate;   /* kbit/s */
  int L[2];      /* blocks of 128 mother-code bits */
  int PI[2];     /* 1-based puncturing indices */
} dabgpu_eep_layout_t;

static inline int dabgpu_eep_size_mul(int level) {
  static const int sm[8] = {12, 8, 6, 4, 27, 21, 18, 15};
  return sm[level & 7];
}

/* bitrate_field: the sub-channel's bitrate as the caller believes it (the reference keys
 * its special case on that field, depuncture.c:113); pass -1 to derive it from the size. */
static inline int dabgpu_eep_layout(int level, int size_cu, int bitrate_field, dabgpu_eep_layout_t *o) {
  /* L1 = a1*n + b1, L2 = a2*n + b2 */
  static const int a1[8] = {6, 2, 6, 4, 24, 24, 24, 24};
  static const int a2[8] = {0, 4, 0, 2, 0, 0, 0, 0};
  static const int pi1[8] = {24, 14, 8, 3, 10, 6, 4, 2};
  level &= 7;
  int n = size_cu / dabgpu_eep_size_mul(level);
  o->n = n;
  o->bitrate = n * (level < 4 ? 8 : 32);
  o->L[0] = a1[level] * n - 3;
  o->L[1] = a2[level] * n + 3;
  o->PI[0] = pi1[level];
  o->PI[1] = pi1[level] - 1;
  if (bitrate_field < 0) bitrate_field = o->bitrate;
  if (level == 1 && bitrate_field == 8) {
    /* 8 kbit/s at 2-A: the standard specifies L=(5,1), PI=(13,12).  The reference
     * (dab_tables.c:98-100, depuncture.c:113-114) uses PI=(4,13) here; parity with
     * the reference is the contract, so its values are used. */
    o->L[0] = 5;
    o->L[1] = 1;
    o->PI[0] = 4;
    o->PI[1] = 13;
  }
  return n;
}

/* ---- a codeword's puncturing layout, the common form of FIC / UEP / EEP -------------
 * Up to 4 regions of whole 128-bit blocks plus the 24-bit tail (always V_8 = 1100).
 * One trellis step = 4 mother-code bits, so a region of L blocks is 32*L steps. */
typedef struct {
  int32_t nbits;       /* information bits; trellis steps = nbits + 6 */
  int32_t in_bits;     /* punctured (transmitted) bits consumed from the channel */
  int32_t n_regions;   /* including the tail region */
  struct {
    int32_t steps;     /* trellis steps in this region */
    int32_t pi;        /* 1-based puncturing index */
    int32_t step0;     /* first trellis step of the region */
    int32_t in0;       /* first input bit of the region */
  } r[5];
} dabgpu_cw_shape;

static inline void dabgpu_shape_finish(dabgpu_cw_shape *s, const int *L, const int *PI, int nreg) {
  int step = 0, in = 0, k = 0;
  for (int i = 0; i < nreg; i++) {
    if (L[i] <= 0) continue;
    s->r[k].steps = 32 * L[i];
    s->r[k].pi = PI[i];
    s->r[k].step0 = step;
    s->r[k].in0 = in;
    step += 32 * L[i];
    in += 4 * L[i] * (8 + PI[i]);
    k++;
  }
  s->r[k].steps = 6; /* tail: 24 mother-code bits punctured with V_8 */
  s->r[k].pi = 8;
  s->r[k].step0 = step;
  s->r[k].in0 = in;
  k++;
  s->n_regions = k;
  s->nbits = step;
  s->in_bits = in + 12;
  for (; k < 5; k++) s->r[k].steps = s->r[k].pi = s->r[k].step0 = s->r[k].in0 = 0;
}

/* FIC, Mode I: 21 blocks at PI=16, 3 blocks at PI=15, tail (clause 11.2; fic.c:160-208) */
static inline void dabgpu_shape_fic(dabgpu_cw_shape *s) {
  const int L[2] = {21, 3}, PI[2] = {16, 15};
  dabgpu_shape_finish(s, L, PI, 2);
}

static inline int dabgpu_shape_uep(dabgpu_cw_shape *s, int uep_index) {
  if (uep_index < 0 || uep_index > 63) return -1;
  const dabgpu_uep_profile *p = &DABGPU_UEP[uep_index];
  int L[4], PI[4];
  for (int i = 0; i < 4; i++) { L[i] = p->L[i]; PI[i] = p->PI[i]; }
  dabgpu_shape_finish(s, L, PI, 4);
  return 0;
}

static inline int dabgpu_shape_eep(dabgpu_cw_shape *s, int level, int size_cu, int bitrate_field) {
  dabgpu_eep_layout_t e;
  if (level < 0 || level > 7) return -1;
  if (dabgpu_eep_layout(level, size_cu, bitrate_field, &e) <= 0 && !(level == 1 && bitrate_field == 8)) return -1;
  if (e.L[0] < 0) return -1;
  dabgpu_shape_finish(s, e.L, e.PI, 2);
  return 0;
}

/* ---- frequency interleaving (clause 14.6) ------------------------------------------
 * P(0)=0, P(i) = (13 P(i-1) + 511) mod 2048; the values in [256,1792] except 1024,
 * in order of appearance, give d_0..d_1535; QPSK symbol n rides carrier k = d_n-1024.
 * rev[c] = n with c = k+768 for k<0 and k+767 for k>0 (carrier index in ascending
 * frequency, DC skipped) -- the order in which a receiver walks the FFT bins. */
static inline void dabgpu_build_freq_deint(uint16_t rev[1536]) {
  int p = 0, n = 0;
  for (int i = 0; i < 2048; i++) {
    if (i > 0) p = (13 * p + 511) & 2047;
    if (p >= 256 && p <= 1792 && p != 1024) {
      int k = p - 1024;
      int c = k < 0 ? k + 768 : k + 767;
      rev[c] = (uint16_t)n++;
    }
  }
}

/* ---- phase reference symbol (clause 14.3.2, tables 43/44 for Mode I) ----------------
 * z_k = exp(j*pi/2*(h[i][k-k'] + n)).  q[c] in 0..3 is the phase in quarter turns for
 * carrier index c (same c as above). */
static inline void dabgpu_build_prs(uint8_t q[1536]) {
  static const uint8_t h[4][32] = {
    {0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1, 0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1},
    {0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0, 0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0},
    {0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3, 0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3},
    {0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2, 0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2}};
  /* per 32-carrier block, from k=-768 upward (24 blocks), then k=1 upward (24 blocks):
   * (i, n) pairs of table 44 */
  static const uint8_t blk_i[48] = {0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3,
                                    0, 3, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1};
  static const uint8_t blk_n[48] = {1, 2, 0, 1, 3, 2, 2, 3, 2, 1, 2, 3, 1, 2, 3, 3, 2, 2, 2, 1, 1, 3, 1, 2,
                                    3, 1, 1, 1, 2, 2, 1, 0, 2, 2, 3, 3, 0, 2, 1, 3, 3, 3, 3, 0, 3, 0, 1, 1};
  for (int c = 0; c < 1536; c++) {
    int b = c >> 5, j = c & 31;
    q[c] = (uint8_t)((h[blk_i[b]][j] + blk_n[b]) & 3);
  }
}

/* ---- energy dispersal PRBS, x^9 + x^5 + 1, all-ones start (clause 10; misc.c:41-58) */
static inline void dabgpu_build_prbs(uint8_t *out, int nbytes) {
  unsigned reg = 0x1ff;
  for (int i = 0; i < nbytes; i++) {
    unsigned byte = 0;
    for (int b = 0; b < 8; b++) {
      unsigned bit = ((reg >> 8) ^ (reg >> 4)) & 1u;
      reg = ((reg << 1) | bit) & 0x1ff;
      byte = (byte << 1) | bit;
    }
    out[i] = (uint8_t)byte;
  }
}

/* ---- CRC-16-CCITT (poly 0x1021), MSB first, as used for FIB, ETI header and EOF ---- */
static inline uint16_t dabgpu_crc16(const uint8_t *p, int n, uint16_t crc) {
  for (int i = 0; i < n; i++) {
    crc ^= (uint16_t)p[i] << 8;
    for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
  }
  return crc;
}

#ifdef __cplusplus
}
#endif
#endif /* DABGPU_TABLES_H */
