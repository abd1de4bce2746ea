// viterbi.cu -- batched K=7 r=1/4 Viterbi (one codeword per lane), step-byte producers,
// FIB CRC and descrambler kernels.  See viterbi.cuh for the data formats.
//
// Why this is bit-exact with src/viterbi.c:352-452
// -------------------------------------------------
// The reference maximises sum(mettab[expected][rx]) with mettab(match)=+3,
// mettab(mismatch)=-7, mettab(erasure)=0 (gen_met(amp=1,noise=1,scale=4)), i.e. a
// branch metric of 3n - 10d with n = transmitted symbols in the step (identical on
// every branch) and d = Hamming distance over them.  Both candidates entering a
// state have seen the same steps, so "m1 > m0" <=> "d1 < d0": predecessor (i>>1)+32
// wins only on a strictly smaller Hamming path metric, ties keep predecessor i>>1.
// The start bias (-999999 on states != 0) only says "unreachable for the first six
// steps"; any value larger than every reachable metric of those steps reproduces the
// same survivors.  After step 6 every state is reachable and max-min <= 24, so 8-bit
// metrics with a periodic common subtraction are exact.
//
// Kernel layout
// -------------
// One lane = one codeword, 64 path metrics as 4 x u8 in 16 registers.  A butterfly pairs the
// predecessors p and p+32 (state bit 5) and produces the successors 2p and 2p+1 (state bit 0), so
// with byte-SWAR arithmetic on registers A (bit 5 = 0) and B (bit 5 = 1)
//     t = a + 0x7f7f7f7f - b   -> bit 7 of each byte = (a > b)      (a,b < 128)
//     m = prmt(t, 0xba98)      -> 0xff where b wins (sign replicate)
//     new = (b & m) | (a & ~m) ; decision bits accumulate as m & C
// the results E (successors 2p) and O (2p+1) hold the new states in the same byte positions as
// their predecessors.  Which two state bits select the byte inside a register therefore moves up
// by one position per step; a byte bit that reaches position 5 has to be swapped with the new bit 0
// (one prmt per register).  Alternating two layouts needs that swap only on every other step:
//     L0 (before even steps): byte = (s0, s2), register = s1 + 2 s3 + 4 s4 + 8 s5
//     L1 (before odd steps):  byte = (s1, s3), register = s0 + 2 s2 + 4 s4 + 8 s5
// Branch distances: generators 0 and 3 are equal, so only the 4 symbol classes {c, ~c} occur, and
// over the 8 register pairs of a step only 8 distinct 4-byte distance patterns do.  A shared-memory
// table indexed by the step byte delivers all 8 with two 128-bit loads per step and layout.
// 123 SASS instructions per 64-state step, 66 of them on the half-rate ALU pipe (LOP3, PRMT, IADD3)
// and the rest on the FMA pipe (IMAD): the ALU pipe is what bounds the kernel, so work is placed on
// the FMA pipe where that is possible even at the price of an instruction (pair_acs).
#include "viterbi.cuh"

#include <algorithm>
#include <cmath>

namespace dabgpu {

__constant__ uint32_t c_prbs_le[288];  // 1152 PRBS bytes as little-endian words

int viterbi_soft_init_constants();
int viterbi_init_constants() {
  int rc0 = viterbi_soft_init_constants();
  if (rc0) return rc0;
  uint8_t prbs[1152];
  dabgpu_build_prbs(prbs, 1152);
  uint32_t w[288];
  memcpy(w, prbs, sizeof w);
  CUDA_TRY(cudaMemcpyToSymbol(c_prbs_le, w, sizeof w));
  return DABGPU_OK;
}

// ---- compile-time code structure ------------------------------------------------------
__host__ __device__ constexpr int cx_parity(unsigned x) {
  x ^= x >> 4;
  x ^= x >> 2;
  x ^= x >> 1;
  return (int)(x & 1u);
}
// expected symbol for 7-bit register value `reg`; bit j <-> generator j
__host__ __device__ constexpr int cx_branch_sym(int reg) {
  return cx_parity(reg & 0x6d) | (cx_parity(reg & 0x4f) << 1) | (cx_parity(reg & 0x53) << 2) |
         (cx_parity(reg & 0x6d) << 3);
}
// predecessor state (bit 5 = 0) held in byte j of register k before a step of phase ph
__host__ __device__ constexpr int cx_pre_state(int ph, int k, int j) {
  const int j0 = j & 1, j1 = j >> 1, k0 = k & 1, k1 = (k >> 1) & 1, k2 = k >> 2;
  return ph == 0 ? (j0 | (k0 << 1) | (j1 << 2) | (k1 << 3) | (k2 << 4))
                 : (k0 | (j0 << 1) | (k1 << 2) | (j1 << 3) | (k2 << 4));
}
// Distance pattern of register pair k as a prmt selector over {D (bytes 0-3), Dc (bytes 4-7)}:
// nibble j = class | polarity << 2 of the branch "predecessor in byte j, input bit 0".  The other
// three branches of the butterfly see this pattern (p+32 -> 2p+1) or its complement (^ 0x4444).
__host__ __device__ constexpr uint32_t cx_pair_sel(int ph, int k) {
  uint32_t s = 0;
  for (int j = 0; j < 4; j++) {
    const int v = cx_branch_sym(2 * cx_pre_state(ph, k, j));
    const int pol = v & 1;  // class representative has bit0 == 0
    const int cls = ((pol ? ~v : v) >> 1) & 3;
    s |= (uint32_t)(cls | (pol << 2)) << (4 * j);
  }
  return s;
}
// the 16 patterns X_0, Y_0, X_1, Y_1, ... of a phase and their 8 distinct values
__host__ __device__ constexpr uint32_t cx_code(int ph, int n) {
  return (n & 1) ? cx_pair_sel(ph, n >> 1) ^ 0x4444u : cx_pair_sel(ph, n >> 1);
}
__host__ __device__ constexpr int cx_first(int ph, int n) {
  for (int q = 0; q < n; q++)
    if (cx_code(ph, q) == cx_code(ph, n)) return q;
  return n;
}
__host__ __device__ constexpr int cx_n_distinct(int ph) {
  int r = 0;
  for (int q = 0; q < 16; q++) r += cx_first(ph, q) == q;
  return r;
}
// index (0..7) of pattern n among the distinct ones
__host__ __device__ constexpr int cx_slot(int ph, int n) {
  const int f = cx_first(ph, n);
  int r = 0;
  for (int q = 0; q < f; q++) r += cx_first(ph, q) == q;
  return r;
}
// selector of distinct pattern i
__host__ __device__ constexpr uint32_t cx_pattern(int ph, int i) {
  int r = 0;
  for (int q = 0; q < 16; q++)
    if (cx_first(ph, q) == q) {
      if (r == i) return cx_code(ph, q);
      r++;
    }
  return 0;
}
static_assert(cx_n_distinct(0) == 8 && cx_n_distinct(1) == 8, "8 distance patterns per layout");
// Where the decision of (byte j, pair k) goes inside byte j of the decision word: any bijection
// of k per byte will do; this one makes the traceback's bit index a single XOR (see tb_step)
__host__ __device__ constexpr uint32_t cx_dec_mask(int k) {
  uint32_t c = 0;
  for (int j = 0; j < 4; j++) c |= 1u << (8 * j + (((k & 3) ^ j) + (k & 4)));
  return c;
}

struct Metrics {
  uint32_t r[16];
};
#ifndef DABGPU_VIT_COMPARE
#define DABGPU_VIT_COMPARE 1
#endif
__constant__ uint32_t c_two = 2;
struct Patterns {
  uint32_t p[8];
  uint32_t ks, two;  // (0x7f - e) in every byte; 2
};

// add-compare-select of register pair K: E = successors 2p, O = successors 2p+1
template <int PH, int K>
__device__ __forceinline__ void pair_acs(const Metrics &m, const Patterns &P, uint32_t &E, uint32_t &O,
                                         uint32_t &decE, uint32_t &decO) {
  constexpr int ix = cx_slot(PH, 2 * K), iy = cx_slot(PH, 2 * K + 1);
  constexpr uint32_t C = cx_dec_mask(K);
  const uint32_t A = m.r[K], B = m.r[K + 8];
  const uint32_t X = P.p[ix], Y = P.p[iy];       // distance of p -> 2p, and its complement
  const uint32_t a0 = A + X, b0 = B + Y;         // into even state: via p / via p+32
  const uint32_t a1 = A + Y, b1 = B + X;         // into odd state
#if DABGPU_VIT_COMPARE == 1
  // bit 7 of a byte of t = (a > b), a, b < 128.  a0 - b0 = (A - B) + (X - Y) and X + Y = e (the
  // number of transmitted symbols of the step), so with D = A - B + 0x7f - e (one three-input add on
  // the half-rate ALU pipe, shared by both compares, independent of a0 .. b1) the compares are
  // t0 = 2 X + D and t1 = 2 Y + D: multiply-adds, which run on the FMA pipe (c_two is a run-time 2
  // so that ptxas cannot turn them back into LEA / IADD3).  |A - B| <= 72 at any time, so no byte of
  // D or t borrows or overflows.
  const uint32_t D = A + P.ks - B;
  const uint32_t t0 = X * P.two + D;
  const uint32_t t1 = Y * P.two + D;
#else
  const uint32_t t0 = a0 + 0x7f7f7f7fu - b0;
  const uint32_t t1 = a1 + 0x7f7f7f7fu - b1;
#endif
  // Measured on B200 (tools/demod_time.py, S = 1024, MSC batch of 2): t = a + 0x7f7f7f7f - b as two
  // IADD3 1.198 ms per launch, the form above 1.165 ms.  Slower or within 2 %: (1) mask/record on the
  // FMA pipe (f = umulhi(t & 0x80808080, 1 << 25), m = f * 255, dec = f * 2^k + dec): -10 %; (2) the
  // compare as t = b * -1 + (a * 1 + K) with run-time factors (two dependent multiply-adds per compare).
  // (3) 2 X + (0x7f - e) read from a second table so that A - B and both compares are two-input adds
  // (no IADD3 left in the loop, 121 instructions per step): 1.217 -> 1.213 ms, not worth 16 KB more
  // shared memory per CTA -- with two warps per scheduler the kernel waits on latency, not on a pipe.
  const uint32_t m0 = prmt(t0, 0u, 0xba98u);
  const uint32_t m1 = prmt(t1, 0u, 0xba98u);
  decE |= m0 & C;
  decO |= m1 & C;
  E = (b0 & m0) | (a0 & ~m0);
  O = (b1 & m1) | (a1 & ~m1);
}

template <int K>
__device__ __forceinline__ void pair_even(const Metrics &m, const Patterns &P, Metrics &n, uint32_t &dE,
                                          uint32_t &dO) {
  uint32_t E, O;
  pair_acs<0, K>(m, P, E, O, dE, dO);
  n.r[2 * (K & 3) + 8 * (K >> 2)] = E;  // L1: register = s0 + 2 s2 + 4 s4 + 8 s5, bytes unchanged
  n.r[2 * (K & 3) + 8 * (K >> 2) + 1] = O;
}
template <int K>
__device__ __forceinline__ void pair_odd(const Metrics &m, const Patterns &P, Metrics &n, uint32_t &dE,
                                         uint32_t &dO) {
  uint32_t E, O;
  pair_acs<1, K>(m, P, E, O, dE, dO);
  // back to L0: the byte bit that reached position 4 (s3 before the step) trades places with bit 0
  n.r[(K & 3) + 8 * (K >> 2)] = prmt(E, O, 0x5140u);
  n.r[(K & 3) + 8 * (K >> 2) + 4] = prmt(E, O, 0x7362u);
}

__device__ __forceinline__ uint2 acs_even(Metrics &m, const Patterns &P) {
  uint32_t e = 0, o = 0;
  Metrics n;
  pair_even<0>(m, P, n, e, o);
  pair_even<1>(m, P, n, e, o);
  pair_even<2>(m, P, n, e, o);
  pair_even<3>(m, P, n, e, o);
  pair_even<4>(m, P, n, e, o);
  pair_even<5>(m, P, n, e, o);
  pair_even<6>(m, P, n, e, o);
  pair_even<7>(m, P, n, e, o);
  m = n;
  return make_uint2(e, o);
}
__device__ __forceinline__ uint2 acs_odd(Metrics &m, const Patterns &P) {
  uint32_t e = 0, o = 0;
  Metrics n;
  pair_odd<0>(m, P, n, e, o);
  pair_odd<1>(m, P, n, e, o);
  pair_odd<2>(m, P, n, e, o);
  pair_odd<3>(m, P, n, e, o);
  pair_odd<4>(m, P, n, e, o);
  pair_odd<5>(m, P, n, e, o);
  pair_odd<6>(m, P, n, e, o);
  pair_odd<7>(m, P, n, e, o);
  m = n;
  return make_uint2(e, o);
}

// shared-memory distance table: lut[phase][half][step byte] = patterns 4*half .. 4*half+3
struct VitLut {
  uint4 t[2][2][256];
};
template <int PH>
__device__ __forceinline__ void lut_fill_row(VitLut &L, uint32_t sb, uint32_t D, uint32_t Dc) {
  L.t[PH][0][sb] = make_uint4(prmt(D, Dc, cx_pattern(PH, 0)), prmt(D, Dc, cx_pattern(PH, 1)),
                              prmt(D, Dc, cx_pattern(PH, 2)), prmt(D, Dc, cx_pattern(PH, 3)));
  L.t[PH][1][sb] = make_uint4(prmt(D, Dc, cx_pattern(PH, 4)), prmt(D, Dc, cx_pattern(PH, 5)),
                              prmt(D, Dc, cx_pattern(PH, 6)), prmt(D, Dc, cx_pattern(PH, 7)));
}
__device__ __forceinline__ void lut_fill(VitLut &L) {
  for (uint32_t sb = threadIdx.x; sb < 256; sb += blockDim.x) {
    const uint32_t r = sb & 15, e = sb >> 4;
    uint32_t D = 0;  // distance to the class representatives 2c (c = 0..3) over the transmitted symbols
    for (int c = 0; c < 4; c++) D |= (uint32_t)__popc(((2u * c) ^ r) & e) << (8 * c);
    const uint32_t Dc = (uint32_t)__popc(e) * 0x01010101u - D;
    lut_fill_row<0>(L, sb, D, Dc);
    lut_fill_row<1>(L, sb, D, Dc);
  }
}
template <int PH>
__device__ __forceinline__ Patterns lut_get(const VitLut &L, uint32_t sb) {
  const uint4 a = L.t[PH][0][sb], b = L.t[PH][1][sb];
  Patterns P;
  P.p[0] = a.x, P.p[1] = a.y, P.p[2] = a.z, P.p[3] = a.w;
  P.p[4] = b.x, P.p[5] = b.y, P.p[6] = b.z, P.p[7] = b.w;
  // patterns 2 K and 2 K + 1 of a pair are complements: their sum is e in every byte
  P.ks = 0x7f7f7f7fu - P.p[cx_slot(PH, 0)] - P.p[cx_slot(PH, 1)];
  P.two = c_two;
  return P;
}

// One traceback step.  The survivor state s (bit 0 = newest input bit) is kept as its even and odd
// bits, P = s0 + 2 s2 + 4 s4 and Q = s1 + 2 s3 + 4 s5.  For the decisions written by trellis step t:
//   word = s0; byte = (s1, s3) for even t, (s2, s4) for odd t; bit within the byte = (P >> 1) ^ Q
// (cx_dec_mask).  The predecessor is (s >> 1) | (bit << 5), i.e. P' = Q, Q' = (P >> 1) | bit << 2.
template <bool ODD>
__device__ __forceinline__ uint32_t tb_step(uint2 d, uint32_t &P, uint32_t &Q) {
  const uint32_t P1 = P >> 1;
  const uint32_t w = (P & 1u) ? d.y : d.x;
  const uint32_t byte = prmt(w, w, ODD ? P1 : Q);
  const uint32_t bit = (byte >> (P1 ^ Q)) & 1u;
  P = Q;
  Q = P1 | (bit << 2);
  return bit;
}

// Decode one group (up to 32 equal-length codewords, one per lane): forward pass + traceback.
__device__ __forceinline__ void decode_group(const VitGroup &g, const VitLut &lut,
                                             const uint8_t *__restrict__ steps, uint8_t *__restrict__ out,
                                             uint2 *__restrict__ dec, const VitJob *__restrict__ jobs, int lane) {
  const bool active = lane < (int)g.nlanes;
  const VitJob job = jobs[g.job0 + (active ? lane : 0)];
  const uint4 *row = reinterpret_cast<const uint4 *>(steps + job.in_off);
  uint2 *decp = dec + g.dec_off + lane;
  const uint32_t nsteps = g.nsteps;

  // ---- forward pass: whole 16-step chunks (rows and the decision scratch are padded to 16 steps;
  // what the padding steps do to the metrics no longer matters) ----
  Metrics m;
#pragma unroll
  for (int i = 0; i < 16; i++) m.r[i] = 0x30303030u;  // "unreachable" = 48
  m.r[0] = 0x30303000u;                               // start state 0 (L0: register 0, byte 0)
  const uint32_t nchunks = (nsteps + 15u) >> 4;
  uint4 nxt = row[0];
  uint2 *dq = decp;
  for (uint32_t c = 0; c < nchunks; c++) {
    uint4 cur = nxt;
    if (c + 1 < nchunks) nxt = row[c + 1];
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      const uint32_t word = cur.x;
      cur.x = cur.y;
      cur.y = cur.z;
      cur.z = cur.w;
      const Patterns p0 = lut_get<0>(lut, word & 0xffu);
      const Patterns p1 = lut_get<1>(lut, (word >> 8) & 0xffu);
      const Patterns p2 = lut_get<0>(lut, (word >> 16) & 0xffu);
      const Patterns p3 = lut_get<1>(lut, word >> 24);
      const uint2 d0 = acs_even(m, p0);
      const uint2 d1 = acs_odd(m, p1);
      const uint2 d2 = acs_even(m, p2);
      const uint2 d3 = acs_odd(m, p3);
      if (active) {
        dq[0] = d0;
        dq[32] = d1;
        dq[64] = d2;
        dq[96] = d3;
      }
      dq += 128;
    }
    // common subtraction: every metric is within 24 of state 0's once t >= 6
    const uint32_t m0 = m.r[0] & 0xffu;
    const uint32_t sub = (m0 > 24u ? m0 - 24u : 0u) * 0x01010101u;
#pragma unroll
    for (int i = 0; i < 16; i++) m.r[i] -= sub;
  }

  // ---- traceback from state 0 (viterbi.c:443-450) fused with descrambling ----
  if (!active) return;
  const uint32_t nbits = job.nbits;
  uint8_t *dst = out + job.out_off;
  const bool scr = job.flags & VIT_DESCRAMBLE;
  uint32_t P = 0, Q = 0, acc = 0;
  int i = (int)nbits - 1;
  // peel so that the main loop runs whole 32-bit words; bit i is decided by trellis step i + 6
  for (; i >= 0 && (i & 31) != 31; --i) {
    const uint2 d = decp[(size_t)(i + 6) * 32];
    const uint32_t bit = (i & 1) ? tb_step<true>(d, P, Q) : tb_step<false>(d, P, Q);
    acc |= bit << (31 - (i & 31));
    if ((i & 31) == 0) {
      uint32_t w = prmt(acc, 0u, 0x0123u);
      if (scr) w ^= c_prbs_le[i >> 5];
      *reinterpret_cast<uint32_t *>(dst + 4 * (i >> 5)) = w;
      acc = 0;
    }
  }
  // main loop: 16 bits per half-iteration; the decision words are fetched two half-iterations
  // ahead (their addresses do not depend on the survivor state), so ~32 loads per lane are in
  // flight while 16 are being walked.  i is odd here, so bit i - j comes from an odd trellis step
  // for even j.
  if (i < 31) return;
  // cur[-32 * n] = decisions of the step that decides bit i - n (constant offsets from one pointer)
  const uint2 *cur = decp + ((size_t)i + 6) * 32;
  uint2 b0[16], b1[16];
#pragma unroll
  for (int j = 0; j < 16; j++) b0[j] = cur[-32 * j];
#pragma unroll
  for (int j = 0; j < 16; j++) b1[j] = cur[-32 * (16 + j)];
  for (; i >= 31; i -= 32, cur -= 32 * 32) {
    uint2 b2[16];
    if (i >= 63) {
#pragma unroll
      for (int j = 0; j < 16; j++) b2[j] = cur[-32 * (32 + j)];
    }
    acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      acc |= tb_step<true>(b0[j], P, Q) << j;  // bit index (i - j) & 31 = 31 - j
      acc |= tb_step<false>(b0[j + 1], P, Q) << (j + 1);
    }
    if (i >= 63) {
#pragma unroll
      for (int j = 0; j < 16; j++) b0[j] = cur[-32 * (48 + j)];
    }
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      acc |= tb_step<true>(b1[j], P, Q) << (16 + j);
      acc |= tb_step<false>(b1[j + 1], P, Q) << (17 + j);
    }
    uint32_t w = prmt(acc, 0u, 0x0123u);
    if (scr) w ^= c_prbs_le[i >> 5];
    *reinterpret_cast<uint32_t *>(dst + 4 * (i >> 5)) = w;
    // rotate: next iteration walks b2 (steps i-32..i-47) then the freshly requested b0 (i-48..)
#pragma unroll
    for (int j = 0; j < 16; j++) {
      b1[j] = b0[j];
      b0[j] = b2[j];
    }
  }
}

// Persistent decoder: VIT_CTAS_PER_SM CTAs of VIT_WARPS warps on every SM (see viterbi.cuh for the
// measured shapes).  Warp w of CTA b owns work list `bin = b * VIT_WARPS + w` (groups
// bin_start[bin] .. bin_start[bin+1]).  The host packs the groups longest-processing-time-first over
// the CTAs' scheduler quarters (warp w runs on scheduler w % 4) and then over each quarter's
// VIT_WARPS/4 warps, so that all warps stay busy until the end of the launch.
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, WARPS == 1 ? 1 : VIT_CTAS_PER_SM) viterbi_kernel(const uint8_t *__restrict__ steps,
                                                                uint8_t *__restrict__ out,
                                                                uint2 *__restrict__ dec,
                                                                const VitJob *__restrict__ jobs,
                                                                const VitGroup *__restrict__ groups,
                                                                const uint32_t *__restrict__ bin_start) {
  __shared__ VitLut lut;  // step byte -> the 8 distance patterns of each layout
  lut_fill(lut);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t bin = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const uint32_t g0 = bin_start[bin], g1 = bin_start[bin + 1];
  for (uint32_t gi = g0; gi < g1; gi++) decode_group(groups[gi], lut, steps, out, dec, jobs, lane);
}

int launch_viterbi(const uint8_t *d_steps, uint8_t *d_out, uint2 *d_dec, const VitJob *d_jobs,
                   const VitGroup *d_groups, const uint32_t *d_bin_start, int n_ctas, int warps_per_cta,
                   cudaStream_t st) {
  if (n_ctas <= 0) return DABGPU_OK;
  if (warps_per_cta == 1)
    viterbi_kernel<1><<<n_ctas, 32, 0, st>>>(d_steps, d_out, d_dec, d_jobs, d_groups, d_bin_start);
  else
    viterbi_kernel<VIT_WARPS><<<n_ctas, 32 * VIT_WARPS, 0, st>>>(d_steps, d_out, d_dec, d_jobs, d_groups,
                                                                 d_bin_start);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

int device_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- producers ---------------------------------------------------------------------------
__global__ void prep_soft_kernel(const uint8_t *__restrict__ soft, uint64_t soft_stride,
                                 uint8_t *__restrict__ steps, uint64_t row_stride, int n_cw,
                                 uint32_t nsteps) {
  const uint32_t rowlen = vit_row_bytes(nsteps);
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t cw = idx / rowlen;
  const uint32_t t = (uint32_t)(idx % rowlen);
  if (cw >= (uint64_t)n_cw) return;
  uint32_t sb = 0;
  if (t < nsteps) {
    const uint8_t *p = soft + cw * soft_stride + 4ull * t;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t v = p[j];
      sb |= (uint32_t)(v > 128u) << j;
      sb |= (uint32_t)(v != 128u) << (4 + j);
    }
  }
  steps[cw * row_stride + t] = (uint8_t)sb;
}

int launch_prep_soft(const uint8_t *d_soft, uint64_t soft_stride, uint8_t *d_steps,
                     uint64_t row_stride, int n_cw, uint32_t nsteps, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_cw * vit_row_bytes(nsteps);
  if (!total) return DABGPU_OK;
  prep_soft_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_soft, soft_stride, d_steps,
                                                                    row_stride, n_cw, nsteps);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

void shape_to_dev(const dabgpu_cw_shape &s, ShapeDev *o) {
  memset(o, 0, sizeof *o);
  o->nbits = s.nbits;
  o->in_bits = s.in_bits;
  o->n_regions = s.n_regions;
  for (int i = 0; i < s.n_regions; i++) {
    o->r[i].steps = s.r[i].steps;
    o->r[i].step0 = s.r[i].step0;
    o->r[i].in0 = s.r[i].in0;
    o->r[i].ones = 8 + s.r[i].pi;
    uint32_t mask = dabgpu_puncture_mask(s.r[i].pi);
    o->r[i].mask = mask;
    // deposit tables; a region shorter than a period (the 6-step tail) only has that many steps
    if (s.r[i].steps < 8) mask &= (1u << (4 * s.r[i].steps)) - 1u;
    uint32_t taken = 0, nt = 0;
    for (int k = 0; k < 8; k++) {
      const uint32_t e = (mask >> (4 * k)) & 15u;
      if (!e) continue;
      const uint32_t shift = 4u * k - taken;  // punctured positions before step k
      taken += (uint32_t)__builtin_popcount(e);
      if (nt && o->r[i].dep_shift[nt - 1] == shift) {
        o->r[i].dep_mask[nt - 1] |= e << (4 * k);
      } else {
        o->r[i].dep_shift[nt] = (uint8_t)shift;
        o->r[i].dep_mask[nt] = e << (4 * k);
        nt++;
      }
    }
    o->r[i].n_terms = nt;
    auto spread = [](uint32_t v) {  // nibbles of the low 16 bits -> high nibbles of 4 bytes
      v = (v | (v << 8)) & 0x00ff00ffu;
      v = (v | (v << 4)) & 0x0f0f0f0fu;
      return v << 4;
    };
    o->r[i].e_lo = spread(mask & 0xffffu);
    o->r[i].e_hi = spread(mask >> 16);
  }
}

// one thread per 8-step puncturing period (the tail region has 6 steps)
__global__ void prep_hard_kernel(const uint8_t *__restrict__ bits, uint64_t bits_stride,
                                 uint32_t per_group, uint64_t group_stride,
                                 const uint32_t *__restrict__ group_index,
                                 uint8_t *__restrict__ steps, uint64_t row_stride, int n_cw,
                                 const ShapeDev *__restrict__ shape, uint32_t nsteps) {
  const uint32_t periods = vit_row_bytes(nsteps) / 8;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t cw = idx / periods;
  const uint32_t per = (uint32_t)(idx % periods);
  if (cw >= (uint64_t)n_cw) return;
  const uint32_t t0 = per * 8;
  uint64_t packed = 0;
  if (t0 < nsteps) {
    int r = 0;
    while (r + 1 < shape->n_regions && (int)t0 >= shape->r[r + 1].step0) r++;
    const uint32_t mask = shape->r[r].mask;
    const int in = shape->r[r].in0 + (int)((t0 - shape->r[r].step0) >> 3) * shape->r[r].ones;
    const uint64_t grp = group_index ? (uint64_t)group_index[cw / per_group] : cw / per_group;
    const uint8_t *p = bits + grp * group_stride + (cw % per_group) * bits_stride + in;
    const int nst = min(8, (int)nsteps - (int)t0);
    int k = 0;
    for (int s = 0; s < nst; s++) {
      uint32_t sb = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if ((mask >> (4 * s + j)) & 1u) {
          sb |= (uint32_t)(p[k++] & 1u) << j;
          sb |= 1u << (4 + j);
        }
      }
      packed |= (uint64_t)sb << (8 * s);
    }
  }
  *reinterpret_cast<uint64_t *>(steps + cw * row_stride + t0) = packed;
}

int launch_prep_hard(const uint8_t *d_bits, uint64_t bits_stride, uint32_t per_group,
                     uint64_t group_stride, const uint32_t *d_group_index, uint8_t *d_steps,
                     uint64_t row_stride, int n_cw,
                     const ShapeDev *d_shape, uint32_t nsteps, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_cw * (vit_row_bytes(nsteps) / 8);
  if (!total) return DABGPU_OK;
  if (per_group == 0) per_group = 1;
  prep_hard_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      d_bits, bits_stride, per_group, group_stride, d_group_index, d_steps, row_stride, n_cw, d_shape, nsteps);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

__global__ void scatter_rows_kernel(const uint8_t *__restrict__ src, uint32_t row_words,
                                    const uint64_t *__restrict__ dst_off, uint8_t *__restrict__ dst, int n_rows) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t row = idx / row_words;
  const uint32_t w = (uint32_t)(idx - row * row_words);
  if (row >= (uint64_t)n_rows) return;
  reinterpret_cast<uint32_t *>(dst + dst_off[row])[w] = reinterpret_cast<const uint32_t *>(src)[idx];
}

int launch_scatter_rows(const uint8_t *d_src, uint32_t row_bytes, const uint64_t *d_dst_off, uint8_t *d_dst,
                        int n_rows, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_rows * (row_bytes / 4);
  if (!total) return DABGPU_OK;
  scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_src, row_bytes / 4, d_dst_off, d_dst,
                                                                       n_rows);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- FIB CRC: CRC-16-CCITT over 32 bytes must leave 0x1d0f (misc.c:145-150) ---------------
__global__ void fib_crc_kernel(const uint8_t *__restrict__ fibs, uint8_t *__restrict__ ok, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t *p = reinterpret_cast<const uint32_t *>(fibs + 32ull * i);
  uint32_t crc = 0xffff;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint32_t v = p[w];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      crc ^= ((v >> (8 * b)) & 0xffu) << 8;
#pragma unroll
      for (int k = 0; k < 8; k++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xffffu : (crc << 1) & 0xffffu;
    }
  }
  ok[i] = crc == 0x1d0fu;
}

int launch_fib_crc(const uint8_t *d_fibs, uint8_t *d_ok, int n_fibs, cudaStream_t st) {
  if (n_fibs <= 0) return DABGPU_OK;
  fib_crc_kernel<<<(n_fibs + 127) / 128, 128, 0, st>>>(d_fibs, d_ok, n_fibs);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

__global__ void descramble_kernel(uint8_t *buf, uint64_t stride, int n_rows, int nbytes) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t row = idx / (uint64_t)nbytes;
  const int b = (int)(idx % (uint64_t)nbytes);
  if (row >= (uint64_t)n_rows) return;
  // the PRBS has period 511 bits = 511 bytes for byte-aligned reads
  const int k = b % 511;
  buf[row * stride + b] ^= (uint8_t)(c_prbs_le[k >> 2] >> (8 * (k & 3)));
}

int launch_descramble(uint8_t *d_buf, uint64_t stride, int n_rows, int nbytes, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_rows * (uint64_t)nbytes;
  if (!total) return DABGPU_OK;
  descramble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_buf, stride, n_rows, nbytes);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// =================================================================================================
// Soft-decision decoder (opt-in, the default path above is untouched)
// =================================================================================================
// The reference's viterbi() is a soft-decision decoder fed with three-valued input: its metric table
// gen_met(amp = 1, noise = 1.0, bias = 0, scale = 4) (viterbi.c:126-191, :455-462) is defined for the
// received symbols 121..135 around the erasure value 128 -- 4,4,4,4,4,4,3,0,-7,-17,...,-73 for a
// transmitted 0, mirrored for a 1, i.e. a log-likelihood ratio of about 10.7 per LSB -- and overflows
// to INT_MIN beyond.  This kernel decodes arbitrary symbols in that range with exactly the
// reference's arithmetic, so the unmodified viterbi.c is its oracle:
//   * per-symbol cost c_b(s) = 4 - mettab[b][s] in 0..77; a branch costs the sum over the four
//     generators (0..308).  Minimising the cost is maximising the reference's metric, and the
//     reference's "take predecessor i/2+32 iff m1 > m0" is "iff cost1 < cost0";
//   * the start bias (-999999) only marks states unreachable in the first six steps: 16384 here;
//   * any two path costs differ by at most 6 x 308 once all states are reachable, so 16-bit costs
//     with a common subtraction every 16 steps are exact.
// One codeword per lane like the hard-decision kernel; 64 costs as 2 x u16 in 32 registers, register r
// = states (2r, 2r+1).  The butterflies 2k and 2k+1 (predecessors p, p+32 -> successors 2p, 2p+1) are
// done together on the packed halves of registers k and k+16.  Symbols outside 121..135 saturate.
enum { SOFT_LO = 121, SOFT_HI = 135 };

static void host_gen_metrics(int mettab[2][256]) {
  // viterbi.c:126-191 with amp = 1, noise = 1.0, bias = 0, scale = 4
  auto phi = [](double x) { return 0.5 + 0.5 * erf(x / M_SQRT2); };
  for (int s = 0; s < 256; s++) {
    const double lo = s - 128 - 0.5, hi = s - 128 + 0.5;
    double p0, p1;
    if (s == 0) {
      p1 = phi(hi - 1);
      p0 = phi(hi + 1);
    } else if (s == 255) {
      p1 = 1 - phi(lo - 1);
      p0 = 1 - phi(lo + 1);
    } else {
      p1 = phi(hi - 1) - phi(lo - 1);
      p0 = phi(hi + 1) - phi(lo + 1);
    }
    const double m0 = log(2 * p0 / (p1 + p0)) * M_LOG2E, m1 = log(2 * p1 / (p1 + p0)) * M_LOG2E;
    mettab[0][s] = (s >= SOFT_LO && s <= SOFT_HI) ? (int)floor(m0 * 4 + 0.5) : INT32_MIN;
    mettab[1][s] = (s >= SOFT_LO && s <= SOFT_HI) ? (int)floor(m1 * 4 + 0.5) : INT32_MIN;
  }
}
void viterbi_soft_metrics(int32_t out[2][256]) {
  int t[2][256];
  host_gen_metrics(t);
  for (int b = 0; b < 2; b++)
    for (int s = 0; s < 256; s++) out[b][s] = t[b][s];
}

__device__ uint32_t g_soft_cost[256];  // cost of (0-bit | 1-bit << 16) for the saturated symbol

int viterbi_soft_init_constants() {
  int t[2][256];
  host_gen_metrics(t);
  uint32_t c[256];
  for (int s = 0; s < 256; s++) {
    const int q = s < SOFT_LO ? SOFT_LO : s > SOFT_HI ? SOFT_HI : s;
    c[s] = (uint32_t)(4 - t[0][q]) | ((uint32_t)(4 - t[1][q]) << 16);
  }
  CUDA_TRY(cudaMemcpyToSymbol(g_soft_cost, c, sizeof c));
  return DABGPU_OK;
}

// expected symbols on the branch into state `s` taken with decision `d`: the encoder register is
// s | d << 6 (viterbi.c:322-347); generators 0 and 3 are equal, so (g0, g1, g2) index 8 cost sums
__host__ __device__ constexpr int sx_combo(int reg7) {
  return cx_parity(reg7 & 0x6d) | (cx_parity(reg7 & 0x4f) << 1) | (cx_parity(reg7 & 0x53) << 2);
}

template <int K>
__device__ __forceinline__ void soft_pair(const uint32_t (&M)[32], const uint32_t (&cst)[8], uint32_t (&N)[32],
                                          uint32_t &accE, uint32_t &accO) {
  // butterflies p = 2K (low halves) and 2K+1 (high halves); X = cost of p -> 2p and p+32 -> 2p+1,
  // Y = cost of the complementary symbols (p -> 2p+1, p+32 -> 2p)
  constexpr int c0 = sx_combo(2 * (2 * K)), c1 = sx_combo(2 * (2 * K + 1));
  const uint32_t X = prmt(cst[c0], cst[c1], 0x5410u), Y = prmt(cst[c0 ^ 7], cst[c1 ^ 7], 0x5410u);
  const uint32_t A = M[K], B = M[K + 16];
  const uint32_t a0 = A + X, b0 = B + Y, a1 = A + Y, b1 = B + X;
  // bit 15 / 31 of t: b >= a, i.e. predecessor p (decision 0) is kept; all costs stay below 32768
  const uint32_t t0 = b0 + 0x80008000u - a0, t1 = b1 + 0x80008000u - a1;
  const uint32_t m0 = prmt(t0, 0u, 0xbb99u), m1 = prmt(t1, 0u, 0xbb99u);  // 0xffff where a is kept
  const uint32_t E = (a0 & m0) | (b0 & ~m0), O = (a1 & m1) | (b1 & ~m1);
  // decisions: after all 16 pairs bit K = butterfly 2K, bit 16 + K = butterfly 2K + 1
  accE = (accE >> 1) | (~t0 & 0x80008000u);
  accO = (accO >> 1) | (~t1 & 0x80008000u);
  N[2 * K] = prmt(E, O, 0x5410u);      // states 4K, 4K+1
  N[2 * K + 1] = prmt(E, O, 0x7632u);  // states 4K+2, 4K+3
}

__device__ __forceinline__ uint2 soft_step(uint32_t (&M)[32], const uint32_t *cost, uint32_t sym4) {
  // per-generator costs -> the 8 sums over (g0 = g3, g1, g2)
  const uint32_t q0 = cost[sym4 & 0xffu], q1 = cost[(sym4 >> 8) & 0xffu];
  const uint32_t q2 = cost[(sym4 >> 16) & 0xffu], q3 = cost[sym4 >> 24];
  const uint32_t u = q0 + q3;  // (u_0 | u_1 << 16): generators 0 and 3 carry the same bit
  uint32_t cst[8];
#pragma unroll
  for (int c = 0; c < 8; c++)
    cst[c] = ((c & 1) ? u >> 16 : u & 0xffffu) + ((c & 2) ? q1 >> 16 : q1 & 0xffffu) +
             ((c & 4) ? q2 >> 16 : q2 & 0xffffu);
  uint32_t N[32], accE = 0, accO = 0;
  soft_pair<0>(M, cst, N, accE, accO);
  soft_pair<1>(M, cst, N, accE, accO);
  soft_pair<2>(M, cst, N, accE, accO);
  soft_pair<3>(M, cst, N, accE, accO);
  soft_pair<4>(M, cst, N, accE, accO);
  soft_pair<5>(M, cst, N, accE, accO);
  soft_pair<6>(M, cst, N, accE, accO);
  soft_pair<7>(M, cst, N, accE, accO);
  soft_pair<8>(M, cst, N, accE, accO);
  soft_pair<9>(M, cst, N, accE, accO);
  soft_pair<10>(M, cst, N, accE, accO);
  soft_pair<11>(M, cst, N, accE, accO);
  soft_pair<12>(M, cst, N, accE, accO);
  soft_pair<13>(M, cst, N, accE, accO);
  soft_pair<14>(M, cst, N, accE, accO);
  soft_pair<15>(M, cst, N, accE, accO);
#pragma unroll
  for (int i = 0; i < 32; i++) M[i] = N[i];
  return make_uint2(accE, accO);
}

// decision of state s in the words of its step: word = s & 1 (E / O), butterfly p = s >> 1
__device__ __forceinline__ uint32_t soft_decision(uint2 d, uint32_t s) {
  const uint32_t w = (s & 1u) ? d.y : d.x, p = s >> 1;
  return (w >> (16u * (p & 1u) + (p >> 1))) & 1u;
}

// input rows: 4 symbols per step (one 32-bit word), rows 16-byte aligned, padded to whole 16-step chunks
__global__ void __launch_bounds__(128) viterbi_soft_kernel(const uint8_t *__restrict__ sym, uint8_t *__restrict__ out,
                                                           uint2 *__restrict__ dec, const VitJob *__restrict__ jobs,
                                                           const VitGroup *__restrict__ groups, int n_groups) {
  __shared__ uint32_t cost[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cost[i] = g_soft_cost[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gi >= n_groups) return;
  const VitGroup g = groups[gi];
  const bool active = lane < (int)g.nlanes;
  const VitJob job = jobs[g.job0 + (active ? lane : 0)];
  const uint4 *row = reinterpret_cast<const uint4 *>(sym + job.in_off);
  uint2 *decp = dec + g.dec_off + lane;
  const uint32_t nsteps = g.nsteps, nchunks = (nsteps + 15u) >> 4;
  uint32_t M[32];
#pragma unroll
  for (int i = 0; i < 32; i++) M[i] = 0x40004000u;  // unreachable
  M[0] = 0x40000000u;                               // start state 0
  for (uint32_t c = 0; c < nchunks; c++) {
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      const uint4 w = row[4 * c + q];
      const uint2 d0 = soft_step(M, cost, w.x), d1 = soft_step(M, cost, w.y), d2 = soft_step(M, cost, w.z), d3 = soft_step(M, cost, w.w);
      if (active) {
        uint2 *dq = decp + (size_t)(16 * c + 4 * q) * 32;
        dq[0] = d0;
        dq[32] = d1;
        dq[64] = d2;
        dq[96] = d3;
      }
    }
    // common subtraction: every cost is within 1848 of state 0's once all states are reachable
    const uint32_t m0 = M[0] & 0xffffu;
    const uint32_t sub = (m0 > 2048u ? m0 - 2048u : 0u) * 0x00010001u;
#pragma unroll
    for (int i = 0; i < 32; i++) M[i] -= sub;
  }
  if (!active) return;
  // traceback from state 0 (viterbi.c:443-450): information bit i is the decision read at step i + 6
  const uint32_t nbits = job.nbits;
  uint8_t *dst = out + job.out_off;
  const bool scr = job.flags & VIT_DESCRAMBLE;
  uint32_t state = 0, acc = 0;
  int i = (int)nbits - 1;
  while (i >= 0) {
    const int n = min(16, i + 1);
    uint2 buf[16];
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j < n) buf[j] = decp[(size_t)(i - j + 6) * 32];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (j < n) {
        const int b = i - j;
        const uint32_t bit = soft_decision(buf[j], state);
        state = (state >> 1) | (bit << 5);
        acc |= bit << (31 - (b & 31));
        if ((b & 31) == 0 || b == 0) {
          uint32_t w = prmt(acc, 0u, 0x0123u);
          if (scr) w ^= c_prbs_le[b >> 5];
          const uint32_t nb = min(4u, (nbits + 7u) / 8u - 4u * (uint32_t)(b >> 5));  // bytes of this word
          if (nb == 4)
            *reinterpret_cast<uint32_t *>(dst + 4 * (b >> 5)) = w;
          else
            for (uint32_t k = 0; k < nb; k++) dst[4 * (b >> 5) + k] = (uint8_t)(w >> (8 * k));
          acc = 0;
        }
      }
    }
    i -= n;
  }
}

// caller's symbol rows (4 bytes per step at any pitch) -> aligned, padded rows (padding = erasures)
__global__ void soft_rows_kernel(const uint8_t *__restrict__ soft, uint64_t soft_stride, uint32_t *__restrict__ rows,
                                 uint32_t row_words, int n_cw, uint32_t nsteps) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t cw = idx / row_words;
  const uint32_t t = (uint32_t)(idx % row_words);
  if (cw >= (uint64_t)n_cw) return;
  uint32_t w = 0x80808080u;
  if (t < nsteps) {
    const uint8_t *p = soft + cw * soft_stride + 4ull * t;
    w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }
  rows[cw * row_words + t] = w;
}
int launch_soft_rows(const uint8_t *d_soft, uint64_t soft_stride, uint8_t *d_rows, int n_cw, uint32_t nsteps,
                     cudaStream_t st) {
  const uint32_t row_words = vit_soft_row_bytes(nsteps) / 4;
  const uint64_t total = (uint64_t)n_cw * row_words;
  if (!total) return DABGPU_OK;
  soft_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_soft, soft_stride, reinterpret_cast<uint32_t *>(d_rows),
                                                                    row_words, n_cw, nsteps);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

int launch_viterbi_soft(const uint8_t *d_sym, uint8_t *d_out, uint2 *d_dec, const VitJob *d_jobs,
                        const VitGroup *d_groups, int n_groups, cudaStream_t st) {
  if (n_groups <= 0) return DABGPU_OK;
  viterbi_soft_kernel<<<(n_groups + 3) / 4, 128, 0, st>>>(d_sym, d_out, d_dec, d_jobs, d_groups, n_groups);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

}  // namespace dabgpu
