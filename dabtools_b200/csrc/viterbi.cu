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
// One lane = one codeword, 64 path metrics as 4 x u8 in 16 registers, state s in
// byte (s & 3) of register (s >> 2).  Butterfly group k (0..7) takes A = R[k]
// (states 4k..4k+3) and B = R[k+8] (those +32) and produces the even successors
// 8k,8k+2,.. (E) and the odd ones (O) with byte-SWAR arithmetic:
//     t = a + 0x7f7f7f7f - b   -> bit 7 of each byte = (a > b)      (a,b < 128)
//     m = prmt(t, 0xba98)      -> 0xff where b wins (sign replicate)
//     new = (b & m) | (a & ~m) ; decision bits accumulate as m & (0x01010101 << k)
// and two prmt re-interleave E/O into registers 2k, 2k+1.  Branch distances for the
// four symbol classes {c, ~c} (generators 0 and 3 are equal, so only 8 of the 16
// symbols occur) come from a 256-entry shared-memory table indexed by the step byte.
#include "viterbi.cuh"

#include <algorithm>

namespace dabgpu {

__constant__ uint32_t c_prbs_le[288];  // 1152 PRBS bytes as little-endian words

int viterbi_init_constants() {
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
// prmt selector that gathers, for butterfly group k, the distance of the "input bit 0
// from predecessor p" branch of p = 4k..4k+3 out of {D (bytes 0-3), Dc (bytes 4-7)}
__host__ __device__ constexpr uint32_t cx_group_sel(int k) {
  uint32_t s = 0;
  for (int j = 0; j < 4; j++) {
    int v = cx_branch_sym(2 * (4 * k + j));
    int pol = v & 1;  // class representative has bit0 == 0
    int cls = ((pol ? ~v : v) >> 1) & 3;
    s |= (uint32_t)(cls | (pol << 2)) << (4 * j);
  }
  return s;
}

// Two formulations of "compare, select, record the decision" for 4 states at once.  t holds
// (a > b) in bit 7 of every byte.
//  - ALU flavour: m = prmt(t, 0xba98) (sign-replicate), dec |= m & (0x01010101 << K)   [2 ALU ops]
//  - FMA flavour: f = (t & 0x80808080) >> 7 via IMAD.HI, m = f * 255, dec = f * 2^K + dec
//                                                                   [1 ALU op + 3 FMA-pipe ops]
// The integer ALU pipe is the kernel's bottleneck (PRMT/LOP3/IADD3 issue at half rate), the FMA
// pipe is mostly idle, so a share of the groups uses the second form to balance the two pipes.
template <int K, bool FMA_FLAVOUR>
__device__ __forceinline__ uint32_t select_and_record(uint32_t t, uint32_t a, uint32_t b, uint32_t &dec) {
  uint32_t m;
  if (FMA_FLAVOUR) {
    const uint32_t f = __umulhi(t & 0x80808080u, 1u << 25);  // 0x01 where b wins
    m = f * 255u;
    dec = f * (1u << K) + dec;
  } else {
    m = prmt(t, 0u, 0xba98u);
    dec |= m & (0x01010101u << K);
  }
  return (b & m) | (a & ~m);
}

template <int K>
__device__ __forceinline__ void butterfly(uint32_t A, uint32_t B, uint32_t D, uint32_t Dc,
                                          uint32_t &R0, uint32_t &R1, uint32_t &decE, uint32_t &decO) {
  constexpr uint32_t sel = cx_group_sel(K);
  const uint32_t X = prmt(D, Dc, sel);           // distance of branch p -> 2p
  const uint32_t Y = prmt(D, Dc, sel ^ 0x4444u); // its complement, n - X
  const uint32_t a0 = A + X, b0 = B + Y;         // into even state: via p / via p+32
  const uint32_t a1 = A + Y, b1 = B + X;         // into odd state
  const uint32_t t0 = a0 + 0x7f7f7f7fu - b0;
  const uint32_t t1 = a1 + 0x7f7f7f7fu - b1;
  // measured on B200: the FMA flavour raises the step from 132 to 154 instructions and the decoder
  // gets 10 % slower -- it is issue-bound, not ALU-pipe-bound -- so every group uses the ALU form
  constexpr bool fma = false;
  const uint32_t E = select_and_record<K, fma>(t0, a0, b0, decE);
  const uint32_t O = select_and_record<K, fma>(t1, a1, b1, decO);
  R0 = prmt(E, O, 0x5140u);
  R1 = prmt(E, O, 0x7362u);
}

struct Metrics {
  uint32_t r[16];
};

__device__ __forceinline__ uint2 acs_step(Metrics &m, uint2 d) {
  uint32_t e = 0, o = 0;
  Metrics n;
  butterfly<0>(m.r[0], m.r[8], d.x, d.y, n.r[0], n.r[1], e, o);
  butterfly<1>(m.r[1], m.r[9], d.x, d.y, n.r[2], n.r[3], e, o);
  butterfly<2>(m.r[2], m.r[10], d.x, d.y, n.r[4], n.r[5], e, o);
  butterfly<3>(m.r[3], m.r[11], d.x, d.y, n.r[6], n.r[7], e, o);
  butterfly<4>(m.r[4], m.r[12], d.x, d.y, n.r[8], n.r[9], e, o);
  butterfly<5>(m.r[5], m.r[13], d.x, d.y, n.r[10], n.r[11], e, o);
  butterfly<6>(m.r[6], m.r[14], d.x, d.y, n.r[12], n.r[13], e, o);
  butterfly<7>(m.r[7], m.r[15], d.x, d.y, n.r[14], n.r[15], e, o);
  m = n;
  return make_uint2(e, o);
}

// decision bit of state s inside the (E,O) pair written by acs_step
__device__ __forceinline__ uint32_t decision_bit(uint2 d, uint32_t s) {
  const uint32_t w = (s & 1u) ? d.y : d.x;
  return (w >> ((((s >> 1) & 3u) << 3) + (s >> 3))) & 1u;
}

// Decode one group (up to 32 equal-length codewords, one per lane): forward pass + traceback.
__device__ __forceinline__ void decode_group(const VitGroup &g, const uint2 *lut,
                                             const uint8_t *__restrict__ steps, uint8_t *__restrict__ out,
                                             uint2 *__restrict__ dec, const VitJob *__restrict__ jobs, int lane) {
  const bool active = lane < (int)g.nlanes;
  const VitJob job = jobs[g.job0 + (active ? lane : 0)];
  const uint4 *row = reinterpret_cast<const uint4 *>(steps + job.in_off);
  uint2 *decp = dec + g.dec_off + lane;
  const uint32_t nsteps = g.nsteps;

  // ---- forward pass ----
  Metrics m;
#pragma unroll
  for (int i = 0; i < 16; i++) m.r[i] = 0x30303030u;  // "unreachable" = 48
  m.r[0] = 0x30303000u;                               // start state 0
  const uint32_t nchunks = (nsteps + 15u) >> 4;
  uint4 nxt = row[0];
  for (uint32_t c = 0; c < nchunks; c++) {
    uint4 cur = nxt;
    if (c + 1 < nchunks) nxt = row[c + 1];
    const uint32_t base = c << 4;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      uint32_t word = cur.x;
      cur.x = cur.y;
      cur.y = cur.z;
      cur.z = cur.w;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const uint32_t t = base + 4 * q + s;
        if (t < nsteps) {
          const uint2 d = lut[word & 0xffu];
          word >>= 8;
          const uint2 dd = acs_step(m, d);
          if (active) decp[(size_t)t * 32] = dd;
        }
      }
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
  uint32_t state = 0, acc = 0;
  int i = (int)nbits - 1;
  // peel so that the main loop runs whole 32-bit words
  for (; i >= 0 && (i & 31) != 31; --i) {
    const uint32_t bit = decision_bit(decp[(size_t)(i + 6) * 32], state);
    acc |= bit << (31 - (i & 31));
    state = (state >> 1) | (bit << 5);
    if ((i & 31) == 0) {
      uint32_t w = prmt(acc, 0u, 0x0123u);
      if (scr) w ^= c_prbs_le[i >> 5];
      *reinterpret_cast<uint32_t *>(dst + 4 * (i >> 5)) = w;
      acc = 0;
    }
  }
  // main loop: 16 bits per half-iteration; the decision words are fetched two half-iterations
  // ahead (their addresses do not depend on the survivor state), so ~32 loads per lane are in
  // flight while 16 are being walked
  if (i < 31) return;
  const uint2 *dp = decp + (size_t)6 * 32;  // dp[step * 32] = decisions of trellis step `step`+6
  uint2 b0[16], b1[16];
#pragma unroll
  for (int j = 0; j < 16; j++) b0[j] = dp[(size_t)(i - j) * 32];
#pragma unroll
  for (int j = 0; j < 16; j++) b1[j] = dp[(size_t)(i - 16 - j) * 32];
  for (; i >= 31; i -= 32) {
    uint2 b2[16];
    if (i >= 63) {
#pragma unroll
      for (int j = 0; j < 16; j++) b2[j] = dp[(size_t)(i - 32 - j) * 32];
    }
    acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t bit = decision_bit(b0[j], state);
      acc |= bit << j;  // bit index (i - j) & 31 = 31 - j
      state = (state >> 1) | (bit << 5);
    }
    if (i >= 63) {
#pragma unroll
      for (int j = 0; j < 16; j++) b0[j] = dp[(size_t)(i - 48 - j) * 32];
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t bit = decision_bit(b1[j], state);
      acc |= bit << (16 + j);
      state = (state >> 1) | (bit << 5);
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

// Persistent decoder: one CTA per SM, VIT_WARPS warps per CTA.  Warp w of CTA b owns work list
// `bin = b * VIT_WARPS + w` (groups bin_start[bin] .. bin_start[bin+1]).  The host packs the
// groups longest-processing-time-first, first over the warp schedulers (warp w runs on scheduler
// w % 4) so that every scheduler gets the same number of trellis steps, then over the scheduler's
// VIT_WARPS/4 warps so that they all stay busy until the end of the launch.
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) viterbi_kernel(const uint8_t *__restrict__ steps,
                                                                uint8_t *__restrict__ out,
                                                                uint2 *__restrict__ dec,
                                                                const VitJob *__restrict__ jobs,
                                                                const VitGroup *__restrict__ groups,
                                                                const uint32_t *__restrict__ bin_start) {
  __shared__ uint2 lut[256];  // step byte -> {D, Dc}: per-class distances and complements
  for (int sb = threadIdx.x; sb < 256; sb += blockDim.x) {
    const uint32_t r = sb & 15, e = sb >> 4;
    uint32_t D = 0;
    for (int c = 0; c < 4; c++) D |= (uint32_t)__popc(((2u * c) ^ r) & e) << (8 * c);
    lut[sb] = make_uint2(D, (uint32_t)__popc(e) * 0x01010101u - D);
  }
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
    o->r[i].mask = dabgpu_puncture_mask(s.r[i].pi);
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

}  // namespace dabgpu
