// msc.cu -- time de-interleave + depuncture gather, CIF plane packing, ETI assembly.
#include "msc.cuh"

namespace dabgpu {

__constant__ uint8_t c_tdi_slot[16];    // window slot that supplies plane m (misc.c:32 map[])
__constant__ uint16_t c_x8pow[16];      // x^(8*2^j) mod (x^16+x^12+x^5+1), for CRC combination

static uint16_t host_mulmod(uint16_t a, uint16_t b) {
  uint32_t r = 0;
  for (int i = 15; i >= 0; i--) {
    r = (r & 0x8000u) ? ((r << 1) ^ 0x1021u) & 0xffffu : (r << 1) & 0xffffu;
    if ((b >> i) & 1u) r ^= a;
  }
  return (uint16_t)r;
}

int msc_init_dep_tables();
int msc_soft_init_constants();
// Wavefinder producer (input_wf.c:23-40): carrier order -> frequency de-interleaved position, and the
// 2304 channel bits of a FIC group that decodes to three NULL FIBs (fic.c:150-175)
__device__ uint16_t g_rev1536[1536];
__device__ uint8_t g_null_fic_group[2304];

static void host_null_fic_group(uint8_t out[2304]) {
  // three NULL FIBs (ff 00 .. 00 a8 a8) -> energy dispersal -> K=7 r=1/4 mother code (viterbi.c:322-347)
  // -> FIC puncturing (depuncture.c:45-82 in reverse)
  uint8_t fibs[96], prbs[96];
  memset(fibs, 0, sizeof fibs);
  for (int k = 0; k < 3; k++) {
    fibs[32 * k] = 0xff;
    fibs[32 * k + 30] = fibs[32 * k + 31] = 0xa8;
  }
  dabgpu_build_prbs(prbs, 96);
  static uint8_t mother[4 * 774];
  uint32_t sr = 0;
  for (int t = 0; t < 774; t++) {
    const int bit = t < 768 ? (((fibs[t >> 3] ^ prbs[t >> 3]) >> (7 - (t & 7))) & 1) : 0;
    sr = ((sr << 1) | (uint32_t)bit) & 0x7fu;
    for (int j = 0; j < 4; j++) mother[4 * t + j] = (uint8_t)(__builtin_popcount(sr & DABGPU_POLYS[j]) & 1);
  }
  dabgpu_cw_shape sh;
  dabgpu_shape_fic(&sh);
  int n = 0;
  for (int r = 0; r < sh.n_regions; r++) {
    const uint32_t mask = dabgpu_puncture_mask(sh.r[r].pi);
    for (int pos = 0; pos < 4 * sh.r[r].steps; pos++)
      if ((mask >> (pos & 31)) & 1u) out[n++] = mother[4 * sh.r[r].step0 + pos];
  }
}

int msc_init_constants() {
  {
    uint16_t rev[1536];
    dabgpu_build_freq_deint(rev);
    CUDA_TRY(cudaMemcpyToSymbol(g_rev1536, rev, sizeof rev));
    static uint8_t grp[2304];
    host_null_fic_group(grp);
    CUDA_TRY(cudaMemcpyToSymbol(g_null_fic_group, grp, sizeof grp));
  }
  CUDA_TRY(cudaMemcpyToSymbol(c_tdi_slot, DABGPU_TDI_DELAY, 16));
  {
    int rc_soft = msc_soft_init_constants();
    if (rc_soft) return rc_soft;
  }
  uint16_t p[16];
  // x^8 mod P: a CRC register holding 1 shifted by one byte
  uint16_t v = 1;
  for (int b = 0; b < 8; b++) v = (v & 0x8000u) ? (uint16_t)((v << 1) ^ 0x1021u) : (uint16_t)(v << 1);
  p[0] = v;
  for (int j = 1; j < 16; j++) p[j] = host_mulmod(p[j - 1], p[j - 1]);
  CUDA_TRY(cudaMemcpyToSymbol(c_x8pow, p, sizeof p));
  return msc_init_dep_tables();
}

// ---- shared first half: 16 planes -> packed bits in logical order ---------------------------
// 16 x 16 bit-matrix transpose (rows in the low halves of M[0..15]): afterwards bit m of M[b] is the
// former bit b of M[m].  Four swap stages of 8 row pairs each.
__device__ __forceinline__ void transpose16x16(uint32_t (&M)[16]) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int j = 8 >> s;
    const uint32_t mask = s == 0 ? 0x00ffu : s == 1 ? 0x0f0fu : s == 2 ? 0x3333u : 0x5555u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k & j) continue;
      const uint32_t t = ((M[k] >> j) ^ M[k + j]) & mask;
      M[k + j] ^= t;
      M[k] ^= t << j;
    }
  }
}

// lin must hold CIF_WORDS + 1 words.  Logical bit i = 16 q + m comes from plane m, bit q, of the
// window slot that supplies plane m; 16 consecutive q of all 16 planes are one 16 x 16 transpose.
__device__ __forceinline__ void deinterleave_to_smem(const uint8_t *__restrict__ cifs, const CifJob &job,
                                                     uint32_t (*pl)[CIF_PLANE_WORDS], uint32_t *lin) {
  for (int idx = threadIdx.x; idx < CIF_WORDS; idx += blockDim.x) {
    const int m = idx / CIF_PLANE_WORDS, w = idx - m * CIF_PLANE_WORDS;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(cifs + job.slot_off[c_tdi_slot[m]]);
    pl[m][w] = __ldg(src + idx);
  }
  __syncthreads();
  for (int item = threadIdx.x; item < 2 * CIF_PLANE_WORDS; item += blockDim.x) {
    const int pw = item >> 1, half = item & 1;
    uint32_t M[16];
#pragma unroll
    for (int m = 0; m < 16; m++) M[m] = half ? pl[m][pw] >> 16 : pl[m][pw] & 0xffffu;
    transpose16x16(M);
    // M[b] = the 16 planes' bits of q = 32 pw + 16 half + b, i.e. logical bits 16 q .. 16 q + 15
    uint32_t *dst = lin + 16 * pw + 8 * half;
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = (M[2 * j] & 0xffffu) | (M[2 * j + 1] << 16);
  }
  if (threadIdx.x == 0) lin[CIF_WORDS] = 0;
  __syncthreads();
}

// K3: one block per logical CIF
__global__ void __launch_bounds__(256) msc_gather_kernel(const uint8_t *__restrict__ cifs,
                                                         const CifJob *__restrict__ jobs,
                                                         const SubJob *__restrict__ subs,
                                                         const ShapeDev *__restrict__ shapes,
                                                         uint8_t *__restrict__ steps) {
  __shared__ uint32_t pl[16][CIF_PLANE_WORDS];
  __shared__ uint32_t lin[CIF_WORDS + 1];
  __shared__ CifJob job;
  if (threadIdx.x < sizeof(CifJob) / 4)
    reinterpret_cast<uint32_t *>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&jobs[blockIdx.x])[threadIdx.x];
  __syncthreads();
  deinterleave_to_smem(cifs, job, pl, lin);

  for (uint32_t s = 0; s < job.nsub; s++) {
    const SubJob sj = subs[job.sub0 + s];
    const ShapeDev *sh = &shapes[sj.shape];
    const uint32_t nsteps = (uint32_t)sh->nbits + 6u;
    const uint32_t periods = vit_row_bytes(nsteps) >> 3;
    const int nreg = sh->n_regions;
    // one thread per puncturing period = 8 trellis steps = 8 step bytes (uep/eep_depuncture,
    // depuncture.c:84-132, without materialising the 4 soft symbols per step)
    for (uint32_t p = threadIdx.x; p < periods; p += blockDim.x) {
      const uint32_t t0 = p << 3;
      uint2 packed = make_uint2(0u, 0u);
      if (t0 < nsteps) {
        int r = 0;
        while (r + 1 < nreg && (int)t0 >= sh->r[r + 1].step0) r++;
        const auto &rg = sh->r[r];
        const uint32_t in = sj.in_bit0 + (uint32_t)rg.in0 + ((t0 - (uint32_t)rg.step0) >> 3) * (uint32_t)rg.ones;
        const uint32_t wi = in >> 5;
        // `ones` <= 32 consecutive channel bits starting at bit `in`
        const uint32_t x = wi < CIF_WORDS ? __funnelshift_r(lin[wi], lin[wi + 1], in & 31u) : 0u;
        uint32_t rn = 0;  // received bits as one nibble per step
        const uint32_t nt = rg.n_terms;
        for (uint32_t t = 0; t < nt; t++) rn |= (x << rg.dep_shift[t]) & rg.dep_mask[t];
        uint32_t lo = rn & 0xffffu, hi = rn >> 16;
        lo = (lo | (lo << 8)) & 0x00ff00ffu;
        hi = (hi | (hi << 8)) & 0x00ff00ffu;
        packed.x = ((lo | (lo << 4)) & 0x0f0f0f0fu) | rg.e_lo;
        packed.y = ((hi | (hi << 4)) & 0x0f0f0f0fu) | rg.e_hi;
      }
      *reinterpret_cast<uint2 *>(steps + sj.row_off + 8ull * p) = packed;
    }
  }
}

// ---- the same with a per-layout period table -----------------------------------------------------
// deposit tables for every puncturing vector: index pi (1..24) for whole periods, 32 + pi for the
// 6-step tail region (only pi = 8 occurs)
struct DepTab {
  uint32_t n_terms, e_lo, e_hi, pad;
  uint32_t shift[8];
  uint32_t mask[8];
};
__device__ DepTab g_dep[64];

static void make_dep(int pi, int steps, DepTab *d) {
  memset(d, 0, sizeof *d);
  uint32_t mask = dabgpu_puncture_mask(pi);
  if (steps < 8) mask &= (1u << (4 * steps)) - 1u;
  uint32_t taken = 0, nt = 0;
  for (int k = 0; k < 8; k++) {
    const uint32_t e = (mask >> (4 * k)) & 15u;
    if (!e) continue;
    const uint32_t shift = 4u * k - taken;  // punctured positions before step k
    taken += (uint32_t)__builtin_popcount(e);
    if (nt && d->shift[nt - 1] == shift) {
      d->mask[nt - 1] |= e << (4 * k);
    } else {
      d->shift[nt] = shift;
      d->mask[nt] = e << (4 * k);
      nt++;
    }
  }
  d->n_terms = nt;
  auto spread = [](uint32_t v) {  // nibbles of the low 16 bits -> high nibbles of 4 bytes
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    return v << 4;
  };
  d->e_lo = spread(mask & 0xffffu);
  d->e_hi = spread(mask >> 16);
}

int msc_init_dep_tables() {
  DepTab t[64];
  memset(t, 0, sizeof t);
  for (int pi = 1; pi <= 24; pi++) {
    make_dep(pi, 8, &t[pi]);
    make_dep(pi, 6, &t[32 + pi]);
  }
  CUDA_TRY(cudaMemcpyToSymbol(g_dep, t, sizeof t));
  return DABGPU_OK;
}

bool append_periods(const dabgpu_cw_shape &sh, uint32_t in_bit0, uint32_t row_off, std::vector<PeriodDesc> &out) {
  const uint32_t nsteps = (uint32_t)sh.nbits + 6u;
  const uint32_t periods = vit_row_bytes(nsteps) >> 3;
  for (uint32_t p = 0; p < periods; p++) {
    const uint32_t t0 = p << 3;
    PeriodDesc d;
    d.row_off8 = (row_off >> 3) + p;
    if (t0 >= nsteps) {
      d.in_tab = 0xffu << 16;
    } else {
      int r = 0;
      while (r + 1 < sh.n_regions && (int)t0 >= sh.r[r + 1].step0) r++;
      const int pi = sh.r[r].pi, st = sh.r[r].steps;
      if (pi < 1 || pi > 24 || (st < 8 && st != 6) || (st >= 8 && (st & 7))) return false;
      const uint32_t in = in_bit0 + (uint32_t)sh.r[r].in0 + ((t0 - (uint32_t)sh.r[r].step0) >> 3) * (8u + (uint32_t)pi);
      if (in >= 0x10000u) return false;
      d.in_tab = in | (uint32_t)(st < 8 ? 32 + pi : pi) << 16;
    }
    out.push_back(d);
  }
  return true;
}

// Host restatement of what one CIF's worth of msc_gather_periods_kernel threads do with these
// tables (for the CPU-side tests of append_periods / make_dep; not used by the product path).
// bits: the de-interleaved CIF, one byte per channel bit; steps: nsteps padded to 16 step bytes.
bool host_periods_to_steps(const dabgpu_cw_shape &shape, uint32_t in_bit0, const uint8_t *bits, uint8_t *steps) {
  std::vector<PeriodDesc> per;
  if (!append_periods(shape, in_bit0, 0, per)) return false;
  for (const PeriodDesc &d : per) {
    const uint32_t idx = d.in_tab >> 16, in = d.in_tab & 0xffffu;
    uint32_t lo = 0, hi = 0;
    if (idx != 0xffu) {
      DepTab t;
      make_dep((int)(idx & 31u), idx >= 32 ? 6 : 8, &t);
      uint32_t x = 0;
      for (int b = 0; b < 32; b++)
        if (in + b < DABGPU_CIF_BITS) x |= (uint32_t)(bits[in + b] & 1u) << b;
      uint32_t rn = 0;
      for (uint32_t k = 0; k < t.n_terms; k++) rn |= (x << t.shift[k]) & t.mask[k];
      lo = rn & 0xffffu;
      hi = rn >> 16;
      lo = (lo | (lo << 8)) & 0x00ff00ffu;
      hi = (hi | (hi << 8)) & 0x00ff00ffu;
      lo = ((lo | (lo << 4)) & 0x0f0f0f0fu) | t.e_lo;
      hi = ((hi | (hi << 4)) & 0x0f0f0f0fu) | t.e_hi;
    }
    memcpy(steps + 8ull * d.row_off8, &lo, 4);
    memcpy(steps + 8ull * d.row_off8 + 4, &hi, 4);
  }
  return true;
}

__global__ void __launch_bounds__(256) msc_gather_periods_kernel(const uint8_t *__restrict__ cifs,
                                                                 const CifJob *__restrict__ jobs,
                                                                 const PeriodDesc *__restrict__ periods,
                                                                 uint8_t *__restrict__ steps) {
  __shared__ uint32_t pl[16][CIF_PLANE_WORDS];
  __shared__ uint32_t lin[CIF_WORDS + 1];
  __shared__ CifJob job;
  __shared__ __align__(16) DepTab dep[64];
  if (threadIdx.x < sizeof(CifJob) / 4)
    reinterpret_cast<uint32_t *>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&jobs[blockIdx.x])[threadIdx.x];
  for (uint32_t i = threadIdx.x; i < sizeof(dep) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(dep)[i] = reinterpret_cast<const uint32_t *>(g_dep)[i];
  __syncthreads();
  deinterleave_to_smem(cifs, job, pl, lin);
  const uint2 *pd = reinterpret_cast<const uint2 *>(periods) + job.per0;
  uint2 *dst = reinterpret_cast<uint2 *>(steps + job.row_base);
  static_assert(sizeof(DepTab) == 80, "five 16-byte vectors per deposit table");
  const uint4 *dq = reinterpret_cast<const uint4 *>(dep);
  const uint32_t nper = job.nper;
  // Branch-free: always eight deposit terms (unused ones have mask 0), and a padding period (table
  // index 0xff) takes row 63, which is all zero, so it stores 0 whatever it read.
  // four periods per thread and pass, descriptors fetched ahead of their use
  for (uint32_t p0 = threadIdx.x; p0 < nper; p0 += 4 * blockDim.x) {
    uint2 d[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t p = p0 + u * blockDim.x;
      d[u] = p < nper ? __ldg(pd + p) : make_uint2(0xffu << 16, 0xffffffffu);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (d[u].y == 0xffffffffu) break;
      const uint32_t idx = min(d[u].x >> 16, 63u), in = d[u].x & 0xffffu;
      const uint32_t wi = min(in >> 5, (uint32_t)CIF_WORDS - 1u);
      // up to 32 consecutive channel bits starting at bit `in`
      const uint32_t x = __funnelshift_r(lin[wi], lin[wi + 1], in & 31u);
      const uint4 hd = dq[5 * idx], s0 = dq[5 * idx + 1], s1 = dq[5 * idx + 2], m0 = dq[5 * idx + 3],
                  m1 = dq[5 * idx + 4];
      // received bits as one nibble per step
      const uint32_t rn = ((x << s0.x) & m0.x) | ((x << s0.y) & m0.y) | ((x << s0.z) & m0.z) |
                          ((x << s0.w) & m0.w) | ((x << s1.x) & m1.x) | ((x << s1.y) & m1.y) |
                          ((x << s1.z) & m1.z) | ((x << s1.w) & m1.w);
      // nibbles n0 .. n7 -> the low nibbles of eight bytes: even and odd nibbles apart, then interleaved
      const uint32_t ev = rn & 0x0f0f0f0fu, od = (rn >> 4) & 0x0f0f0f0fu;
      dst[d[u].y] = make_uint2(__byte_perm(ev, od, 0x5140) | hd.y, __byte_perm(ev, od, 0x7362) | hd.z);
    }
  }
}

int launch_msc_gather_periods(const uint8_t *d_cifs, const CifJob *d_jobs, const PeriodDesc *d_periods,
                              uint8_t *d_steps, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return DABGPU_OK;
  msc_gather_periods_kernel<<<n_jobs, 256, 0, st>>>(d_cifs, d_jobs, d_periods, d_steps);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

int launch_msc_gather(const uint8_t *d_cifs, const CifJob *d_jobs, const SubJob *d_subs,
                      const ShapeDev *d_shapes, uint8_t *d_steps, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return DABGPU_OK;
  msc_gather_kernel<<<n_jobs, 256, 0, st>>>(d_cifs, d_jobs, d_subs, d_shapes, d_steps);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

__global__ void __launch_bounds__(256) deinterleave_bytes_kernel(const uint8_t *__restrict__ cifs,
                                                                 const CifJob *__restrict__ jobs,
                                                                 uint8_t *__restrict__ out) {
  __shared__ uint32_t pl[16][CIF_PLANE_WORDS];
  __shared__ uint32_t lin[CIF_WORDS + 1];
  __shared__ CifJob job;
  if (threadIdx.x < sizeof(CifJob) / 4)
    reinterpret_cast<uint32_t *>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&jobs[blockIdx.x])[threadIdx.x];
  __syncthreads();
  deinterleave_to_smem(cifs, job, pl, lin);
  uint8_t *dst = out + (size_t)blockIdx.x * DABGPU_CIF_BITS;
  for (int i = threadIdx.x; i < DABGPU_CIF_BITS; i += blockDim.x) dst[i] = (lin[i >> 5] >> (i & 31)) & 1u;
}

int launch_deinterleave_bytes(const uint8_t *d_cifs, const CifJob *d_jobs, uint8_t *d_out, int n_jobs,
                              cudaStream_t st) {
  if (n_jobs <= 0) return DABGPU_OK;
  deinterleave_bytes_kernel<<<n_jobs, 256, 0, st>>>(d_cifs, d_jobs, d_out);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- demapped bytes -> planes -------------------------------------------------------------------
__global__ void pack_planes_kernel(const uint8_t *__restrict__ msc, uint64_t tf_stride,
                                   const uint64_t *__restrict__ dst_off, uint8_t *__restrict__ cifs, int n_tf) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (tf, cif, word)
  const uint64_t cif = idx / CIF_WORDS;
  const int word = (int)(idx - cif * CIF_WORDS);
  if (cif >= (uint64_t)n_tf * 4) return;
  const int m = word / CIF_PLANE_WORDS, w = word - m * CIF_PLANE_WORDS;
  const uint8_t *src = msc + (cif >> 2) * tf_stride + (cif & 3) * (uint64_t)DABGPU_CIF_BITS;
  uint32_t v = 0;
#pragma unroll 8
  for (int b = 0; b < 32; b++) v |= (uint32_t)(src[16 * (32 * w + b) + m] & 1u) << b;
  reinterpret_cast<uint32_t *>(cifs + dst_off[cif])[word] = v;
}

int launch_pack_planes(const uint8_t *d_msc_bytes, uint64_t tf_stride, const uint64_t *d_dst_off,
                       uint8_t *d_cifs, int n_tf, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_tf * 4 * CIF_WORDS;
  if (!total) return DABGPU_OK;
  pack_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_msc_bytes, tf_stride, d_dst_off,
                                                                      d_cifs, n_tf);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- ETI(NI) assembly: one warp per frame -----------------------------------------------------
__device__ __forceinline__ uint32_t crc_byte(uint32_t crc, uint32_t byte) {
  crc ^= byte << 8;
#pragma unroll
  for (int k = 0; k < 8; k++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xffffu : (crc << 1) & 0xffffu;
  return crc;
}
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll
  for (int i = 15; i >= 0; i--) {
    r = (r & 0x8000u) ? ((r << 1) ^ 0x1021u) & 0xffffu : (r << 1) & 0xffffu;
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}
// v * x^(8*nbytes) mod P
__device__ __forceinline__ uint32_t crc_shift(uint32_t v, uint32_t nbytes) {
  for (int j = 0; nbytes; j++, nbytes >>= 1)
    if (nbytes & 1u) v = gf_mulmod(v, c_x8pow[j]);
  return v;
}
// byte-at-a-time CRC step with the 256-entry table T[b] = CRC register after byte b from 0
__device__ __forceinline__ uint32_t crc_byte_tab(const uint16_t *tab, uint32_t crc, uint32_t byte) {
  return ((crc << 8) & 0xffffu) ^ tab[(crc >> 8) ^ byte];
}
// CRC-16-CCITT (init 0xffff, no final xor) of n bytes (n % 4 == 0, p 4-byte aligned) by one warp
__device__ uint32_t warp_crc16(const uint16_t *tab, const uint8_t *p, uint32_t n, int lane) {
  const uint32_t words = n >> 2;
  const uint32_t per = (words + 31) / 32;
  const uint32_t w0 = min(words, per * lane), w1 = min(words, per * (lane + 1));
  uint32_t crc = 0;
  const uint32_t *pw = reinterpret_cast<const uint32_t *>(p);
  for (uint32_t w = w0; w < w1; w++) {
    const uint32_t v = pw[w];
    crc = crc_byte_tab(tab, crc, v & 0xffu);
    crc = crc_byte_tab(tab, crc, (v >> 8) & 0xffu);
    crc = crc_byte_tab(tab, crc, (v >> 16) & 0xffu);
    crc = crc_byte_tab(tab, crc, v >> 24);
  }
  uint32_t part = crc_shift(crc, 4 * (words - w1));
  if (lane == 0) part ^= crc_shift(0xffffu, n);
#pragma unroll
  for (int o = 16; o; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
  return part;
}

__global__ void __launch_bounds__(128) eti_pack_kernel(const EtiJob *__restrict__ jobs,
                                                       const EnsDev *__restrict__ ens,
                                                       const uint8_t *__restrict__ fibs,
                                                       uint8_t *__restrict__ eti_all, int n_frames) {
  __shared__ uint16_t crc_tab[256];
  for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) crc_tab[b] = (uint16_t)crc_byte(0u, b);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const EtiJob job = jobs[f];
  const EnsDev *en = &ens[job.ens];
  uint8_t *eti = eti_all + (size_t)f * DABGPU_ETI_BYTES;
  const uint32_t nst = en->nst, fl = en->fl;
  // SYNC + FC (misc.c:158-183)
  if (lane == 0) {
    const bool odd = job.cif_lo & 1;
    eti[0] = 0xff;
    eti[1] = odd ? 0xf8 : 0x07;
    eti[2] = odd ? 0xc5 : 0x3a;
    eti[3] = odd ? 0x49 : 0xb6;
    eti[4] = job.cif_lo;
    eti[5] = (uint8_t)(0x80u | nst);
    const uint32_t fp = (job.cif_hi * 250u + job.cif_lo) & 7u;
    eti[6] = (uint8_t)((fp << 5) | (1u << 3) | ((fl >> 8) & 7u));
    eti[7] = (uint8_t)(fl & 0xffu);
  }
  // STC (misc.c:185-202)
  for (uint32_t j = lane; j < nst; j += 32)
    *reinterpret_cast<uint32_t *>(eti + 8 + 4 * j) = *reinterpret_cast<const uint32_t *>(en->stc[j]);
  __syncwarp();
  const uint32_t eoh = 8 + 4 * nst;
  // EOH: MNSC + HCRC over eti[4 .. eoh+2) (misc.c:203-211)
  if (lane == 0) {
    eti[eoh] = 0xff;
    eti[eoh + 1] = 0xff;
    uint32_t crc = 0xffffu;
    for (uint32_t i = 4; i < eoh + 2; i++) crc = crc_byte_tab(crc_tab, crc, eti[i]);
    crc = ~crc & 0xffffu;
    eti[eoh + 2] = (uint8_t)(crc >> 8);
    eti[eoh + 3] = (uint8_t)(crc & 0xffu);
  }
  const uint32_t e1 = eoh + 4;
  // FIC of the oldest CIF (misc.c:239)
  if (lane < 24)
    *reinterpret_cast<uint32_t *>(eti + e1 + 4 * lane) = *reinterpret_cast<const uint32_t *>(fibs + job.fib_off + 4 * lane);
  __syncwarp();
  // EOF: CRC over MST = FIC + sub-channel payload (misc.c:280-296)
  const uint32_t mst = 96 + en->payload;
  uint32_t crc = ~warp_crc16(crc_tab, eti + e1, mst, lane) & 0xffffu;
  uint32_t e = e1 + mst;
  if (lane == 0) {
    eti[e] = (uint8_t)(crc >> 8);
    eti[e + 1] = (uint8_t)(crc & 0xffu);
    eti[e + 2] = 0xff;  // RFU
    eti[e + 3] = 0xff;
    *reinterpret_cast<uint32_t *>(eti + e + 4) = 0xffffffffu;  // TIST unused
  }
  e += 8;
  for (uint32_t i = e + 4 * lane; i < DABGPU_ETI_BYTES; i += 128) *reinterpret_cast<uint32_t *>(eti + i) = 0x55555555u;
}

int launch_eti_pack(const EtiJob *d_jobs, const EnsDev *d_ens, const uint8_t *d_fibs, uint8_t *d_eti,
                    int n_frames, cudaStream_t st) {
  if (n_frames <= 0) return DABGPU_OK;
  eti_pack_kernel<<<(n_frames + 3) / 4, 128, 0, st>>>(d_jobs, d_ens, d_fibs, d_eti, n_frames);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- soft-decision data path (opt-in: DABGPU_ENGINE_SOFT) ------------------------------
// The same receiver with one byte per channel bit instead of one bit: a received symbol value around
// the erasure level 128 (what the reference's depuncturers would hand to viterbi() if its demapper
// did not slice, depuncture.c:36-43).  CIFs are stored as 55296 bytes in logical order; the time
// de-interleaver + depuncturer below produce rows of four symbols per trellis step for
// viterbi_soft_kernel, with the same job / period descriptors as the hard-decision gather.
enum { SOFT_CIF_BYTES = 55296 };
__constant__ uint32_t c_pmask[25];  // puncturing vectors PI = 1..24 (dab_tables.c:102-127)

int msc_soft_init_constants() {
  uint32_t m[25] = {0};
  for (int pi = 1; pi <= 24; pi++) m[pi] = dabgpu_puncture_mask(pi);
  CUDA_TRY(cudaMemcpyToSymbol(c_pmask, m, sizeof m));
  return DABGPU_OK;
}

// the 4 CIFs of a demapped soft transmission frame -> their slots of the soft CIF store
__global__ void __launch_bounds__(256) soft_store_kernel(const uint8_t *__restrict__ msc, uint64_t tf_stride,
                                                         const uint64_t *__restrict__ dst_off,
                                                         uint8_t *__restrict__ cifs_soft) {
  const int cif = blockIdx.y;  // 4 * tf + c
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= SOFT_CIF_BYTES / 16) return;
  const uint8_t *src = msc + (uint64_t)(cif >> 2) * tf_stride + (uint64_t)(cif & 3) * SOFT_CIF_BYTES;
  // hard-store offsets (slot * CIF_BYTES) address the soft store slot by slot
  uint8_t *dst = cifs_soft + dst_off[cif] / CIF_BYTES * SOFT_CIF_BYTES;
  reinterpret_cast<uint4 *>(dst)[v] = reinterpret_cast<const uint4 *>(src)[v];
}
int launch_soft_store(const uint8_t *d_msc_bytes, uint64_t tf_stride, const uint64_t *d_dst_off, uint8_t *d_cifs_soft,
                      int n_tf, cudaStream_t st) {
  if (n_tf <= 0) return DABGPU_OK;
  soft_store_kernel<<<dim3((SOFT_CIF_BYTES / 16 + 255) / 256, 4 * n_tf), 256, 0, st>>>(d_msc_bytes, tf_stride,
                                                                                       d_dst_off, d_cifs_soft);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// time_deinterleave + uep/eep_depuncture (misc.c:29-39, depuncture.c:84-132) on symbols: one thread
// per puncturing period (8 trellis steps -> 8 words of 4 symbols; punctured positions = 128)
__global__ void __launch_bounds__(128) msc_soft_gather_kernel(const uint8_t *__restrict__ cifs_soft,
                                                              const CifJob *__restrict__ jobs,
                                                              const PeriodDesc *__restrict__ periods,
                                                              uint8_t *__restrict__ rows) {
  const CifJob &job = jobs[blockIdx.x];
  for (uint32_t p = threadIdx.x; p < job.nper; p += blockDim.x) {
    const PeriodDesc d = periods[job.per0 + p];
    const uint32_t idx = d.in_tab >> 16, in = d.in_tab & 0xffffu;
    uint32_t w[8];
#pragma unroll
    for (int s = 0; s < 8; s++) w[s] = 0x80808080u;
    if (idx != 0xffu) {
      const uint32_t mask = c_pmask[idx & 31u];
      const int nst = idx >= 32u ? 6 : 8;
      uint32_t i = in;
#pragma unroll
      for (int s = 0; s < 8; s++) {
        if (s < nst) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if ((mask >> (4 * s + j)) & 1u) {
              const uint8_t *slot = cifs_soft + job.slot_off[c_tdi_slot[i & 15u]] / CIF_BYTES * SOFT_CIF_BYTES;
              const uint32_t v = i < (uint32_t)DABGPU_CIF_BITS ? slot[i] : 128u;
              w[s] = (w[s] & ~(0xffu << (8 * j))) | (v << (8 * j));
              i++;
            }
          }
        }
      }
    }
    uint4 *dst = reinterpret_cast<uint4 *>(rows + 4ull * (job.row_base + 8ull * d.row_off8));
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
int launch_msc_soft_gather(const uint8_t *d_cifs_soft, const CifJob *d_jobs, const PeriodDesc *d_periods,
                           uint8_t *d_rows, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return DABGPU_OK;
  msc_soft_gather_kernel<<<n_jobs, 128, 0, st>>>(d_cifs_soft, d_jobs, d_periods, d_rows);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// fic_depuncture (depuncture.c:45-82) on symbols: codeword cw = 4 * (frame index) + group reads its
// 2304 symbols at fic + d_index[frame] * stride + group * 2304
__global__ void __launch_bounds__(256) fic_soft_rows_kernel(const uint8_t *__restrict__ fic, uint64_t stride,
                                                            const uint32_t *__restrict__ index,
                                                            uint32_t *__restrict__ rows, int n_cw,
                                                            const ShapeDev *__restrict__ shape) {
  const uint32_t row_words = vit_soft_row_bytes(774) / 4;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t cw = idx / row_words;
  const uint32_t t = (uint32_t)(idx % row_words);
  if (cw >= (uint64_t)n_cw) return;
  uint32_t w = 0x80808080u;
  if (t < 774u) {
    int r = 0;
    while (r + 1 < shape->n_regions && (int)t >= shape->r[r + 1].step0) r++;
    const uint32_t mask = shape->r[r].mask;
    const uint32_t rel = 4u * (t - (uint32_t)shape->r[r].step0), per = rel >> 5, pos0 = rel & 31u;
    const uint64_t frame = index ? (uint64_t)index[cw >> 2] : cw >> 2;
    const uint8_t *src = fic + frame * stride + (cw & 3u) * 2304u + (uint32_t)shape->r[r].in0 +
                         per * (uint32_t)shape->r[r].ones;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t pos = pos0 + j;
      if ((mask >> pos) & 1u) {
        const uint32_t v = src[__popc(mask & ((1u << pos) - 1u))];
        w = (w & ~(0xffu << (8 * j))) | (v << (8 * j));
      }
    }
  }
  rows[cw * row_words + t] = w;
}
int launch_fic_soft_rows(const uint8_t *d_fic, uint64_t stride, const uint32_t *d_index, uint8_t *d_rows, int n_cw,
                         const ShapeDev *d_shape, cudaStream_t st) {
  const uint64_t total = (uint64_t)n_cw * (vit_soft_row_bytes(774) / 4);
  if (!total) return DABGPU_OK;
  fic_soft_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_fic, stride, d_index,
                                                                        reinterpret_cast<uint32_t *>(d_rows), n_cw, d_shape);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- Wavefinder producer (input_wf.c:23-115) -----------------------------------------
// One CTA per 524-byte USB packet: byte 2 is the symbol number (2..4 FIC, 5..76 MSC), bytes 12..395
// are 192 little-endian words of 8 DQPSK decisions each in carrier order; wf_demap_symbol's combined
// frequency de-interleave + demap (input_wf.c:23-40) writes them as the common byte-per-bit symbol.
// tf_ring: [S][5][230400], the reference's dab->tfs[5]; slot[s] = dab->tfidx of stream s.
__global__ void __launch_bounds__(192) wf_demap_kernel(const uint8_t *__restrict__ packets, uint64_t pitch,
                                                       const int32_t *__restrict__ n_packets,
                                                       const int32_t *__restrict__ slot, uint8_t *__restrict__ tf_ring,
                                                       uint32_t *__restrict__ fic_seen) {
  const int s = blockIdx.y;
  if ((int)blockIdx.x >= n_packets[s]) return;
  const uint8_t *pkt = packets + (uint64_t)s * pitch + (uint64_t)blockIdx.x * 524u;
  const int sym = pkt[2];
  if (sym < 2 || sym > 76) return;  // NULL and phase reference symbols carry no data
  uint8_t *dst = tf_ring + ((uint64_t)s * 5u + (uint32_t)slot[s]) * 230400u + (uint32_t)(sym - 2) * 3072u;
  const int i = threadIdx.x;
  const uint32_t k = (uint32_t)pkt[12 + 2 * i] | ((uint32_t)pkt[13 + 2 * i] << 8);
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const uint32_t qq = g_rev1536[8 * i + p];
    dst[qq] = (uint8_t)((k >> (15 - 2 * p)) & 1u);
    dst[qq + 1536] = (uint8_t)((k >> (14 - 2 * p)) & 1u);
  }
  if (i == 0 && sym <= 4) atomicOr(&fic_seen[s], 1u << (sym - 2));
}
// tfs[tfidx] of every stream -> the contiguous frame batch process_demapped() takes; a stream that
// did not receive all three FIC symbols gets the channel bits of NULL FIBs instead, which the FIC
// chain then decodes to exactly what fic_decode() substitutes for has_fic == 0 (fic.c:167-175)
__global__ void __launch_bounds__(256) wf_finish_kernel(const uint8_t *__restrict__ tf_ring,
                                                        const int32_t *__restrict__ slot,
                                                        const uint32_t *__restrict__ fic_seen,
                                                        uint8_t *__restrict__ out) {
  const int s = blockIdx.y;
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte vector of the 230400-byte frame
  if (v >= 230400u / 16u) return;
  const uint8_t *src = tf_ring + ((uint64_t)s * 5u + (uint32_t)slot[s]) * 230400u;
  uint4 val = reinterpret_cast<const uint4 *>(src)[v];
  if (v < 9216u / 16u && fic_seen[s] != 7u) val = reinterpret_cast<const uint4 *>(g_null_fic_group)[v % 144u];
  reinterpret_cast<uint4 *>(out + (uint64_t)s * 230400u)[v] = val;
}

int launch_wf_demap(const uint8_t *d_packets, uint64_t pitch, const int32_t *d_n_packets, int max_packets,
                    const int32_t *d_slot, uint8_t *d_tf_ring, uint32_t *d_fic_seen, uint8_t *d_tf_out, int n_streams,
                    cudaStream_t st) {
  if (n_streams <= 0) return DABGPU_OK;
  CUDA_TRY(cudaMemsetAsync(d_fic_seen, 0, (size_t)n_streams * 4, st));
  if (max_packets > 0) {
    wf_demap_kernel<<<dim3(max_packets, n_streams), 192, 0, st>>>(d_packets, pitch, d_n_packets, d_slot, d_tf_ring,
                                                                   d_fic_seen);
    LAUNCH_CHECK();
  }
  wf_finish_kernel<<<dim3((230400 / 16 + 255) / 256, n_streams), 256, 0, st>>>(d_tf_ring, d_slot, d_fic_seen, d_tf_out);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// ---- ETI consumers on the device ---------------------------------------------------
// eti2mpa.c:32-67 per frame: FICF / NST from the FC, walk the STC for the SubChId, copy the
// sub-channel's STL*8 bytes from the MST.  One warp per frame.  Unlike eti2mpa, which latches offset
// and length from its first frame, every frame is parsed, so a re-organised multiplex is followed.
// out_len[f] = bytes written for frame f, or -1 when the frame does not carry the sub-channel.
__global__ void __launch_bounds__(128) eti_extract_kernel(const uint8_t *__restrict__ eti_all, int n_frames,
                                                          int subchid, uint8_t *__restrict__ out, uint64_t out_pitch,
                                                          int32_t *__restrict__ out_len) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const uint8_t *eti = eti_all + (size_t)f * DABGPU_ETI_BYTES;
  const uint32_t ficf = eti[5] >> 7, nst = eti[5] & 0x7fu;
  // lane j looks at STC entries j, j+32: SubChId match and the STLs before it
  uint32_t before = 0;
  int32_t len = -1, hit = 0x7fffffff;
  for (uint32_t j = lane; j < nst; j += 32) {
    const uint8_t *w = eti + 8 + 4 * j;
    if ((int)(w[0] >> 2) == subchid && (int32_t)j < hit) hit = (int32_t)j;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) hit = min(hit, __shfl_xor_sync(0xffffffffu, hit, o));
  if (hit != 0x7fffffff) {
    for (uint32_t j = lane; j < (uint32_t)hit; j += 32) {
      const uint8_t *w = eti + 8 + 4 * j;
      before += ((((uint32_t)w[2] & 3u) << 8) | w[3]) * 8u;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    const uint8_t *w = eti + 8 + 4 * hit;
    len = (int32_t)(((((uint32_t)w[2] & 3u) << 8) | w[3]) * 8u);
    const uint32_t src0 = 12u + 4u * nst + ficf * 96u + before;
    if (src0 + (uint32_t)len > DABGPU_ETI_BYTES || (uint64_t)len > out_pitch) {
      len = -1;  // a corrupt header must not read or write out of bounds
    } else {
      uint8_t *dst = out + (uint64_t)f * out_pitch;
      // the MST starts 4-byte aligned (12 + 4 NST + 96) and sub-channel sizes are multiples of 8
      const uint32_t *s4 = reinterpret_cast<const uint32_t *>(eti + src0);
      if ((out_pitch & 3) == 0 && ((uintptr_t)out & 3) == 0)
        for (uint32_t i = lane; i < (uint32_t)len / 4; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = s4[i];
      else
        for (uint32_t i = lane; i < (uint32_t)len; i += 32) dst[i] = eti[src0 + i];
    }
  }
  if (lane == 0) out_len[f] = len;
}

int launch_eti_extract(const uint8_t *d_eti, int n_frames, int subchid, uint8_t *d_out, uint64_t out_pitch,
                       int32_t *d_len, cudaStream_t st) {
  if (n_frames <= 0) return DABGPU_OK;
  eti_extract_kernel<<<(n_frames + 3) / 4, 128, 0, st>>>(d_eti, n_frames, subchid, d_out, out_pitch, d_len);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// Frame check ("Check Sync etc", eti2mpa.c:37; TODO.md:10-11): per frame a bit mask of what is wrong
// with an ETI(NI) frame as misc.c:153-296 builds it -- ERR/FSYNC, header CRC, FL against the STC,
// end-of-frame CRC over the MST, padding.  0 = frame is consistent.
__global__ void __launch_bounds__(128) eti_check_kernel(const uint8_t *__restrict__ eti_all, int n_frames,
                                                        uint32_t *__restrict__ flags) {
  __shared__ uint16_t crc_tab[256];
  for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) crc_tab[b] = (uint16_t)crc_byte(0u, b);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const uint8_t *eti = eti_all + (size_t)f * DABGPU_ETI_BYTES;
  uint32_t bad = 0;
  const uint32_t sync = ((uint32_t)eti[1] << 16) | ((uint32_t)eti[2] << 8) | eti[3];
  if (eti[0] != 0xff || (sync != 0x073ab6u && sync != 0xf8c549u)) bad |= DABGPU_ETI_BAD_SYNC;
  if (((sync == 0xf8c549u) ? 1u : 0u) != (eti[4] & 1u)) bad |= DABGPU_ETI_BAD_SYNC;  // FSYNC alternates with FCT
  const uint32_t ficf = eti[5] >> 7, nst = eti[5] & 0x7fu;
  const uint32_t fl = (((uint32_t)eti[6] & 7u) << 8) | eti[7];
  const uint32_t eoh = 8 + 4 * nst;
  uint32_t stl_sum = 0;
  for (uint32_t j = lane; j < nst; j += 32) stl_sum += (((uint32_t)eti[8 + 4 * j + 2] & 3u) << 8) | eti[8 + 4 * j + 3];
#pragma unroll
  for (int o = 16; o; o >>= 1) stl_sum += __shfl_xor_sync(0xffffffffu, stl_sum, o);
  const uint32_t mst = ficf * 96u + 8u * stl_sum;
  if (fl != nst + 1 + ficf * 24u + 2u * stl_sum || ((eti[6] >> 3) & 3u) != 1u) bad |= DABGPU_ETI_BAD_FC;
  if (eoh + 4 + mst + 8 > DABGPU_ETI_BYTES) {
    bad |= DABGPU_ETI_BAD_FC;
  } else {
    uint32_t crc = 0xffffu;  // header CRC over FC + STC + MNSC (misc.c:207-211)
    if (lane == 0) {
      for (uint32_t i = 4; i < eoh + 2; i++) crc = crc_byte_tab(crc_tab, crc, eti[i]);
      crc = ~crc & 0xffffu;
      if (crc != (((uint32_t)eti[eoh + 2] << 8) | eti[eoh + 3])) bad |= DABGPU_ETI_BAD_HCRC;
    }
    const uint32_t e1 = eoh + 4, e = e1 + mst;
    const uint32_t c2 = ~warp_crc16(crc_tab, eti + e1, mst, lane) & 0xffffu;
    if (c2 != (((uint32_t)eti[e] << 8) | eti[e + 1])) bad |= DABGPU_ETI_BAD_EOF_CRC;
    uint32_t padbad = 0;
    for (uint32_t i = e + 8 + lane; i < DABGPU_ETI_BYTES; i += 32) padbad |= eti[i] != 0x55;
    if (__any_sync(0xffffffffu, padbad)) bad |= DABGPU_ETI_BAD_PADDING;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  if (lane == 0) flags[f] = bad;
}

int launch_eti_check(const uint8_t *d_eti, int n_frames, uint32_t *d_flags, cudaStream_t st) {
  if (n_frames <= 0) return DABGPU_OK;
  eti_check_kernel<<<(n_frames + 3) / 4, 128, 0, st>>>(d_eti, n_frames, d_flags);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

}  // namespace dabgpu
