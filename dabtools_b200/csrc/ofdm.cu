// ofdm.cu -- RTL-SDR front-end kernels (see ofdm.cuh).
//
//  ingest_kernel     rtlsdr_callback + cbWrite loop  (dab2eti.c:125, input_sdr.c:36-38), with the
//                    virtual tuner (the reference retunes hardware, dab2eti.c:75-103)
//  fifo_read_kernel  sdr_read_fifo                   (sdr_fifo.c:43-61)
//  sync_kernel       dab_coarse_time_sync, dab_fine_time_sync, dab_coarse_freq_sync_2,
//                    dab_fine_freq_corr              (sdr_sync.c:34-302, input_sdr.c:60-112)
//  demod_kernel      76 x FFT2048 + DQPSK + frequency de-interleave + hard slicing
//                                                    (input_sdr.c:114-162)
//
// The demodulator never materialises a spectrum in memory: each CTA walks the symbols of one CIF
// (or of the FIC), keeps the previous symbol's bins in registers, and only the sliced bits leave
// the SM -- FIC as the reference's byte-per-bit array, MSC as packed planes (msc.cuh).
#include "ofdm.cuh"

#include "msc.cuh"

namespace dabgpu {

// ---- constant tables ---------------------------------------------------------------------------
__device__ float2 g_tw2048[2048];    // exp(-2*pi*i*k/2048)
__device__ float2 g_tw1536[1536];    // exp(-2*pi*i*k/1536)
__device__ uint16_t g_bin_dst[2048]; // FFT bin -> n = rev_freq_deint_tab[c] (0xffff: unused bin)
__device__ uint8_t g_prs_q[1536];    // phase reference symbol, quarter turns, by carrier index c

int ofdm_init_constants() {
  static float2 tw[2048];
  for (int k = 0; k < 2048; k++) {
    const double a = -2.0 * M_PI * (double)k / 2048.0;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  CUDA_TRY(cudaMemcpyToSymbol(g_tw2048, tw, sizeof tw));
  for (int k = 0; k < 1536; k++) {
    const double a = -2.0 * M_PI * (double)k / 1536.0;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  CUDA_TRY(cudaMemcpyToSymbol(g_tw1536, tw, 1536 * sizeof(float2)));
  uint16_t rev[1536];
  static uint16_t dst[2048];
  dabgpu_build_freq_deint(rev);
  for (int b = 0; b < 2048; b++) {
    // fftshifted index i = (b + 1024) % 2048 runs 256..1792 over the carriers, 1024 = DC
    // (input_sdr.c:152-158); carrier index c counts them in that order
    const int i = (b + 1024) & 2047;
    if (i < 256 || i > 1792 || i == 1024)
      dst[b] = 0xffff;
    else
      dst[b] = rev[i < 1024 ? i - 256 : i - 257];
  }
  CUDA_TRY(cudaMemcpyToSymbol(g_bin_dst, dst, sizeof dst));
  uint8_t q[1536];
  dabgpu_build_prs(q);
  CUDA_TRY(cudaMemcpyToSymbol(g_prs_q, q, sizeof q));
  return DABGPU_OK;
}

// uint8 sample -> the reference's int8 (input_sdr.c:61-62: buffer-127 stored in int8_t, so 255 wraps
// to -128): (b - 127) mod 256 == (b + 129) mod 256, read back as a signed byte.
__device__ __forceinline__ float u8_to_sample(uint32_t b) {
  return (float)(int)(int8_t)((b + 129u) & 0xffu);
}
// one I/Q pair (I in bits 0..7, Q in bits 8..15): add 129 to each byte in its own copy (the carry of
// the low byte must not reach the high one), then prmt sign-extends the byte of interest
__device__ __forceinline__ float2 iq_to_sample(uint32_t w) {
  const int re = (int)prmt(w + 0x81u, 0u, 0x8880u);    // byte 0, sign-replicated above
  const int im = (int)prmt(w + 0x8100u, 0u, 0x9991u);  // byte 1, sign-replicated above
  return make_float2((float)re, (float)im);
}

// =================================================================================================
// ingest
// =================================================================================================
__global__ void __launch_bounds__(256) ingest_kernel(const uint8_t *__restrict__ src, uint64_t pitch,
                                                     uint32_t chunk_len, uint8_t *__restrict__ ring,
                                                     const StepCtl *__restrict__ ctl) {
  const int s = blockIdx.y;
  const StepCtl *c = &ctl[s];
  const uint32_t vec = blockIdx.x * blockDim.x + threadIdx.x;  // 16-byte vector index
  if (vec * 16u >= chunk_len) return;
  uint4 v = *reinterpret_cast<const uint4 *>(src + (uint64_t)s * pitch + 16ull * vec);
  const int32_t df = c->nco_hz;
  if (df != 0) {
    // virtual tuner: multiply by exp(-j*2*pi*df*n/fs) and re-quantise, as the oracle harness does
    const int64_t M = 2048000;
    const int64_t a = ((int64_t)df % M + M) % M;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint64_t n = c->nco_sample0 + 8ull * vec + k;
      const int64_t r = (int64_t)(((uint64_t)a * (n % (uint64_t)M)) % (uint64_t)M);
      float sn, cs;
      sincospif(-2.0f * ((float)r / 2048000.0f), &sn, &cs);
      const uint32_t pair = (w[k >> 1] >> (16 * (k & 1))) & 0xffffu;
      const float xr = (float)(int)(pair & 0xffu) - 127.0f, xi = (float)(int)(pair >> 8) - 127.0f;
      const float yr = floorf(xr * cs - xi * sn + 0.5f) + 127.0f, yi = floorf(xr * sn + xi * cs + 0.5f) + 127.0f;
      const uint32_t qr = (uint32_t)fminf(fmaxf(yr, 0.0f), 255.0f), qi = (uint32_t)fminf(fmaxf(yi, 0.0f), 255.0f);
      const uint32_t np = qr | (qi << 8);
      w[k >> 1] = (w[k >> 1] & ~(0xffffu << (16 * (k & 1)))) | (np << (16 * (k & 1)));
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  const uint32_t pos = (c->wr_pos + 16u * vec) % IQ_RING_BYTES;
  *reinterpret_cast<uint4 *>(ring + (uint64_t)s * IQ_RING_BYTES + pos) = v;
}

int launch_ingest(const uint8_t *d_src, uint64_t src_pitch, uint32_t chunk_len, uint8_t *d_ring,
                  const StepCtl *d_ctl, int n_streams, cudaStream_t st) {
  if (n_streams <= 0 || !chunk_len) return DABGPU_OK;
  if ((src_pitch & 15) || ((uintptr_t)d_src & 15)) {
    set_error(DABGPU_ERR_ARG, "ingest: device IQ buffer and pitch must be 16-byte aligned");
    return DABGPU_ERR_ARG;
  }
  dim3 grid((chunk_len / 16 + 255) / 256, n_streams);
  ingest_kernel<<<grid, 256, 0, st>>>(d_src, src_pitch, chunk_len, d_ring, d_ctl);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// =================================================================================================
// FIFO read into the persistent per-stream frame buffer
// =================================================================================================
// 16 bytes of a ring window starting at the even offset `p` (wraps at word granularity)
__device__ __forceinline__ uint4 ring_load16(const uint8_t *rs, uint32_t mod, uint32_t p) {
  const uint32_t sh = (p & 3u) * 8u;
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; k++) w[k] = *reinterpret_cast<const uint32_t *>(rs + ring_wrap((p & ~3u) + 4u * k, mod));
  return make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh),
                    __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
}

template <bool MATERIALISE>
__global__ void __launch_bounds__(256) fifo_read_kernel(RingGeom ring, const uint8_t *__restrict__ tails,
                                                        uint8_t *__restrict__ frames,
                                                        const StepCtl *__restrict__ ctl) {
  const int s = blockIdx.y;
  const StepCtl *c = &ctl[s];
  const uint8_t *rs = ring.base + (uint64_t)s * ring.pitch;
  uint8_t *fs = frames + (uint64_t)s * DABGPU_TF_BYTES;
  const uint32_t vec = blockIdx.x * blockDim.x + threadIdx.x;
  if (MATERIALISE) {
    // the previous frame was never copied (ring mode): rebuild sdr->buffer as it stood, because a
    // short read below leaves part of it in place
    if (!c->mat) return;
    const uint32_t b = 16u * vec;
    uint4 v;
    if (b < TAIL_OFF)
      v = ring_load16(rs, ring.mod, c->mat_pos + b);
    else
      v = *reinterpret_cast<const uint4 *>(tails + (uint64_t)s * TAIL_BYTES + (b - TAIL_OFF));
    *reinterpret_cast<uint4 *>(fs + b) = v;
    return;
  }
  if (c->src_ring) return;
  // segment 0 always lands at offset 0: 16 destination bytes per thread from an even, possibly
  // unaligned ring position (all shifts are even byte counts)
  const uint32_t n0 = c->rd_bytes[0];
  if (16u * vec < n0) {
    const uint4 o4 = ring_load16(rs, ring.mod, c->rd_pos[0] + 16u * vec);
    if (16u * vec + 16u <= n0) {
      *reinterpret_cast<uint4 *>(fs + 16u * vec) = o4;
    } else {
      const uint32_t o[4] = {o4.x, o4.y, o4.z, o4.w};
      for (uint32_t b = 0; 16u * vec + b < n0; b++) fs[16u * vec + b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
    }
  }
  // segment 1 (rare: FIFO ran dry during a large positive shift): plain byte copy
  const uint32_t n1 = c->rd_bytes[1];
  for (uint32_t b = 16u * vec; b < min(n1, 16u * vec + 16u); b++)
    fs[c->rd_dst[1] + b] = rs[ring_wrap(c->rd_pos[1] + b, ring.mod)];
}

int launch_fifo_read(RingGeom d_ring, const uint8_t *d_tails, uint8_t *d_frames, const StepCtl *d_ctl,
                     int n_streams, bool materialise, cudaStream_t st) {
  if (n_streams <= 0) return DABGPU_OK;
  dim3 grid(DABGPU_TF_BYTES / 16 / 256, n_streams);
  if (materialise)
    fifo_read_kernel<true><<<grid, 256, 0, st>>>(d_ring, d_tails, d_frames, d_ctl);
  else
    fifo_read_kernel<false><<<grid, 256, 0, st>>>(d_ring, d_tails, d_frames, d_ctl);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// tail[j] = byte TAIL_OFF + j of the stream's logical frame buffer after this step's read
__global__ void __launch_bounds__(128) tail_update_kernel(RingGeom ring, const uint8_t *__restrict__ frames,
                                                          uint8_t *__restrict__ tails,
                                                          const StepCtl *__restrict__ ctl) {
  const int s = blockIdx.x;
  const StepCtl *c = &ctl[s];
  if (!c->rd_bytes[0] && !c->rd_bytes[1]) return;  // no read this step
  uint8_t *t = tails + (uint64_t)s * TAIL_BYTES;
  const uint32_t j = 16u * threadIdx.x;  // 128 threads x 16 bytes
  if (!c->src_ring) {
    *reinterpret_cast<uint4 *>(t + j) =
        *reinterpret_cast<const uint4 *>(frames + (uint64_t)s * DABGPU_TF_BYTES + TAIL_OFF + j);
    return;
  }
  const uint32_t n0 = c->rd_bytes[0];  // fresh bytes; the rest of the buffer keeps its old content
  const uint32_t b = TAIL_OFF + j;
  if (b >= n0) return;
  const uint4 v = ring_load16(ring.base + (uint64_t)s * ring.pitch, ring.mod, c->src_pos + b);
  if (b + 16u <= n0) {
    *reinterpret_cast<uint4 *>(t + j) = v;
  } else {
    const uint32_t o[4] = {v.x, v.y, v.z, v.w};
    for (uint32_t k = 0; b + k < n0; k++) t[j + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
  }
}

int launch_tail_update(RingGeom d_ring, const uint8_t *d_frames, uint8_t *d_tails, const StepCtl *d_ctl,
                       int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return DABGPU_OK;
  tail_update_kernel<<<n_streams, 128, 0, st>>>(d_ring, d_frames, d_tails, d_ctl);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// =================================================================================================
// complex helpers and in-register DFTs
// =================================================================================================
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (forward) : (x,y) -> (y,-x)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// forward 4-point DFT in place: v[k] = sum_j v[j] * exp(-2*pi*i*j*k/4)
__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = mul_mi(csub(b, d));
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = cadd(s1, s3);
  d = csub(s1, s3);
}

// forward 8-point DFT: in natural order, out natural order
__device__ __forceinline__ void dft8(float2 *v) {
  const float h = 0.70710678118654752f;
  // decimation in time: evens (0,2,4,6) and odds (1,3,5,7)
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  // twiddles W8^k: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
  o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
  o2 = mul_mi(o2);
  o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1);
  v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2);
  v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3);
  v[7] = csub(e3, o3);
}

// forward 16-point DFT, natural order in and out (4 x 4 decomposition)
__device__ __forceinline__ void dft16(float2 *v) {
  const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
  // columns: n = 4*a + b -> first DFT4 over a for each b
  float2 x[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    x[b][0] = v[b];
    x[b][1] = v[4 + b];
    x[b][2] = v[8 + b];
    x[b][3] = v[12 + b];
    dft4(x[b][0], x[b][1], x[b][2], x[b][3]);  // x[b][p], p = output index of the a-DFT
  }
  // twiddle W16^(b*p)
  // b=1: p=1: W^1, p=2: W^2, p=3: W^3 ; b=2: W^2, W^4, W^6 ; b=3: W^3, W^6, W^9
  const float2 W1 = make_float2(c1, -s1), W2 = make_float2(h, -h), W3 = make_float2(s1, -c1);
  const float2 W6 = make_float2(-h, -h), W9 = make_float2(-c1, s1);
  x[1][1] = cmul(x[1][1], W1);
  x[1][2] = cmul(x[1][2], W2);
  x[1][3] = cmul(x[1][3], W3);
  x[2][1] = cmul(x[2][1], W2);
  x[2][2] = mul_mi(x[2][2]);
  x[2][3] = cmul(x[2][3], W6);
  x[3][1] = cmul(x[3][1], W3);
  x[3][2] = cmul(x[3][2], W6);
  x[3][3] = cmul(x[3][3], W9);
  // second DFT4 over b for each p: output k = p + 4*q
#pragma unroll
  for (int p = 0; p < 4; p++) {
    dft4(x[0][p], x[1][p], x[2][p], x[3][p]);
    v[p] = x[0][p];
    v[p + 4] = x[1][p];
    v[p + 8] = x[2][p];
    v[p + 12] = x[3][p];
  }
}

// =================================================================================================
// 2048-point forward FFT by 128 threads: 16 x 16 x 8, decimation in frequency.
//   stage 1: thread t: DFT16 over x[t + 128 j]            -> k1, twiddle W2048^(t k1)
//   stage 2: thread (u = p>>4, k1 = p&15): DFT16 over j2 of z_k1[u + 8 j2] -> k2, twiddle W128^(u k2)
//   stage 3: thread p: two DFT8 over u for q = p and p+128 (q = 16 k2 + k1)  -> k3
// result: thread p holds bins p + 128 m, m = 0..15  (m = 2 k3 for q = p, 2 k3 + 1 for q = p + 128)
// `xch` is a 2080-element float2 exchange buffer.
// =================================================================================================
enum { FFT_THREADS = 128, XCH_PAD = 129, XCH_ELEMS = 16 * XCH_PAD };  // 129: both half-warps conflict-free

struct FftTwiddles {  // per-thread twiddles, constant over symbols
  float2 s1[15];      // W2048^(t*k1), k1 = 1..15
  float2 s2[15];      // W128^(u*k2),  k2 = 1..15
};
__device__ __forceinline__ void load_twiddles(FftTwiddles &tw, int p) {
  const int t = p, u = p >> 4;
#pragma unroll
  for (int k = 1; k < 16; k++) {
    tw.s1[k - 1] = g_tw2048[(t * k) & 2047];
    tw.s2[k - 1] = g_tw2048[(16 * u * k) & 2047];
  }
}

// the same twiddles fetched from the table when they are needed (two FFTs per frame in the
// synchroniser: registers matter more than 30 cached loads)
struct FftTwiddlesMem {
  int t, u16;
  struct Row {
    const FftTwiddlesMem &o;
    bool second;
    __device__ __forceinline__ float2 operator[](int km1) const {
      return g_tw2048[((second ? o.u16 : o.t) * (km1 + 1)) & 2047];
    }
  };
  Row s1{*this, false}, s2{*this, true};
  __device__ __forceinline__ explicit FftTwiddlesMem(int p) : t(p), u16(16 * (p >> 4)) {}
};

template <typename Tw>
__device__ __forceinline__ void fft2048_from_regs(float2 *v, const Tw &tw, float2 *xch, int p) {
  // stage 1 (v = x[p + 128 j])
  dft16(v);
#pragma unroll
  for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw.s1[k - 1]);
#pragma unroll
  for (int k = 0; k < 16; k++) xch[k * XCH_PAD + p] = v[k];
  __syncthreads();
  const int u = p >> 4, k1 = p & 15;
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = xch[k1 * XCH_PAD + u + 8 * j];
  __syncthreads();
  // stage 2
  dft16(v);
#pragma unroll
  for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw.s2[k - 1]);
#pragma unroll
  for (int k = 0; k < 16; k++) xch[u * 256 + k * 16 + k1] = v[k];
  __syncthreads();
  // stage 3
  float2 a[8], b[8];
#pragma unroll
  for (int uu = 0; uu < 8; uu++) {
    a[uu] = xch[uu * 256 + p];
    b[uu] = xch[uu * 256 + p + 128];
  }
  __syncthreads();
  dft8(a);
  dft8(b);
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    v[2 * k3] = a[k3];
    v[2 * k3 + 1] = b[k3];
  }
}

// load the 16 samples x[p + 128 j] of one symbol from a byte buffer (I,Q uint8 pairs)
__device__ __forceinline__ void load_symbol(float2 *v, const uint8_t *sym, int p) {
  const uint16_t *h = reinterpret_cast<const uint16_t *>(sym);
#pragma unroll
  for (int j = 0; j < 16; j++) {
    v[j] = iq_to_sample(h[p + 128 * j]);
  }
}

// =================================================================================================
// demodulator
// =================================================================================================
// TMA (1-D bulk copy) + mbarrier plumbing
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- packed single precision (sm_100 FADD2 / FMUL2 / FFMA2) -------------------------------------
// The demodulator transforms TWO consecutive OFDM symbols per thread: every real quantity of the FFT
// is a 64-bit register pair (low half: symbol A = l, high half: symbol B = l + 1).  The two
// transforms share every twiddle factor, which the packed instructions take as a broadcast scalar
// operand, so one issue slot does the work of two and nothing has to be shuffled between halves.
typedef unsigned long long pk_t;
__device__ __forceinline__ pk_t pk(float lo, float hi) {
  pk_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ pk_t pku(uint32_t lo, uint32_t hi) {
  pk_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ float pk_lo(pk_t v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a;
}
__device__ __forceinline__ float pk_hi(pk_t v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return b;
}
__device__ __forceinline__ pk_t add2(pk_t a, pk_t b) {
  pk_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t sub2(pk_t a, pk_t b) {
  pk_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t mul2(pk_t a, pk_t b) {
  pk_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk_t fma2(pk_t a, pk_t b, pk_t c) {
  pk_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ pk_t bc(float s) { return pk(s, s); }  // becomes a scalar (.F32) operand

struct C2 {  // two complex numbers: (re.lo, im.lo) of symbol A and (re.hi, im.hi) of symbol B
  pk_t re, im;
};
__device__ __forceinline__ C2 cadd2(C2 a, C2 b) { return C2{add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 csub2(C2 a, C2 b) { return C2{sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a * (wr + i wi), the same factor for both halves: 2 FMUL2 + 2 FFMA2
__device__ __forceinline__ C2 cmul2s(C2 a, float wr, float wi) {
  const pk_t t = mul2(a.re, bc(wr)), u = mul2(a.re, bc(wi));
  return C2{fma2(a.im, bc(-wi), t), fma2(a.im, bc(wr), u)};
}
// forward 4-point DFT in place; the rotation by -i is a renaming of halves folded into the adds
__device__ __forceinline__ void dft4p(C2 &a, C2 &b, C2 &c, C2 &d) {
  const C2 s0 = cadd2(a, c), s1 = csub2(a, c), s2 = cadd2(b, d), t = csub2(b, d);
  a = cadd2(s0, s2);
  c = csub2(s0, s2);
  b = C2{add2(s1.re, t.im), sub2(s1.im, t.re)};
  d = C2{sub2(s1.re, t.im), add2(s1.im, t.re)};
}
// forward 8-point DFT, natural order in and out; the 1/sqrt2 factors ride in the last butterflies
__device__ __forceinline__ void dft8p(C2 *v) {
  const float h = 0.70710678118654752f;
  C2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  C2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4p(e0, e1, e2, e3);
  dft4p(o0, o1, o2, o3);
  // W8^1 o1 = h (x + y, y - x) ; W8^2 o2 = (y, -x) ; W8^3 o3 = h (y - x, -(x + y))
  const pk_t s1 = add2(o1.re, o1.im), d1 = sub2(o1.im, o1.re);
  const pk_t s3 = add2(o3.re, o3.im), d3 = sub2(o3.im, o3.re);
  v[0] = cadd2(e0, o0);
  v[4] = csub2(e0, o0);
  v[1] = C2{fma2(s1, bc(h), e1.re), fma2(d1, bc(h), e1.im)};
  v[5] = C2{fma2(s1, bc(-h), e1.re), fma2(d1, bc(-h), e1.im)};
  v[2] = C2{add2(e2.re, o2.im), sub2(e2.im, o2.re)};
  v[6] = C2{sub2(e2.re, o2.im), add2(e2.im, o2.re)};
  v[3] = C2{fma2(d3, bc(h), e3.re), fma2(s3, bc(-h), e3.im)};
  v[7] = C2{fma2(d3, bc(-h), e3.re), fma2(s3, bc(h), e3.im)};
}
// forward 16-point DFT, natural order in and out (4 x 4)
__device__ __forceinline__ void dft16p(C2 *v) {
  const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
  C2 x[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    x[b][0] = v[b];
    x[b][1] = v[4 + b];
    x[b][2] = v[8 + b];
    x[b][3] = v[12 + b];
    dft4p(x[b][0], x[b][1], x[b][2], x[b][3]);
  }
  // twiddles W16^(b p): W1 = (c1,-s1) W2 = (h,-h) W3 = (s1,-c1) W4 = -i W6 = (-h,-h) W9 = (-c1,s1)
  x[1][1] = cmul2s(x[1][1], c1, -s1);
  x[1][3] = cmul2s(x[1][3], s1, -c1);
  x[3][1] = cmul2s(x[3][1], s1, -c1);
  x[3][3] = cmul2s(x[3][3], -c1, s1);
  {  // W2 z = h (x + y, y - x) ; W6 z = h (y - x, -(x + y))
    C2 z = x[1][2];
    x[1][2] = C2{mul2(add2(z.re, z.im), bc(h)), mul2(sub2(z.im, z.re), bc(h))};
    z = x[2][1];
    x[2][1] = C2{mul2(add2(z.re, z.im), bc(h)), mul2(sub2(z.im, z.re), bc(h))};
    z = x[2][3];
    x[2][3] = C2{mul2(sub2(z.im, z.re), bc(h)), mul2(add2(z.re, z.im), bc(-h))};
    z = x[3][2];
    x[3][2] = C2{mul2(sub2(z.im, z.re), bc(h)), mul2(add2(z.re, z.im), bc(-h))};
    z = x[2][2];
    x[2][2] = C2{z.im, sub2(bc(0.f), z.re)};  // -i z
  }
#pragma unroll
  for (int p = 0; p < 4; p++) {
    dft4p(x[0][p], x[1][p], x[2][p], x[3][p]);
    v[p] = x[0][p];
    v[p + 4] = x[1][p];
    v[p + 8] = x[2][p];
    v[p + 12] = x[3][p];
  }
}


#ifndef DABGPU_DEMOD_CTAS
#define DABGPU_DEMOD_CTAS 3  // launch bound: 3 -> up to 168 registers, 4 -> 128 (kernel experiments: build.py --variant)
#endif
enum { SYM_BYTES = 4096, STAGE_BYTES = SYM_BYTES + 16, N_STAGES = 3, DEMOD_CTAS_PER_SM = DABGPU_DEMOD_CTAS };
enum { N_SLOTS = 12 };  // carriers per thread and symbol

struct DemodSmem {
  // exchange buffer, real and imaginary parts apart: a 128-bit store would need (re, im) in four
  // consecutive registers, which costs three moves per store; two 64-bit stores cost none
  pk_t xre[XCH_ELEMS], xim[XCH_ELEMS];
  // a symbol is fetched from the 16-byte aligned address below its first byte (TMA needs aligned
  // sources; ring windows start at any even offset), hence 16 spare bytes per buffer
  __align__(16) uint8_t stage[N_STAGES][2][STAGE_BYTES];
  __align__(16) uint8_t tailbuf[TAIL_BYTES];   // second half of symbol 75 in ring mode
  __align__(16) uint8_t nib[1536];             // sliced bits of a symbol pair: A.b0 | A.b1<<1 | B.b0<<2 | B.b1<<3
  uint32_t planes[16][CIF_PLANE_WORDS];        // the CIF being assembled
  __align__(16) float2 tw2[8][16];             // W128^(u k): stage-2 twiddles (shared by 16 threads each)
  uint64_t full[N_STAGES];
};

// symbol l of a frame starts (useful part) at this byte offset (input_sdr.c:116)
__device__ __forceinline__ uint32_t sym_byte_off(int l) { return 2u * (2656u + 2552u * (uint32_t)l + 504u); }

// where a stream's frame lives this step (see StepCtl) and how one symbol is fetched from it
struct FrameSrc {
  const uint8_t *ring, *tail, *frame;
  uint32_t ring_mode, pos, delta, mod;
};

// thread 0: start the TMA copies of symbol l into `buf`; each symbol arrives once on the barrier
__device__ __forceinline__ void issue_symbol_load(DemodSmem &sm, const FrameSrc &src, int l, uint8_t *buf,
                                                  uint64_t *bar) {
  const uint32_t off = sym_byte_off(l);
  if (!src.ring_mode) {
    mbar_expect_tx(bar, SYM_BYTES);
    tma_load_1d(buf, src.frame + off, SYM_BYTES, bar);
    return;
  }
  const bool last = l == 75;  // its second half may be stale: it comes from the tail store
  const uint32_t need = (last ? SYM_BYTES - TAIL_BYTES : SYM_BYTES) + 16u;
  const uint32_t a0 = ring_wrap(src.pos + off, src.mod) & ~15u;
  mbar_expect_tx(bar, need + (last ? TAIL_BYTES : 0u));
  const uint32_t first = min(need, src.mod - a0);
  tma_load_1d(buf, src.ring + a0, first, bar);
  if (first < need) tma_load_1d(buf + first, src.ring, need - first, bar);
  if (last) tma_load_1d(sm.tailbuf, src.tail, TAIL_BYTES, bar);
}

// 16 samples x[p + 128 j] of symbols A and B -> packed floats carrying a common offset.
// The reference's sample is (int8)(b - 127) (input_sdr.c:61-62: 255 wraps to -128) = ((b + 1) & 255)
// - 128.  The byte increment is done on four bytes at once; each byte then becomes the float
// 32768 + byte by placing it in mantissa bits 8..15 under the exponent of 2^15 (one PRMT, no
// integer-to-float conversion).  The common offset 32896 (1 + i) only reaches the DC path of the
// transform, where every operation is an exact integer sum below 2^24; it is taken out again at
// (k1, k2) = (0, 0) after stage 2 (see below), so all 2048 bins are bit-identical to a transform
// of the signed samples.
__device__ __forceinline__ void load_pair(C2 *v, const uint8_t *aLo, const uint8_t *aHi, const uint8_t *bLo,
                                          const uint8_t *bHi, int p) {
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t o = 2u * (uint32_t)(p + 128 * j);
    const uint32_t ha = *reinterpret_cast<const uint16_t *>((j < 8 ? aLo : aHi) + o);
    const uint32_t hb = *reinterpret_cast<const uint16_t *>((j < 8 ? bLo : bHi) + o);
    const uint32_t w = prmt(ha, hb, 0x5410u);  // IA QA IB QB
    const uint32_t r = (((w & 0x7f7f7f7fu) + 0x01010101u) ^ (w & 0x80808080u));
    const uint32_t e = 0x47000000u;
    v[j].re = pku(prmt(r, e, 0x7404u), prmt(r, e, 0x7424u));
    v[j].im = pku(prmt(r, e, 0x7414u), prmt(r, e, 0x7434u));
  }
}
#define DEMOD_DC_OFFSET 8421376.0f  // 256 samples x (32768 + 128)

// 2048-point forward FFT of a symbol pair by 128 threads, 16 x 16 x 8 as fft2048_from_regs.
// In: v[j] = x[p + 128 j].  Out: v[m] = bin p + 128 m.  Three barriers; the caller must place one
// more between this call's last shared-memory read and the next call.
__device__ __forceinline__ void fft2048_pair(C2 *v, const float2 *tw1, DemodSmem &sm, int p) {
  pk_t *xre = sm.xre, *xim = sm.xim;
  dft16p(v);
#pragma unroll
  for (int k = 1; k < 16; k++) v[k] = cmul2s(v[k], tw1[k - 1].x, tw1[k - 1].y);
#pragma unroll
  for (int k = 0; k < 16; k++) {
    xre[k * XCH_PAD + p] = v[k].re;
    xim[k * XCH_PAD + p] = v[k].im;
  }
  __syncthreads();
  const int u = p >> 4, k1 = p & 15;
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = C2{xre[k1 * XCH_PAD + u + 8 * j], xim[k1 * XCH_PAD + u + 8 * j]};
  __syncthreads();
  dft16p(v);
  {  // remove the input offset where it has accumulated (exact: see load_pair)
    const pk_t dc = bc(k1 == 0 ? DEMOD_DC_OFFSET : 0.f);
    v[0].re = sub2(v[0].re, dc);
    v[0].im = sub2(v[0].im, dc);
  }
#pragma unroll
  for (int k = 1; k < 16; k += 2) {
    const float4 t = *reinterpret_cast<const float4 *>(&sm.tw2[u][k - 1]);  // factors k-1 (1 for k = 1) and k
    if (k > 1) v[k - 1] = cmul2s(v[k - 1], t.x, t.y);
    v[k] = cmul2s(v[k], t.z, t.w);
  }
#pragma unroll
  for (int k = 0; k < 16; k++) {
    xre[u * 256 + k * 16 + k1] = v[k].re;
    xim[u * 256 + k * 16 + k1] = v[k].im;
  }
  __syncthreads();
  C2 a[8], b[8];
#pragma unroll
  for (int uu = 0; uu < 8; uu++) {
    a[uu] = C2{xre[uu * 256 + p], xim[uu * 256 + p]};
    b[uu] = C2{xre[uu * 256 + p + 128], xim[uu * 256 + p + 128]};
  }
  dft8p(a);
  dft8p(b);
#pragma unroll
  for (int k3 = 0; k3 < 8; k3++) {
    v[2 * k3] = a[k3];
    v[2 * k3 + 1] = b[k3];
  }
}

// Grid (segments, streams): segment 0 = PRS + the three FIC symbols, segment c = 1..4 = CIF c-1
// (its 18 symbols and the one before them).  A CTA walks its symbols two at a time.
template <bool DEBUG, bool SOFT>
__global__ void __launch_bounds__(FFT_THREADS, DEMOD_CTAS_PER_SM)
    demod_kernel(RingGeom ring, const uint8_t *__restrict__ tails, const uint8_t *__restrict__ frames,
                 const StepCtl *__restrict__ ctl, const SyncOut *__restrict__ sync, uint8_t *__restrict__ fic_bits,
                 uint8_t *__restrict__ cifs, float2 *__restrict__ dbg_sym, float2 *__restrict__ dbg_symd,
                 uint8_t *__restrict__ dbg_bits, int seg_first) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  DemodSmem &sm = *reinterpret_cast<DemodSmem *>(smem_raw);
  const int p = threadIdx.x;
  const int s = blockIdx.y, seg = blockIdx.x + seg_first;
  if (!DEBUG) {
    if (!ctl[s].run || sync[s].ok != 1) return;
  }
  FrameSrc src;
  src.frame = frames + (uint64_t)s * DABGPU_TF_BYTES;
  src.ring_mode = DEBUG ? 0u : ctl[s].src_ring;
  src.pos = DEBUG ? 0u : ctl[s].src_pos;
  src.delta = src.ring_mode ? (src.pos & 15u) : 0u;
  src.ring = DEBUG ? nullptr : ring.base + (uint64_t)s * ring.pitch;
  src.mod = ring.mod;
  src.tail = DEBUG ? nullptr : tails + (uint64_t)s * TAIL_BYTES;
  const int l0 = seg == 0 ? 0 : 3 + 18 * (seg - 1);
  const int nsym = seg == 0 ? 4 : 19;
  const int npair = (nsym + 1) / 2;

  if (p == 0) {
    for (int i = 0; i < N_STAGES; i++) mbar_init(&sm.full[i], 2);
    fence_barrier_init();
  }
  sm.tw2[p >> 4][p & 15] = g_tw2048[(16 * (p >> 4) * (p & 15)) & 2047];  // W128^(u k)
  __syncthreads();
  // pair i = symbols (l0 + 2 i, l0 + 2 i + 1); the last pair of a CIF segment has no second symbol
  auto issue_pair = [&](int i, int st) {
    const int la = l0 + 2 * i;
    issue_symbol_load(sm, src, la, sm.stage[st][0], &sm.full[st]);
    if (2 * i + 1 < nsym)
      issue_symbol_load(sm, src, la + 1, sm.stage[st][1], &sm.full[st]);
    else
      mbar_expect_tx(&sm.full[st], 0);
  };
  if (p == 0) {
    for (int i = 0; i < N_STAGES && i < npair; i++) issue_pair(i, i);
  }
  float2 tw1[15];  // W2048^(p k), k = 1..15: the same for every symbol
#pragma unroll
  for (int k = 1; k < 16; k++) tw1[k - 1] = g_tw2048[(p * k) & 2047];
  // Thread p ends every FFT with bins p + 128 m.  Carriers are bins 1..768 and 1280..2047: m = 0..5
  // and 10..15 always (except bin 0 = DC for p = 0) and m = 6 for p = 0 only (bin 768), which takes
  // the place of thread 0's DC bin: slot k <-> m = k (k < 6) or k + 4.  A carrier's four bits of a
  // symbol pair go to one byte of sm.nib whose position is constant over symbols: plane order for
  // the MSC (msc.cuh), n itself for the FIC symbols.
  const bool plane_order = !(seg == 0 || DEBUG || SOFT);
  uint32_t pos[N_SLOTS];
#pragma unroll
  for (int k = 0; k < N_SLOTS; k++) {
    const int m = k < 6 ? k : k + 4;
    const uint32_t n = g_bin_dst[(k == 0 && p == 0) ? 768 : p + 128 * m];
    pos[k] = plane_order ? (n & 15u) * 96u + (n >> 4) : n;
  }
  float px[N_SLOTS], py[N_SLOTS];  // the previous symbol's carriers
#pragma unroll
  for (int k = 0; k < N_SLOTS; k++) px[k] = py[k] = 0.f;

  for (int i = 0; i < npair; i++) {
    const int la = l0 + 2 * i, st = i % N_STAGES;
    const bool b_valid = 2 * i + 1 < nsym;
    C2 v[16];
    mbar_wait(&sm.full[st], (uint32_t)(i / N_STAGES) & 1u);
    {
      const uint8_t *a = sm.stage[st][0] + src.delta;
      const uint8_t *b = b_valid ? sm.stage[st][1] + src.delta : a;
      const uint8_t *ah = a, *bh = b;
      if (src.ring_mode) {  // symbol 75: samples 1024.. come from the tail store
        if (la == 75) ah = sm.tailbuf - 2048;
        if (la + 1 == 75) bh = sm.tailbuf - 2048;
      }
      load_pair(v, a, ah, b, bh, p);
    }
    fft2048_pair(v, tw1, sm, p);  // (its first barrier also says: staging buffer consumed)
    if (DEBUG) {
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const int bin = (p + 128 * m + 1024) & 2047;
        dbg_sym[(size_t)la * 2048 + bin] = make_float2(pk_lo(v[m].re), pk_lo(v[m].im));
        if (b_valid) dbg_sym[(size_t)(la + 1) * 2048 + bin] = make_float2(pk_hi(v[m].re), pk_hi(v[m].im));
      }
    }
    // DQPSK against the previous symbol and hard slicing (input_sdr.c:132-158):
    //   re = Re(s_l conj(s_l-1)),  im' = -Im(s_l conj(s_l-1))   (the reference divides both by
    //   |s_l-1|^2 > 0, which cannot change a sign);  bit0 = !(re > 0),  bit1 = (im' > 0).
    // Both are read off sign bits: r = re - tiny is negative exactly when !(re > 0), q = tiny - im'
    // exactly when im' > 0 (tiny = 1e-30 is far below the rounding unit of any non-zero product of
    // two spectrum values, so it only decides the case of an exact zero, as `> 0` does).
    const float tiny = 1e-30f;
#pragma unroll
    for (int k = 0; k < N_SLOTS; k++) {
      const int m = k < 6 ? k : k + 4;
      C2 c = v[m];
      if (k == 0) {
        if (p == 0) c = v[6];
      }
      const float xa = pk_lo(c.re), ya = pk_lo(c.im), xb = pk_hi(c.re), yb = pk_hi(c.im);
      const float ra = fmaf(xa, px[k], fmaf(ya, py[k], -tiny));
      const float qa = fmaf(ya, px[k], fmaf(-xa, py[k], tiny));
      const float rb = fmaf(xb, xa, fmaf(yb, ya, -tiny));
      const float qb = fmaf(yb, xa, fmaf(-xb, ya, tiny));
      if (DEBUG) {
        const int bin = (((k == 0 && p == 0) ? 768 : p + 128 * m) + 1024) & 2047;
        {
          const float da = px[k] * px[k] + py[k] * py[k], db = xa * xa + ya * ya;
          if (i > 0)
            dbg_symd[(size_t)la * 2048 + bin] =
                make_float2((xa * px[k] + ya * py[k]) / da, (xa * py[k] - ya * px[k]) / da);
          if (b_valid)
            dbg_symd[(size_t)(la + 1) * 2048 + bin] = make_float2((xb * xa + yb * ya) / db, (xb * ya - yb * xa) / db);
        }
      }
      if (SOFT) {
        // Soft decisions (opt-in): instead of the signs, the normalised differential product itself
        // (what input_sdr.c:139-142 computes and :157-158 slices), as a symbol 128 -+ round(8 x)
        // clipped to 121..135: < 128 says "bit 0", like the reference's 127 / 129.
        uint8_t *sb = reinterpret_cast<uint8_t *>(sm.planes);  // [4][1536]: A.b0, A.b1, B.b0, B.b1 by n
        const float da = px[k] * px[k] + py[k] * py[k], db = xa * xa + ya * ya;
        const float ga = da > 0.f ? 8.f / da : 0.f, gb = db > 0.f ? 8.f / db : 0.f;
        const float ea = xa * px[k] + ya * py[k], fa = xa * py[k] - ya * px[k];
        const float eb = xb * xa + yb * ya, fb = xb * ya - yb * xa;
        const uint32_t n = pos[k];
        sb[n] = (uint8_t)(128 - (int)fminf(fmaxf(rintf(ea * ga), -7.f), 7.f));
        sb[1536 + n] = (uint8_t)(128 + (int)fminf(fmaxf(rintf(fa * ga), -7.f), 7.f));
        sb[3072 + n] = (uint8_t)(128 - (int)fminf(fmaxf(rintf(eb * gb), -7.f), 7.f));
        sb[4608 + n] = (uint8_t)(128 + (int)fminf(fmaxf(rintf(fb * gb), -7.f), 7.f));
      } else {
        uint32_t nb = __funnelshift_l(__float_as_uint(qb), 0u, 1);
        nb = __funnelshift_l(__float_as_uint(rb), nb, 1);
        nb = __funnelshift_l(__float_as_uint(qa), nb, 1);
        nb = __funnelshift_l(__float_as_uint(ra), nb, 1);
        sm.nib[pos[k]] = (uint8_t)nb;
      }
      px[k] = xb;
      py[k] = yb;
    }
    __syncthreads();
    // Refill the staging buffers this pair was read from (every thread is past the FFT's barriers).
    // Issued from the fourth warp, which has no part in the bit packing below, so that the ~100
    // instructions of address arithmetic and copy issue do not sit on the packing warps' path.
    if (p == 96 && i + N_STAGES < npair) {
      fence_proxy_async();
      issue_pair(i + N_STAGES, st);
    }
    // rows: symbol A is data symbol la - 1 of the frame, B is la
    if (SOFT) {
      // symbols in logical order: FIC as 3 rows of 3072 per stream, a CIF as 18 rows in its slot of the
      // soft CIF store (55296 bytes; the hard store's offsets address it slot by slot)
      uint8_t *out = seg == 0 ? fic_bits + (uint64_t)s * 9216
                              : cifs + ctl[s].cif_off[seg - 1] / CIF_BYTES * 55296ull;
      const int row_a = seg == 0 ? la - 1 : 2 * i - 1;
      const uint4 *sb = reinterpret_cast<const uint4 *>(sm.planes);
      for (int c = p; c < 384; c += FFT_THREADS) {  // 4 kinds x 96 vectors of 16 bytes
        const int kind = c / 96, v = c % 96;
        const bool is_b = kind >= 2;
        if (is_b ? !b_valid : i == 0) continue;
        const int row = row_a + (is_b ? 1 : 0);
        reinterpret_cast<uint4 *>(out + (size_t)row * 3072 + (kind & 1) * 1536)[v] = sb[c];
      }
    } else if (!plane_order) {
      uint8_t *out = DEBUG ? dbg_bits : fic_bits + (uint64_t)s * 9216;
      const bool a_valid = i > 0;  // the first symbol of a segment is only the phase reference
      for (int c = p; c < 384; c += FFT_THREADS) {
        const uint32_t w = reinterpret_cast<const uint32_t *>(sm.nib)[c];
        if (a_valid) {
          uint32_t *o = reinterpret_cast<uint32_t *>(out + (size_t)(la - 1) * 3072);
          o[c] = w & 0x01010101u;
          o[384 + c] = (w >> 1) & 0x01010101u;
        }
        if (b_valid) {
          uint32_t *o = reinterpret_cast<uint32_t *>(out + (size_t)la * 3072);
          o[c] = (w >> 2) & 0x01010101u;
          o[384 + c] = (w >> 3) & 0x01010101u;
        }
      }
    } else if (p < 96) {
      // 32 carriers' nibbles -> the bit-0 and the bit-1 word of one symbol: plane m, word wq of
      // the symbol's 96 positions.  (x & 0x01010101 << k) * (0x10204080 >> k) gathers bit k of
      // four bytes into bits 28..31.
      const int m = p / 6, r = p % 6, wq = r % 3, ab = r / 3;
      const uint4 *q = reinterpret_cast<const uint4 *>(sm.nib + m * 96 + 32 * wq);
      const uint4 q0 = q[0], q1 = q[1];
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const uint32_t k0 = 2u * (uint32_t)ab;
      const uint32_t mask0 = 0x01010101u << k0, mul0 = 0x10204080u >> k0;
      const uint32_t mask1 = mask0 << 1, mul1 = mul0 >> 1;
      uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
      for (int j = 7; j >= 0; j--) {
        acc0 = __funnelshift_l((w[j] & mask0) * mul0, acc0, 4);
        acc1 = __funnelshift_l((w[j] & mask1) * mul1, acc1, 4);
      }
      const int d = 2 * i - 1 + ab;  // data symbol of the CIF: A = 2 i - 1, B = 2 i
      if (d >= 0 && d < 18) {
        sm.planes[m][d * 6 + wq] = acc0;
        sm.planes[m][d * 6 + 3 + wq] = acc1;
      }
    }
    // the next pair's slicing must not overwrite sm.nib before it has been read, and its first
    // exchange must not overwrite sm.xch before this pair's last read: the barrier above and the
    // barriers inside the next fft2048_pair order both
  }
  if (plane_order) {
    __syncthreads();
    uint32_t *o = reinterpret_cast<uint32_t *>(cifs + ctl[s].cif_off[seg - 1]);
    const uint32_t *pl = &sm.planes[0][0];
    for (int k = p; k < CIF_WORDS; k += FFT_THREADS) o[k] = pl[k];
  }
}

int launch_demod(RingGeom d_ring, const uint8_t *d_tails, const uint8_t *d_frames, const StepCtl *d_ctl,
                 const SyncOut *d_sync, uint8_t *d_fic_bits, uint8_t *d_cifs, int n_streams, int seg_first,
                 int seg_count, bool soft, cudaStream_t st) {
  if (n_streams <= 0) return DABGPU_OK;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(demod_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(DemodSmem)));
    CUDA_TRY(cudaFuncSetAttribute(demod_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(DemodSmem)));
    attr_set = true;
  }
  dim3 grid(seg_count, n_streams);
  if (soft)
    demod_kernel<false, true><<<grid, FFT_THREADS, sizeof(DemodSmem), st>>>(d_ring, d_tails, d_frames, d_ctl, d_sync,
                                                                            d_fic_bits, d_cifs, nullptr, nullptr,
                                                                            nullptr, seg_first);
  else
    demod_kernel<false, false><<<grid, FFT_THREADS, sizeof(DemodSmem), st>>>(d_ring, d_tails, d_frames, d_ctl, d_sync,
                                                                             d_fic_bits, d_cifs, nullptr, nullptr,
                                                                             nullptr, seg_first);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

int launch_demod_debug(const uint8_t *d_frame, float2 *d_symbols, float2 *d_symbols_d, uint8_t *d_bits,
                       cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(demod_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(DemodSmem)));
    attr_set = true;
  }
  // one "segment" walking all 76 symbols is what the debug variant needs: reuse seg 0 semantics
  // by launching the five segments; each writes its own rows
  dim3 grid(5, 1);
  demod_kernel<true, false><<<grid, FFT_THREADS, sizeof(DemodSmem), st>>>(RingGeom{nullptr, 0, IQ_RING_BYTES}, nullptr, d_frame, nullptr, nullptr,
                                                                  nullptr, nullptr, d_symbols, d_symbols_d, d_bits, 0);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// =================================================================================================
// synchronisers: one CTA of 128 threads per stream
// =================================================================================================
// Three forward 512-point FFTs (transform r at d + FFT512_PITCH * r, natural order in and out, in
// place) by the whole CTA, in registers: 512 = 8 x 8 x 8, decimation in frequency, 64 "virtual
// threads" of 8 points per transform, so 192 of them over 128 threads (vt = p and p + 128).
//   stage 1: vt (r, t):          DFT8 over y[t + 64 j]               -> k1, twiddle W512^(t k1)
//   stage 2: vt (r, k1, u):      DFT8 over z_k1[u + 8 j2]            -> k2, twiddle W64^(u k2)
//   stage 3: vt (r, q = k1 + 8 k2): DFT8 over u                      -> bin q + 64 k3
// Every stage reads all its inputs, then (barrier) writes its outputs into the same region; the row
// pitches 72 and 66 make the 64-bit accesses of both sides of each exchange conflict-free.
enum { FFT512_PITCH = 576 };
__device__ void block_fft512x3(float2 *d) {
  const int p = threadIdx.x;
  const bool two = p < 64;  // threads 0..63 also run virtual threads 128..191 (transform 2)
  float2 *d0 = d + FFT512_PITCH * (p >> 6), *d1 = d + FFT512_PITCH * 2;
  const int t = p & 63;
  float2 a[8], b[8];
  // stage 1
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = d0[t + 64 * j];
    if (two) b[j] = d1[t + 64 * j];
  }
  __syncthreads();
  dft8(a);
  if (two) dft8(b);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float2 w = g_tw2048[4 * t * k];
    d0[k * 72 + t] = k ? cmul(a[k], w) : a[k];
    if (two) d1[k * 72 + t] = k ? cmul(b[k], w) : b[k];
  }
  __syncthreads();
  // stage 2
  const int k1 = t >> 3, u = t & 7;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = d0[k1 * 72 + u + 8 * j];
    if (two) b[j] = d1[k1 * 72 + u + 8 * j];
  }
  __syncthreads();
  dft8(a);
  if (two) dft8(b);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float2 w = g_tw2048[32 * u * k];
    d0[u * 66 + k1 + 8 * k] = k ? cmul(a[k], w) : a[k];
    if (two) d1[u * 66 + k1 + 8 * k] = k ? cmul(b[k], w) : b[k];
  }
  __syncthreads();
  // stage 3
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = d0[j * 66 + t];
    if (two) b[j] = d1[j * 66 + t];
  }
  __syncthreads();
  dft8(a);
  if (two) dft8(b);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    d0[t + 64 * k] = a[k];
    if (two) d1[t + 64 * k] = b[k];
  }
  __syncthreads();
}

// Largest magnitude of the 128-point DFT of x[l + 8 j] (j = 0..15 in v) by a group of 8 consecutive
// lanes l, in registers: 128 = 16 x 8, decimation in frequency with one exchange through `buf` (136
// float2 per group, groups of a warp 136 apart: conflict-free on both sides).  Warp-synchronous.
__device__ __forceinline__ float group8_fft128_peak(float2 *v, float2 *buf, int l) {
  dft16(v);
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++) {
    const float2 w = g_tw2048[16 * l * n1];
    buf[(n1 >> 1) * 17 + (n1 & 1) * 8 + l] = n1 ? cmul(v[n1], w) : v[n1];
  }
  __syncwarp();
  float2 a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    a[j] = buf[l * 17 + j];
    b[j] = buf[l * 17 + 8 + j];
  }
  __syncwarp();
  dft8(a);
  dft8(b);
  float m = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    m = fmaxf(m, a[j].x * a[j].x + a[j].y * a[j].y);
    m = fmaxf(m, b[j].x * b[j].x + b[j].y * b[j].y);
  }
#pragma unroll
  for (int o = 4; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  return sqrtf(m);  // sqrt is monotonic: the largest magnitude is the root of the largest square
}

#ifndef SYNC_CTAS_PER_SM
#define SYNC_CTAS_PER_SM 7   // 1024 streams in one wave on 148 SMs; needs <= 72 registers and < 31.5 KB
#endif
enum { SPEC_GAP0 = 784, SPEC_GAP1 = 1280, SPEC_SLOTS = 2048 - (SPEC_GAP1 - SPEC_GAP0) };
// slot of FFT bin `bin` in SyncSmem::spec, -1 for the unused bins between the two carrier blocks
__device__ __forceinline__ int spec_slot(int bin) {
  return bin < SPEC_GAP0 ? bin : (bin >= SPEC_GAP1 ? bin - (SPEC_GAP1 - SPEC_GAP0) : -1);
}
struct SyncSmem {
  union {              // never live at the same time:
    float2 xch[XCH_ELEMS];  // FFT exchange buffer (inside fft_window)
    float2 work[16 * 136];  // correlation buffers (between the FFTs): 3 x FFT512_PITCH or 16 groups x 136
  };
  // FFT output by bin, without the bins nobody reads: the carriers are bins 1..768 and 1280..2047
  // (fine time sync reads 3..770 and 1280..2047, the coarse frequency search 1280..1435)
  float2 spec[SPEC_SLOTS];
  float red_v[FFT_THREADS];
  int red_i[FFT_THREADS];
};

__device__ __forceinline__ float2 prs_value(int c) {
  const int q = g_prs_q[c];
  return make_float2(q == 0 ? 1.f : q == 2 ? -1.f : 0.f, q == 1 ? 1.f : q == 3 ? -1.f : 0.f);
}

// sample sources: the uint8 frame buffer (batched path, sdr_demod) or caller-provided arrays
// (the reference-signature entry points of sdr_sync.h)
struct SrcU8 {
  const uint8_t *f;
  __device__ __forceinline__ float real(int n) const { return u8_to_sample(f[2 * n]); }
  __device__ __forceinline__ float2 at(int n) const {
    return iq_to_sample(reinterpret_cast<const uint16_t *>(f)[n]);
  }
};
// the frame through the ring window + tail store (StepCtl.src_ring == 1)
struct SrcRing {
  const uint8_t *ring, *tail;
  uint32_t pos, mod;
  __device__ __forceinline__ const uint8_t *byte_ptr(uint32_t b) const {
    return b < TAIL_OFF ? ring + ring_wrap(pos + b, mod) : tail + (b - TAIL_OFF);
  }
  __device__ __forceinline__ float real(int n) const { return u8_to_sample(*byte_ptr(2u * (uint32_t)n)); }
  __device__ __forceinline__ float2 at(int n) const {
    return iq_to_sample(*reinterpret_cast<const uint16_t *>(byte_ptr(2u * (uint32_t)n)));
  }
};
struct SrcI8 {
  const int8_t *r;
  __device__ __forceinline__ float real(int n) const { return (float)r[n]; }
  __device__ __forceinline__ float2 at(int) const { return make_float2(0.f, 0.f); }
};
struct SrcF2 {
  const float2 *f;
  __device__ __forceinline__ float real(int n) const { return f[n].x; }
  __device__ __forceinline__ float2 at(int n) const { return f[n]; }
};

__device__ float block_sum(SyncSmem &sm, float v) {
  const int p = threadIdx.x;
  sm.red_v[p] = v;
  __syncthreads();
  for (int o = FFT_THREADS / 2; o; o >>= 1) {
    if (p < o) sm.red_v[p] += sm.red_v[p + o];
    __syncthreads();
  }
  const float r = sm.red_v[0];
  __syncthreads();
  return r;
}

// block-wide arg-extremum with lowest-index tie-break; is_max selects max or min
__device__ void block_arg_reduce(SyncSmem &sm, float v, int idx, bool is_max, float *out_v, int *out_i) {
  sm.red_v[threadIdx.x] = v;
  sm.red_i[threadIdx.x] = idx;
  __syncthreads();
  for (int o = FFT_THREADS / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const float a = sm.red_v[threadIdx.x], b = sm.red_v[threadIdx.x + o];
      const int ia = sm.red_i[threadIdx.x], ib = sm.red_i[threadIdx.x + o];
      const bool take_b = is_max ? (b > a || (b == a && ib < ia)) : (b < a || (b == a && ib < ia));
      if (take_b) {
        sm.red_v[threadIdx.x] = b;
        sm.red_i[threadIdx.x] = ib;
      }
    }
    __syncthreads();
  }
  *out_v = sm.red_v[0];
  *out_i = sm.red_i[0];
  __syncthreads();
}

// sdr_sync.c:34-68: null-symbol energy gate; on a miss the first minimum of a 266-tap sliding
// sum over |real[10 n]| locates the null symbol.  Returns the shift in bytes.
template <typename Src>
__device__ int coarse_time_sync(SyncSmem &sm, const Src &src, bool force, float *energy) {
  const int p = threadIdx.x;
  float e = 0.f;
  for (int k = p; k < 266; k += FFT_THREADS) e += fabsf(src.real(10 * k));
  const float ev = block_sum(sm, e);  // small integers: exact in float in any order
  *energy = ev;
  if (ev < 5000.f && !force) return 0;
  const int nwin = (196608 - 2656) / 10;  // the reference searches 19395 windows
  const int per = (nwin + FFT_THREADS - 1) / FFT_THREADS;
  const int w0 = p * per, w1 = min(nwin, w0 + per);
  float best = 9999999.f;
  int best_w = 0x7fffffff;
  if (w0 < w1) {
    float sum = 0.f;
    for (int j = 0; j < 266; j++) sum += fabsf(src.real(10 * (w0 + j)));
    for (int w = w0; w < w1; w++) {
      if (sum < best) {
        best = sum;
        best_w = w;
      }
      sum += fabsf(src.real(10 * (w + 266))) - fabsf(src.real(10 * w));
    }
  }
  float bv;
  int bi;
  block_arg_reduce(sm, best, best_w, false, &bv, &bi);
  return bv < 9999999.f ? 20 * bi : 0;
}

// FFT of the 2048 samples starting at `start` into sm.spec (natural bin order)
template <typename Src, typename Tw>
__device__ void fft_window(SyncSmem &sm, const Src &src, int start, const Tw &tw) {
  const int p = threadIdx.x;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = src.at(start + p + 128 * j);
  fft2048_from_regs(v, tw, sm.xch, p);
#pragma unroll
  for (int m = 0; m < 16; m++) {
    const int slot = spec_slot(p + 128 * m);
    if (slot >= 0) sm.spec[slot] = v[m];
  }
  __syncthreads();
}

// sdr_sync.c:71-202 on the PRS spectrum in sm.spec: correlate 1536 carriers with conj(PRS),
// 1536-point inverse DFT, first maximum of the magnitude.  Only magnitudes are used, and
// |IDFT(c)| = |DFT(conj c)|: the forward transform of conj(c) = conj(spec) * PRS is computed, as
// three 512-point transforms of the samples i = 3 q + r and a radix-3 combine.
__device__ int fine_time_from_spec(SyncSmem &sm) {
  const int p = threadIdx.x;
  for (int i = p; i < 1536; i += FFT_THREADS) {
    const int bin = i < 768 ? i + 1280 : i - 765;  // sic: off by two in the upper half
    const float2 c = cmulc(sm.spec[spec_slot(bin)], prs_value(i));
    sm.work[(i % 3) * FFT512_PITCH + i / 3] = make_float2(c.x, -c.y);
  }
  __syncthreads();
  block_fft512x3(sm.work);
  float best = -99999.f;
  int best_i = 0x7fffffff;
  for (int k = p; k < 1536; k += FFT_THREADS) {
    const int km = k & 511;
    float2 acc = sm.work[km];
    acc = cadd(acc, cmul(sm.work[FFT512_PITCH + km], g_tw1536[k]));
    acc = cadd(acc, cmul(sm.work[2 * FFT512_PITCH + km], g_tw1536[2 * k < 1536 ? 2 * k : 2 * k - 1536]));
    const float mag = sqrtf(acc.x * acc.x + acc.y * acc.y);
    if (mag > best) {  // ascending k per thread: the first maximum is kept
      best = mag;
      best_i = k;
    }
  }
  float bv;
  int bi;
  block_arg_reduce(sm, best, best_i, true, &bv, &bi);
  return bi < 768 ? 2 * bi + 16 : 2 * (bi - 1536);
}

// sdr_sync.c:205-258 on the spectrum in sm.spec (natural order; the reference's fftshifted
// index i is bin (i + 1024) mod 2048).  The 29 hypotheses are independent: a group of 8 lanes takes
// one (16 at a time), holds its 128 correlation products in registers and only needs the largest
// magnitude of their inverse DFT, which is that of the forward DFT of the conjugates.  The reference
// keeps the first k whose peak is strictly larger, i.e. the largest peak with the smallest k on ties.
__device__ int coarse_freq_from_spec(SyncSmem &sm) {
  const int p = threadIdx.x, l = p & 7, g = p >> 3;
  float2 *buf = sm.work + 136 * g;
  float best = -99999.f;
  int best_k = 99;
  for (int round = 0; round < 2; round++) {  // a group's hypotheses in ascending order; all lanes take
    const int kk = -14 + g + 16 * round;     // part in both rounds (warp-wide exchange and shuffles)
    const bool valid = kk <= 14;
    const int k = valid ? kk : 14;
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int s = l + 8 * j;
      const float2 c = cmulc(sm.spec[spec_slot((14 + k + 256 + s + 1024) & 2047)], prs_value(14 + s));
      v[j] = make_float2(c.x, -c.y);
    }
    const float mag = group8_fft128_peak(v, buf, l);
    if (valid && mag > best) {
      best = mag;
      best_k = k;
    }
  }
  float bv;
  int bi;
  block_arg_reduce(sm, best, best_k + 14, true, &bv, &bi);  // ties -> lowest k
  return bi - 14;
}

// sdr_sync.c:259-302: mean phase of x[n+2048] conj(x[n]) over the PRS guard interval, in Hz.
// In double like the reference: the products of 8-bit samples are exact and atan2 is taken in double,
// because the tuner feedback truncates `frequency + ffs / 3` to an integer (dab2eti.c:98-101) and a
// float-level difference there would move a borderline stream by 1 Hz for the rest of its life.  The
// 504 angles are added in a fixed tree order (per thread ascending i, then lanes, then warps) rather
// than the reference's sequential one: the sums differ by a few units of 1e-16 relative, far below
// anything the truncation can see, and a 504-long dependent chain in one thread was 12 % of the kernel.
template <typename Src>
__device__ double fine_freq(SyncSmem &sm, const Src &src) {
  double *part = reinterpret_cast<double *>(sm.work);  // the correlation buffers are free here
  double acc = 0.0;
  for (int i = threadIdx.x; i < 504; i += FFT_THREADS) {
    const float2 l = src.at(2656 + 2048 + i), r = src.at(2656 + i);
    const double lx = l.x, ly = l.y, rx = r.x, ry = r.y;
    acc += atan2(ly * rx - lx * ry, lx * rx + ly * ry);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  const double mean = (part[0] + part[1]) + (part[2] + part[3]);
  __syncthreads();
  return mean / 504 / (2 * 3.14159265358979323846) * 1000;
}

// the synchroniser half of sdr_demod (input_sdr.c:65-112) on one frame
template <typename Src>
__device__ void sync_frame(SyncSmem &sm, const Src &src, bool force, SyncOut &r) {
  r.ok = 0;
  r.coarse_freq_shift = 0;
  r.stage = 1;
  r.coarse_timeshift = coarse_time_sync(sm, src, force, &r.null_energy);
  if (r.coarse_timeshift != 0) return;
  const FftTwiddlesMem tw(threadIdx.x);
  fft_window(sm, src, 2656 + 504, tw);
  r.fine_timeshift = fine_time_from_spec(sm);
  // input_sdr.c:91: the reference indexes the frame with the *byte* shift here
  fft_window(sm, src, 2656 + 505 + r.fine_timeshift, tw);
  r.coarse_freq_shift = coarse_freq_from_spec(sm);
  r.stage = 2;
  if (abs(r.coarse_freq_shift) > 1) return;
  r.fine_freq_shift = fine_freq(sm, src);
  r.ok = 1;
  r.stage = 3;
}

// ... for every stream with ctl.run
__global__ void __launch_bounds__(FFT_THREADS, SYNC_CTAS_PER_SM) sync_kernel(RingGeom ring, const uint8_t *__restrict__ tails,
                                                           const uint8_t *__restrict__ frames,
                                                           const StepCtl *__restrict__ ctl,
                                                           SyncOut *__restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SyncSmem &sm = *reinterpret_cast<SyncSmem *>(smem_raw);
  const int s = blockIdx.x;
  if (!ctl[s].run) return;
  SyncOut r = out[s];  // fine_timeshift / fine_freq_shift persist across early exits (sdr_state_t)
  const bool force = ctl[s].force_timesync != 0;
  if (ctl[s].src_ring)
    sync_frame(sm, SrcRing{ring.base + (uint64_t)s * ring.pitch, tails + (uint64_t)s * TAIL_BYTES, ctl[s].src_pos,
                           ring.mod},
               force, r);
  else
    sync_frame(sm, SrcU8{frames + (uint64_t)s * DABGPU_TF_BYTES}, force, r);
  if (threadIdx.x == 0) out[s] = r;
}

int launch_sync(RingGeom d_ring, const uint8_t *d_tails, const uint8_t *d_frames, const StepCtl *d_ctl,
                SyncOut *d_out, int n_streams, cudaStream_t st) {
  if (n_streams <= 0) return DABGPU_OK;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SyncSmem)));
    attr_set = true;
  }
  sync_kernel<<<n_streams, FFT_THREADS, sizeof(SyncSmem), st>>>(d_ring, d_tails, d_frames, d_ctl, d_out);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

// single-function variants behind the reference's sdr_sync.h entry points
//   mode 0: dab_coarse_time_sync(int8 real[196608])      -> res[0] (bytes)
//   mode 1: dab_fine_time_sync(frame as float2[196608])  -> res[0] (bytes)
//   mode 2: dab_coarse_freq_sync_2(shifted spectrum float2[2048]) -> res[0] (carriers)
//   mode 3: dab_fine_freq_corr(frame as float2[>= 5208]) -> fres[0] (Hz)
__global__ void __launch_bounds__(FFT_THREADS) sync_single_kernel(int mode, const void *in, int force, int *res,
                                                                  double *fres) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SyncSmem &sm = *reinterpret_cast<SyncSmem *>(smem_raw);
  const int p = threadIdx.x;
  int r = 0;
  double fr = 0.0;
  if (mode == 0) {
    float e;
    r = coarse_time_sync(sm, SrcI8{(const int8_t *)in}, force != 0, &e);
    fr = e;
  } else if (mode == 1) {
    FftTwiddles tw;
    load_twiddles(tw, p);
    fft_window(sm, SrcF2{(const float2 *)in}, 2656 + 504, tw);
    r = fine_time_from_spec(sm);
  } else if (mode == 2) {
    const float2 *sh = (const float2 *)in;
    for (int i = p; i < 2048; i += FFT_THREADS) {
      const int slot = spec_slot((i + 1024) & 2047);
      if (slot >= 0) sm.spec[slot] = sh[i];
    }
    __syncthreads();
    r = coarse_freq_from_spec(sm);
  } else {
    fr = fine_freq(sm, SrcF2{(const float2 *)in});
  }
  if (p == 0) {
    res[0] = r;
    fres[0] = fr;
  }
}

int launch_sync_single(int mode, const void *d_in, int force, int *d_res, double *d_fres, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(sync_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(SyncSmem)));
    attr_set = true;
  }
  sync_single_kernel<<<1, FFT_THREADS, sizeof(SyncSmem), st>>>(mode, d_in, force, d_res, d_fres);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

}  // namespace dabgpu
