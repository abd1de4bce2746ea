// viterbi.cuh -- batched K=7 rate-1/4 Viterbi decoder for DAB (sm_100a).
//
// Replaces the reference's viterbi() (src/viterbi.c:352-452) plus the energy-dispersal
// descrambler that always follows it (src/misc.c:41-58), bit-exactly, for the symbol
// alphabet the reference's depuncturers emit ({127,128,129} = 0 / erasure / 1).
//
// Input format ("step bytes"): one byte per trellis step,
//     bits 0..3 = received hard bit for generator j (0 if punctured)
//     bits 4..7 = 1 where generator j's symbol was transmitted (0 = erasure)
// one row of step bytes per codeword, rows 16-byte aligned and padded to 16 bytes.
#pragma once
#include "common.cuh"

namespace dabgpu {

struct VitJob {
  uint64_t in_off;   // byte offset of this codeword's step-byte row
  uint64_t out_off;  // byte offset of the decoded bytes (4-byte aligned)
  uint32_t nbits;    // information bits (trellis steps = nbits + 6)
  uint32_t flags;    // VIT_DESCRAMBLE
};
enum : uint32_t { VIT_DESCRAMBLE = 1u };

// One warp decodes up to 32 codewords of identical length, one per lane.
struct VitGroup {
  uint32_t job0;     // index of lane 0's job; lanes use job0 + lane
  uint32_t nlanes;   // active lanes (<= 32)
  uint32_t nsteps;   // nbits + 6, identical for all lanes of the group
  uint32_t pad;
  uint64_t dec_off;  // offset into the decision scratch, in uint2 units
};

// decision scratch needed by a group: 32 uint2 per step, steps padded to whole 16-step chunks
__host__ __device__ static inline uint64_t vit_group_dec_words(uint32_t nsteps) {
  return (uint64_t)((nsteps + 15u) & ~15u) * 32u;
}
__host__ __device__ static inline uint32_t vit_row_bytes(uint32_t nsteps) { return (nsteps + 15u) & ~15u; }

// Persistent launch shape: VIT_CTAS_PER_SM CTAs of VIT_WARPS warps on every SM, each warp with its
// own work list.  Measured on B200 (1024 streams, ETI frames/s of the whole receiver, bench.py):
//   1 x  8 warps  3.98 M      3 x 4 warps  3.94 M      1 x 12 / 1 x 16 warps  3.82 M
//   2 x  8 warps  3.78 M      2 x 4 warps  3.50 M (nothing stops three of its CTAs from landing on one SM)
// Two warps per scheduler are enough to keep the integer pipe busy, and one 8-warp CTA (35 K
// registers) leaves room for a demodulator CTA of the following frames (21.5 K) on every SM, whose
// FP32 and shared-memory work then runs beside the decoder's integer work.  (With round 1's
// 128-register demodulator the best shape was 3 x 4.)
#ifndef DABGPU_VIT_WARPS
#define DABGPU_VIT_WARPS 8
#endif
#ifndef DABGPU_VIT_CTAS_PER_SM
#define DABGPU_VIT_CTAS_PER_SM 1
#endif
enum { VIT_WARPS = DABGPU_VIT_WARPS, VIT_CTAS_PER_SM = DABGPU_VIT_CTAS_PER_SM };
int device_sm_count();
// persistent launch: n_ctas CTAs of VIT_WARPS warps; warp-bin b owns groups
// d_bin_start[b] .. d_bin_start[b+1] (n_ctas * VIT_WARPS + 1 entries)
// warps_per_cta: VIT_WARPS (persistent, VIT_CTAS_PER_SM CTAs per SM) or 1 (one small CTA per work list: a
// footprint that can share SMs with other kernels, used for the latency-critical FIC batches)
int launch_viterbi(const uint8_t *d_steps, uint8_t *d_out, uint2 *d_dec, const VitJob *d_jobs,
                   const VitGroup *d_groups, const uint32_t *d_bin_start, int n_ctas, int warps_per_cta,
                   cudaStream_t st);

// Soft-decision decoder (opt-in): rows of 4 received symbols per trellis step (one 32-bit word per step,
// rows 16-byte aligned and padded to whole 16-step chunks = 64 bytes), decoded with the reference's own
// metric table gen_met(1, 1.0, 0, 4) (viterbi.c:126-191); symbols outside 121..135 saturate.
// Same jobs / groups / decision scratch as the hard-decision kernel; one group per warp, no work lists.
__host__ __device__ static inline uint32_t vit_soft_row_bytes(uint32_t nsteps) { return 4u * ((nsteps + 15u) & ~15u); }
int launch_viterbi_soft(const uint8_t *d_sym, uint8_t *d_out, uint2 *d_dec, const VitJob *d_jobs,
                        const VitGroup *d_groups, int n_groups, cudaStream_t st);
int launch_soft_rows(const uint8_t *d_soft, uint64_t soft_stride, uint8_t *d_rows, int n_cw, uint32_t nsteps,
                     cudaStream_t st);
void viterbi_soft_metrics(int32_t out[2][256]);  // the table the kernel uses (INT32_MIN outside 121..135)

// step-byte producers -------------------------------------------------------------------
// (a) from the reference's soft-symbol bytes (4 per step; <128 -> 0, 128 -> erasure, >128 -> 1)
int launch_prep_soft(const uint8_t *d_soft, uint64_t soft_stride, uint8_t *d_steps, uint64_t row_stride,
                     int n_cw, uint32_t nsteps, cudaStream_t st);
// (b) from contiguous punctured hard bits (1 byte per bit) and a puncturing layout
struct ShapeDev {  // device copy of dabgpu_cw_shape with expanded masks
  int32_t nbits, in_bits, n_regions, pad;
  struct {
    int32_t steps, step0, in0, ones;  // ones = kept bits per 32-bit period
    uint32_t mask;                    // puncturing vector of the period (bit p = coded bit p is sent)
    // Depositing the `ones` received bits x of a period into the mask positions (nibble k = step k):
    // the kept symbols of a step are its first popc() ones, so every nibble is a left shift of x:
    //     r_nibbles = OR_t (x << dep_shift[t]) & dep_mask[t]
    uint32_t n_terms;
    uint32_t e_lo, e_hi;              // the mask nibbles as the high nibbles of 8 step bytes
    uint8_t dep_shift[8];
    uint32_t dep_mask[8];
  } r[5];
};
void shape_to_dev(const dabgpu_cw_shape &s, ShapeDev *o);
// codeword cw reads its bits at G * group_stride + (cw % per_group) * bits_stride with
// G = cw / per_group, or d_group_index[cw / per_group] when an index is given
int launch_prep_hard(const uint8_t *d_bits, uint64_t bits_stride, uint32_t per_group, uint64_t group_stride,
                     const uint32_t *d_group_index, uint8_t *d_steps, uint64_t row_stride, int n_cw,
                     const ShapeDev *d_shape, uint32_t nsteps, cudaStream_t st);
// rows[i] of `row_bytes` bytes (multiple of 4): dst + dst_off[i] <- src + i * row_bytes
int launch_scatter_rows(const uint8_t *d_src, uint32_t row_bytes, const uint64_t *d_dst_off, uint8_t *d_dst,
                        int n_rows, cudaStream_t st);

// FIB CRC check (src/misc.c:145-150) over n FIBs of 32 bytes -> 1/0 per FIB
int launch_fib_crc(const uint8_t *d_fibs, uint8_t *d_ok, int n_fibs, cudaStream_t st);

// in-place XOR with the energy-dispersal PRBS (standalone form of misc.c:41-58)
int launch_descramble(uint8_t *d_buf, uint64_t stride, int n_rows, int nbytes, cudaStream_t st);

}  // namespace dabgpu
