// msc.cuh -- MSC side of the back-end: CIF storage, time de-interleave + depuncture
// gather (feeds the Viterbi kernel) and ETI(NI) frame assembly.
//
// Device format of one received CIF ("planes"): the 55296 channel bits are kept packed
// and split by i mod 16, because the time de-interleaver (misc.c:29-39) takes bit i of
// the logical frame from window slot map[i & 15]: plane m holds bits i = 16q + m,
// q = 0..3455, bit q at word q>>5, bit q&31 (LSB first).  A CIF is 16 planes x 108
// words = 6912 bytes, and de-interleaving one logical frame is 16 contiguous 432-byte
// reads, one plane from each of the 16 window slots.
#pragma once
#include <vector>

#include "viterbi.cuh"

namespace dabgpu {

enum { CIF_PLANE_WORDS = 108, CIF_WORDS = 16 * 108, CIF_BYTES = 6912 };

// one logical (output) CIF to de-interleave and turn into Viterbi step bytes
struct CifJob {
  uint64_t slot_off[16];  // byte offsets of the 16 window CIFs (slot 0 = oldest) in the CIF store
  uint32_t sub0, nsub;    // range in the SubJob array (launch_msc_gather)
  uint64_t row_base;      // step-byte rows of this CIF start here (launch_msc_gather_periods)
  uint32_t per0, nper;    // range in the PeriodDesc array
};
// One puncturing period (8 trellis steps -> 8 step bytes) of a multiplex layout.  A layout's
// periods depend only on its sub-channel table, so the list is built once per layout on the host
// and shared by every CIF (and every stream) that uses the layout.
struct PeriodDesc {
  uint32_t in_tab;    // first channel bit of the period | deposit table index << 16 (0xff: padding, all zero)
  uint32_t row_off8;  // byte offset of the 8 step bytes relative to CifJob::row_base, divided by 8
};
// append the periods of one sub-channel codeword (uep/eep_depuncture, depuncture.c:84-132);
// returns false for a puncturing layout outside EN 300 401 (no deposit table)
bool append_periods(const dabgpu_cw_shape &shape, uint32_t in_bit0, uint32_t row_off, std::vector<PeriodDesc> &out);
bool host_periods_to_steps(const dabgpu_cw_shape &shape, uint32_t in_bit0, const uint8_t *bits55296,
                           uint8_t *steps);
int launch_msc_gather_periods(const uint8_t *d_cifs, const CifJob *d_jobs, const PeriodDesc *d_periods,
                              uint8_t *d_steps, int n_jobs, cudaStream_t st);
struct SubJob {
  uint64_t row_off;   // step-byte row of this codeword (16-byte aligned)
  uint32_t in_bit0;   // first channel bit of the sub-channel inside the CIF (start_cu * 64)
  uint32_t shape;     // index into the ShapeDev table
};

int launch_msc_gather(const uint8_t *d_cifs, const CifJob *d_jobs, const SubJob *d_subs,
                      const ShapeDev *d_shapes, uint8_t *d_steps, int n_jobs, cudaStream_t st);

// 1 byte/bit demapped MSC (72 x 3072 per TF, the reference's msc_symbols_demapped) ->
// 4 CIF plane buffers; dst_off[4*i + c] = byte offset of CIF c of TF i in the CIF store
int launch_pack_planes(const uint8_t *d_msc_bytes, uint64_t tf_stride, const uint64_t *d_dst_off,
                       uint8_t *d_cifs, int n_tf, cudaStream_t st);
// inverse of the gather's first half, for the time_deinterleave() drop-in and tests:
// 16 window CIFs (planes) -> 55296 bytes of 0/1 in logical order
int launch_deinterleave_bytes(const uint8_t *d_cifs, const CifJob *d_jobs, uint8_t *d_out, int n_jobs,
                              cudaStream_t st);

// ---- ETI assembly (misc.c:153-314) -------------------------------------------------------
struct EnsDev {          // per stream, changes only when the multiplex description changes
  uint32_t nst;          // number of sub-channels
  uint32_t fl;           // FL field (words)
  uint32_t payload;      // sum of sub-channel bytes in the MST
  uint32_t pad;
  uint8_t stc[64][4];    // STC words in SubChId order, as they appear in the frame
};
struct EtiJob {
  uint64_t fib_off;      // byte offset of the 96 FIB bytes of the oldest CIF in the FIB store
  uint32_t ens;          // index into EnsDev array (stream)
  uint8_t cif_hi, cif_lo, pad[2];
};
// frame f is written to d_eti + 6144*f; the sub-channel payload must already be in place
int launch_eti_pack(const EtiJob *d_jobs, const EnsDev *d_ens, const uint8_t *d_fibs, uint8_t *d_eti,
                    int n_frames, cudaStream_t st);

// ---- soft-decision data path (opt-in): CIFs as 55296 symbol bytes, rows of 4 symbols per step ----
int launch_soft_store(const uint8_t *d_msc_bytes, uint64_t tf_stride, const uint64_t *d_dst_off, uint8_t *d_cifs_soft,
                      int n_tf, cudaStream_t st);
int launch_msc_soft_gather(const uint8_t *d_cifs_soft, const CifJob *d_jobs, const PeriodDesc *d_periods,
                           uint8_t *d_rows, int n_jobs, cudaStream_t st);
int launch_fic_soft_rows(const uint8_t *d_fic, uint64_t stride, const uint32_t *d_index, uint8_t *d_rows, int n_cw,
                         const ShapeDev *d_shape, cudaStream_t st);

// ---- Wavefinder producer (input_wf.c:23-115): USB packets -> demapped transmission frames --------
int launch_wf_demap(const uint8_t *d_packets, uint64_t pitch, const int32_t *d_n_packets, int max_packets,
                    const int32_t *d_slot, uint8_t *d_tf_ring, uint32_t *d_fic_seen, uint8_t *d_tf_out, int n_streams,
                    cudaStream_t st);

// ---- ETI consumers on the device (eti2mpa.c:32-67; TODO.md:10-11) ---------------------------
int launch_eti_extract(const uint8_t *d_eti, int n_frames, int subchid, uint8_t *d_out, uint64_t out_pitch,
                       int32_t *d_len, cudaStream_t st);
int launch_eti_check(const uint8_t *d_eti, int n_frames, uint32_t *d_flags, cudaStream_t st);

}  // namespace dabgpu
