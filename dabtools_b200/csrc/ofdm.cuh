// ofdm.cuh -- RTL-SDR front-end on the GPU: IQ ingest (virtual tuner), FIFO frame read,
// synchronisers and the OFDM demodulator.  Replaces src/input_sdr.c, src/sdr_sync.c and
// src/sdr_fifo.c for batches of independent streams.
#pragma once
#include "common.cuh"

namespace dabgpu {

enum : uint32_t {
  IQ_RING_BYTES = 196608u * 2u * 4u,  // input_sdr.c:170: FIFO of four frames
  TAIL_BYTES = 2048u,                 // >= the largest stale tail (negative fine time shift: 1536 bytes)
  TAIL_OFF = 393216u - TAIL_BYTES
};

// Where the streams' sample FIFOs live: the engine's own ring (pitch = mod = IQ_RING_BYTES), or a
// caller-owned capture that is consumed in place (dabgpu_engine_attach_capture: no ingest copy).
// Stream s, FIFO offset p -> base + s * pitch + (p mod `mod`); base, pitch and mod are multiples of 16.
struct RingGeom {
  const uint8_t *base;
  uint64_t pitch;
  uint32_t mod;
};
// offsets stay below 2 * mod everywhere (a position plus less than one frame)
__host__ __device__ static inline uint32_t ring_wrap(uint32_t x, uint32_t mod) { return x >= mod ? x - mod : x; }

// Per-stream, per-step control block, written by the host FSM, read by the kernels.
struct StepCtl {
  // ---- ingest (rtlsdr_callback + cbWrite loop, dab2eti.c:125, input_sdr.c:36-38)
  uint32_t wr_pos;       // ring offset the chunk is appended at
  int32_t nco_hz;        // virtual tuner: mix by -nco_hz while ingesting (0 = plain copy)
  uint64_t nco_sample0;  // absolute index of the chunk's first sample (phase continuity)
  // ---- frame read (sdr_read_fifo, sdr_fifo.c:43-61): up to two copies into the frame buffer
  uint32_t rd_pos[2];    // ring offsets
  uint32_t rd_dst[2];    // destination offsets inside the 393216-byte frame buffer
  uint32_t rd_bytes[2];  // 0 = nothing
  // ---- where the kernels find the frame (sdr->buffer).  src_ring = 1: the read was one plain
  //      window, so byte b of the frame is ring[(src_pos + b) % RING] for b < TAIL_OFF and
  //      tail[b - TAIL_OFF] above (the tail store carries the stale bytes a negative shift leaves
  //      behind); no copy is made.  src_ring = 0: the persistent frame buffer (rare: FIFO ran dry).
  uint32_t src_ring;
  uint32_t src_pos;
  uint32_t mat;          // materialise the previous (ring-mode) frame into the frame buffer first
  uint32_t mat_pos;
  // ---- sync + demod (sdr_demod, input_sdr.c:60-162)
  uint32_t run;             // frame goes through the synchronisers / demodulator
  uint32_t force_timesync;  // sdr->force_timesync on entry
  uint64_t cif_off[4];      // CIF store offsets for the 4 CIFs of this transmission frame
};

// What sdr_demod leaves behind in sdr_state_t (input_sdr.h:17-20) plus its return value.
struct SyncOut {
  int32_t ok;                // 1 = frame demodulated
  int32_t coarse_timeshift;  // bytes
  int32_t fine_timeshift;    // bytes
  int32_t coarse_freq_shift; // carriers
  double fine_freq_shift;    // Hz (double like sdr_state_t: the tuner feedback truncates it)
  int32_t stage;             // how far the frame got: 1 coarse-time miss, 2 coarse-freq miss, 3 demodulated
  float null_energy;
};

int launch_ingest(const uint8_t *d_src, uint64_t src_pitch, uint32_t chunk_len, uint8_t *d_ring,
                  const StepCtl *d_ctl, int n_streams, cudaStream_t st);
// fallback copy into the frame buffer for the streams with src_ring == 0 (materialise = true runs
// the "previous frame" pass for the streams with mat == 1 instead)
int launch_fifo_read(RingGeom ring, const uint8_t *d_tails, uint8_t *d_frames, const StepCtl *d_ctl,
                     int n_streams, bool materialise, cudaStream_t st);
// keep the tail store equal to the last TAIL_BYTES of every stream's logical frame buffer
int launch_tail_update(RingGeom ring, const uint8_t *d_frames, uint8_t *d_tails, const StepCtl *d_ctl,
                       int n_streams, cudaStream_t st);
int launch_sync(RingGeom ring, const uint8_t *d_tails, const uint8_t *d_frames, const StepCtl *d_ctl,
                SyncOut *d_out, int n_streams, cudaStream_t st);
// fic_bits: [n_streams][9216] one byte per bit (reference layout); MSC goes to the CIF store as planes
// segments: 0 = PRS + the 3 FIC symbols, 1..4 = the 4 CIFs; launched as [seg_first, seg_first+seg_count)
// soft: symbols (128 -+ round(8 x), 121..135) instead of bits: fic_bits gets 9216 symbol bytes, d_cifs is the
// soft CIF store (55296 bytes per CIF, logical order)
int launch_demod(RingGeom ring, const uint8_t *d_tails, const uint8_t *d_frames, const StepCtl *d_ctl,
                 const SyncOut *d_sync, uint8_t *d_fic_bits, uint8_t *d_cifs, int n_streams, int seg_first,
                 int seg_count, bool soft, cudaStream_t st);

// debug / parity variants on a single frame buffer: raw spectra (fftshifted, 76x2048 complex
// float), DQPSK products (rows 1..75) and the reference's byte-per-bit demapped output
int launch_demod_debug(const uint8_t *d_frame, float2 *d_symbols, float2 *d_symbols_d, uint8_t *d_bits230400,
                       cudaStream_t st);

// the four synchronisers on their own (reference-signature entry points of sdr_sync.h)
int launch_sync_single(int mode, const void *d_in, int force, int *d_res, double *d_fres, cudaStream_t st);

}  // namespace dabgpu
