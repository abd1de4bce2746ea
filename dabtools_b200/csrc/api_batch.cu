// api_batch.cu -- stateless batched entry points of include/dabgpu.h.
#include "../../include/dabgpu.h"
#include "msc.cuh"
#include "ofdm.cuh"
#include "vitbatch.cuh"

using namespace dabgpu;

namespace {

// per-thread workspace of the stateless batched calls
struct Workspace {
  DevBuf in, steps, out, aux, shape;
  VitBatch vb, vb_soft;
  uint64_t last_steps = 0;
};
thread_local Workspace t_ws;

}  // namespace

namespace dabgpu {
Workspace &thread_workspace() { return t_ws; }
}  // namespace dabgpu

DABGPU_EXPORT uint64_t dabgpu_last_trellis_steps(void) { return t_ws.last_steps; }

// ---- tables -----------------------------------------------------------------------------------
DABGPU_EXPORT int dabgpu_tab_shape(int kind, int a, int b, int32_t *out23) {
  dabgpu_cw_shape sh;
  int rc = 0;
  if (kind == 0)
    dabgpu_shape_fic(&sh);
  else if (kind == 1)
    rc = dabgpu_shape_uep(&sh, a);
  else if (kind == 2)
    rc = dabgpu_shape_eep(&sh, a, b, -1);
  else
    rc = -1;
  if (rc) {
    set_error(DABGPU_ERR_ARG, "dabgpu_tab_shape: no such profile (kind=%d a=%d b=%d)", kind, a, b);
    return DABGPU_ERR_ARG;
  }
  static_assert(sizeof(sh) == 23 * sizeof(int32_t), "shape layout");
  memcpy(out23, &sh, sizeof sh);
  return DABGPU_OK;
}
// host-only: the step bytes the MSC gather produces for one sub-channel of a de-interleaved CIF
DABGPU_EXPORT int dabgpu_tab_depuncture_steps(int kind, int a, int b, int start_cu, const uint8_t *cif_bits55296,
                                              uint8_t *steps, int steps_cap) {
  dabgpu_cw_shape sh;
  int32_t raw[23];
  if (dabgpu_tab_shape(kind, a, b, raw)) return DABGPU_ERR_ARG;
  memcpy(&sh, raw, sizeof sh);
  const int need = (int)vit_row_bytes((uint32_t)sh.nbits + 6u);
  if (steps_cap < need || start_cu < 0 || start_cu * 64 + sh.in_bits > DABGPU_CIF_BITS) {
    set_error(DABGPU_ERR_ARG, "dabgpu_tab_depuncture_steps: buffer too small or sub-channel outside the CIF");
    return DABGPU_ERR_ARG;
  }
  if (!host_periods_to_steps(sh, (uint32_t)start_cu * 64u, cif_bits55296, steps)) {
    set_error(DABGPU_ERR_ARG, "dabgpu_tab_depuncture_steps: puncturing outside EN 300 401");
    return DABGPU_ERR_ARG;
  }
  return need;
}
DABGPU_EXPORT void dabgpu_tab_uep(int32_t *o) {
  for (int i = 0; i < 64; i++) {
    const dabgpu_uep_profile &p = DABGPU_UEP[i];
    int32_t *r = o + 12 * i;
    r[0] = p.bitrate;
    r[1] = p.size_cu;
    r[2] = p.prot_level;
    for (int k = 0; k < 4; k++) {
      r[3 + k] = p.L[k];
      r[7 + k] = p.PI[k];
    }
    r[11] = p.pad_bits;
  }
}
DABGPU_EXPORT uint32_t dabgpu_tab_puncture_mask(int pi) { return dabgpu_puncture_mask(pi); }
DABGPU_EXPORT void dabgpu_tab_freq_deint(uint16_t *rev) { dabgpu_build_freq_deint(rev); }
DABGPU_EXPORT void dabgpu_tab_prs(uint8_t *q) { dabgpu_build_prs(q); }
DABGPU_EXPORT void dabgpu_tab_prbs(uint8_t *out, int nbytes) { dabgpu_build_prbs(out, nbytes); }

// ---- viterbi batch -------------------------------------------------------------------------------
DABGPU_EXPORT int dabgpu_viterbi_batch(const uint8_t *soft, size_t soft_pitch, int n, int nbits,
                                       uint8_t *out, size_t out_pitch, int descramble, int on_device) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n < 0 || nbits <= 0 || !soft || !out) {
    set_error(DABGPU_ERR_ARG, "dabgpu_viterbi_batch: bad argument");
    return DABGPU_ERR_ARG;
  }
  const uint32_t nsteps = (uint32_t)nbits + 6;
  const size_t out_row = 4 * (((size_t)nbits + 31) / 32);
  if (soft_pitch < 4ull * nsteps || (descramble && nbits > 9216) ||
      (on_device && (out_pitch < out_row || (out_pitch & 3)))) {
    set_error(DABGPU_ERR_ARG, "dabgpu_viterbi_batch: pitch/size constraint violated");
    return DABGPU_ERR_ARG;
  }
  if (n == 0) return DABGPU_OK;
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  const uint32_t row = vit_row_bytes(nsteps);
  if ((rc = ws.steps.reserve((size_t)n * row))) return rc;
  const uint8_t *d_soft = soft;
  uint8_t *d_out = out;
  size_t d_out_pitch = out_pitch;
  if (!on_device) {
    if ((rc = ws.in.reserve((size_t)n * soft_pitch))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ws.in.p, soft, (size_t)n * soft_pitch, cudaMemcpyHostToDevice, st));
    d_soft = ws.in.as<uint8_t>();
    d_out_pitch = out_row;
    if ((rc = ws.out.reserve((size_t)n * d_out_pitch))) return rc;
    d_out = ws.out.as<uint8_t>();
  }
  if ((rc = launch_prep_soft(d_soft, soft_pitch, ws.steps.as<uint8_t>(), row, n, nsteps, st))) return rc;
  ws.vb.clear();
  for (int i = 0; i < n; i++)
    ws.vb.add((uint64_t)i * row, (uint64_t)i * d_out_pitch, (uint32_t)nbits, descramble ? VIT_DESCRAMBLE : 0);
  if ((rc = ws.vb.run(ws.steps.as<uint8_t>(), d_out, st))) return rc;
  ws.last_steps = ws.vb.total_steps;
  if (!on_device) {
    const size_t nbytes = ((size_t)nbits + 7) / 8;
    CUDA_TRY(cudaMemcpy2DAsync(out, out_pitch ? out_pitch : nbytes, d_out, d_out_pitch, nbytes, n,
                               cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DABGPU_OK;
}

// Soft-decision variant: `soft` really is soft -- any symbol value, weighted with the reference's own
// metric table (viterbi.c:126-191; defined for 121..135, saturating beyond) -- instead of being sliced
// to 0 / erasure / 1.  Bit-exact with the reference's viterbi() for symbols inside the table.
DABGPU_EXPORT int dabgpu_viterbi_soft_batch(const uint8_t *soft, size_t soft_pitch, int n, int nbits,
                                            uint8_t *out, size_t out_pitch, int descramble, int on_device) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n < 0 || nbits <= 0 || !soft || !out) {
    set_error(DABGPU_ERR_ARG, "dabgpu_viterbi_soft_batch: bad argument");
    return DABGPU_ERR_ARG;
  }
  const uint32_t nsteps = (uint32_t)nbits + 6;
  const size_t out_row = 4 * (((size_t)nbits + 31) / 32);
  if (soft_pitch < 4ull * nsteps || (descramble && nbits > 9216) ||
      (on_device && (out_pitch < out_row || (out_pitch & 3)))) {
    set_error(DABGPU_ERR_ARG, "dabgpu_viterbi_soft_batch: pitch/size constraint violated");
    return DABGPU_ERR_ARG;
  }
  if (n == 0) return DABGPU_OK;
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  ws.vb_soft.soft = true;
  const uint32_t row = vit_soft_row_bytes(nsteps);
  if ((rc = ws.steps.reserve((size_t)n * row))) return rc;
  const uint8_t *d_soft = soft;
  uint8_t *d_out = out;
  size_t d_out_pitch = out_pitch;
  if (!on_device) {
    if ((rc = ws.in.reserve((size_t)n * soft_pitch))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ws.in.p, soft, (size_t)n * soft_pitch, cudaMemcpyHostToDevice, st));
    d_soft = ws.in.as<uint8_t>();
    d_out_pitch = out_row;
    if ((rc = ws.out.reserve((size_t)n * d_out_pitch))) return rc;
    d_out = ws.out.as<uint8_t>();
  }
  if ((rc = launch_soft_rows(d_soft, soft_pitch, ws.steps.as<uint8_t>(), n, nsteps, st))) return rc;
  ws.vb_soft.clear();
  for (int i = 0; i < n; i++)
    ws.vb_soft.add((uint64_t)i * row, (uint64_t)i * d_out_pitch, (uint32_t)nbits, descramble ? VIT_DESCRAMBLE : 0);
  if ((rc = ws.vb_soft.run(ws.steps.as<uint8_t>(), d_out, st))) return rc;
  ws.last_steps = ws.vb_soft.total_steps;
  if (!on_device) {
    const size_t nbytes = ((size_t)nbits + 7) / 8;
    CUDA_TRY(cudaMemcpy2DAsync(out, out_pitch ? out_pitch : nbytes, d_out, d_out_pitch, nbytes, n,
                               cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DABGPU_OK;
}
DABGPU_EXPORT void dabgpu_tab_soft_metrics(int32_t *out2x256) {
  viterbi_soft_metrics(reinterpret_cast<int32_t(*)[256]>(out2x256));
}

// ---- FIC decode batch ------------------------------------------------------------------------------
DABGPU_EXPORT int dabgpu_fic_decode_batch(const uint8_t *fic_bits, int n_groups, uint8_t *fibs,
                                          uint8_t *crc_ok, int on_device) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n_groups < 0 || !fic_bits || !fibs || !crc_ok) {
    set_error(DABGPU_ERR_ARG, "dabgpu_fic_decode_batch: bad argument");
    return DABGPU_ERR_ARG;
  }
  if (n_groups == 0) return DABGPU_OK;
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  const uint32_t nsteps = 774, row = vit_row_bytes(nsteps);
  if ((rc = ws.steps.reserve((size_t)n_groups * row))) return rc;
  if (!ws.shape.p) {
    dabgpu_cw_shape sh;
    ShapeDev sd;
    dabgpu_shape_fic(&sh);
    shape_to_dev(sh, &sd);
    if ((rc = ws.shape.reserve(sizeof sd))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ws.shape.p, &sd, sizeof sd, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));  // sd lives on this stack frame
  }
  const uint8_t *d_bits = fic_bits;
  uint8_t *d_fibs = fibs, *d_ok = crc_ok;
  if (!on_device) {
    if ((rc = ws.in.reserve((size_t)n_groups * 2304))) return rc;
    if ((rc = ws.out.reserve((size_t)n_groups * 96))) return rc;
    if ((rc = ws.aux.reserve((size_t)n_groups * 3))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ws.in.p, fic_bits, (size_t)n_groups * 2304, cudaMemcpyHostToDevice, st));
    d_bits = ws.in.as<uint8_t>();
    d_fibs = ws.out.as<uint8_t>();
    d_ok = ws.aux.as<uint8_t>();
  }
  if ((rc = launch_prep_hard(d_bits, 2304, 1, 2304, nullptr, ws.steps.as<uint8_t>(), row, n_groups, ws.shape.as<ShapeDev>(),
                             nsteps, st)))
    return rc;
  ws.vb.clear();
  for (int i = 0; i < n_groups; i++) ws.vb.add((uint64_t)i * row, (uint64_t)i * 96, 768, VIT_DESCRAMBLE);
  if ((rc = ws.vb.run(ws.steps.as<uint8_t>(), d_fibs, st))) return rc;
  ws.last_steps = ws.vb.total_steps;
  if ((rc = launch_fib_crc(d_fibs, d_ok, 3 * n_groups, st))) return rc;
  if (!on_device) {
    CUDA_TRY(cudaMemcpyAsync(fibs, d_fibs, (size_t)n_groups * 96, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(crc_ok, d_ok, (size_t)n_groups * 3, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DABGPU_OK;
}

// ---- single-frame front-end ---------------------------------------------------------------------------
DABGPU_EXPORT int dabgpu_sync_frame(const uint8_t *frame, int force_timesync, int32_t *out4, double *fine_freq_hz) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  if ((rc = ws.in.reserve(DABGPU_TF_BYTES))) return rc;
  if ((rc = ws.aux.reserve(sizeof(StepCtl) + sizeof(SyncOut)))) return rc;
  StepCtl ctl;
  memset(&ctl, 0, sizeof ctl);
  ctl.run = 1;
  ctl.force_timesync = force_timesync ? 1u : 0u;
  SyncOut so;
  memset(&so, 0, sizeof so);
  CUDA_TRY(cudaMemcpyAsync(ws.in.p, frame, DABGPU_TF_BYTES, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws.aux.p, &ctl, sizeof ctl, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws.aux.as<uint8_t>() + sizeof ctl, &so, sizeof so, cudaMemcpyHostToDevice, st));
  SyncOut *d_so = reinterpret_cast<SyncOut *>(ws.aux.as<uint8_t>() + sizeof ctl);
  if ((rc = launch_sync(RingGeom{nullptr, 0, IQ_RING_BYTES}, nullptr, ws.in.as<uint8_t>(), ws.aux.as<StepCtl>(), d_so, 1, st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(&so, d_so, sizeof so, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  out4[0] = so.coarse_timeshift;
  out4[1] = so.fine_timeshift;
  out4[2] = so.coarse_freq_shift;
  out4[3] = so.ok;
  if (fine_freq_hz) *fine_freq_hz = so.fine_freq_shift;
  return DABGPU_OK;
}

DABGPU_EXPORT int dabgpu_demod_frame_debug(const uint8_t *frame, float *symbols, float *symbols_d, uint8_t *bits) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  const size_t nsym = (size_t)76 * 2048 * sizeof(float2);
  if ((rc = ws.in.reserve(DABGPU_TF_BYTES))) return rc;
  if ((rc = ws.out.reserve(2 * nsym + 230400))) return rc;
  float2 *d_sym = ws.out.as<float2>();
  float2 *d_symd = d_sym + 76 * 2048;
  uint8_t *d_bits = ws.out.as<uint8_t>() + 2 * nsym;
  CUDA_TRY(cudaMemcpyAsync(ws.in.p, frame, DABGPU_TF_BYTES, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ws.out.p, 0, 2 * nsym + 230400, st));
  if ((rc = launch_demod_debug(ws.in.as<uint8_t>(), d_sym, d_symd, d_bits, st))) return rc;
  if (symbols) CUDA_TRY(cudaMemcpyAsync(symbols, d_sym, nsym, cudaMemcpyDeviceToHost, st));
  if (symbols_d) CUDA_TRY(cudaMemcpyAsync(symbols_d, d_symd, nsym, cudaMemcpyDeviceToHost, st));
  if (bits) CUDA_TRY(cudaMemcpyAsync(bits, d_bits, 230400, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DABGPU_OK;
}

// Soft demapper on one frame (the DABGPU_ENGINE_SOFT demodulator, batch of one): 393216 bytes of uint8 I/Q
// as sdr_read_fifo leaves them -> 230400 symbols (fic 9216 + msc 221184, the layout of
// demapped_transmission_frame_t) 128 -+ round(8 x), 121..135.  No synchronisers: the frame is taken as is.
DABGPU_EXPORT int dabgpu_demod_frame_soft(const uint8_t *frame, uint8_t *soft230400) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (!frame || !soft230400) {
    set_error(DABGPU_ERR_ARG, "demod_frame_soft: null pointer");
    return DABGPU_ERR_ARG;
  }
  cudaStream_t st = current_stream();
  Workspace &ws = t_ws;
  if ((rc = ws.in.reserve(DABGPU_TF_BYTES))) return rc;
  if ((rc = ws.out.reserve(230400))) return rc;
  if ((rc = ws.aux.reserve(sizeof(StepCtl) + sizeof(SyncOut) + 64))) return rc;
  StepCtl ctl;
  memset(&ctl, 0, sizeof ctl);
  ctl.run = 1;
  for (int k = 0; k < 4; k++) ctl.cif_off[k] = (uint64_t)k * CIF_BYTES;  // slot k of the soft store
  SyncOut so;
  memset(&so, 0, sizeof so);
  so.ok = 1;
  CUDA_TRY(cudaMemcpyAsync(ws.in.p, frame, DABGPU_TF_BYTES, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws.aux.p, &ctl, sizeof ctl, cudaMemcpyHostToDevice, st));
  SyncOut *d_so = reinterpret_cast<SyncOut *>(ws.aux.as<uint8_t>() + ((sizeof ctl + 15) & ~(size_t)15));
  CUDA_TRY(cudaMemcpyAsync(d_so, &so, sizeof so, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ws.out.p, 128, 230400, st));
  if ((rc = launch_demod(RingGeom{nullptr, 0, IQ_RING_BYTES}, nullptr, ws.in.as<uint8_t>(), ws.aux.as<StepCtl>(), d_so,
                         ws.out.as<uint8_t>(), ws.out.as<uint8_t>() + 9216, 1, 0, 5, true, st)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(soft230400, ws.out.p, 230400, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DABGPU_OK;
}

// ---- ETI consumers (eti2mpa.c:32-67 and the frame check) ------------------------------------------------
DABGPU_EXPORT int dabgpu_eti_extract_subchannel(const uint8_t *eti, int n_frames, int subchid, uint8_t *out,
                                                size_t out_pitch, int32_t *out_len, int on_device) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n_frames <= 0) return DABGPU_OK;
  if (!eti || !out || !out_len || subchid < 0 || subchid > 63 || out_pitch == 0) {
    set_error(DABGPU_ERR_ARG, "eti_extract_subchannel: null pointer, SubChId outside 0..63 or zero pitch");
    return DABGPU_ERR_ARG;
  }
  cudaStream_t st = current_stream();
  if (on_device) return launch_eti_extract(eti, n_frames, subchid, out, out_pitch, out_len, st);
  Workspace &ws = t_ws;
  const size_t eb = (size_t)n_frames * DABGPU_ETI_BYTES, ob = (size_t)n_frames * out_pitch;
  if ((rc = ws.in.reserve(eb))) return rc;
  if ((rc = ws.out.reserve(ob))) return rc;
  if ((rc = ws.aux.reserve((size_t)n_frames * 4 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(ws.in.p, eti, eb, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ws.out.p, 0, ob, st));
  if ((rc = launch_eti_extract(ws.in.as<uint8_t>(), n_frames, subchid, ws.out.as<uint8_t>(), out_pitch,
                               ws.aux.as<int32_t>(), st)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(out, ws.out.p, ob, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_len, ws.aux.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DABGPU_OK;
}

DABGPU_EXPORT int dabgpu_eti_check(const uint8_t *eti, int n_frames, uint32_t *flags, int on_device) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n_frames <= 0) return DABGPU_OK;
  if (!eti || !flags) {
    set_error(DABGPU_ERR_ARG, "eti_check: null pointer");
    return DABGPU_ERR_ARG;
  }
  cudaStream_t st = current_stream();
  if (on_device) return launch_eti_check(eti, n_frames, flags, st);
  Workspace &ws = t_ws;
  const size_t eb = (size_t)n_frames * DABGPU_ETI_BYTES;
  if ((rc = ws.in.reserve(eb))) return rc;
  if ((rc = ws.aux.reserve((size_t)n_frames * 4 + 64))) return rc;
  CUDA_TRY(cudaMemcpyAsync(ws.in.p, eti, eb, cudaMemcpyHostToDevice, st));
  if ((rc = launch_eti_check(ws.in.as<uint8_t>(), n_frames, ws.aux.as<uint32_t>(), st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(flags, ws.aux.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DABGPU_OK;
}
