// engine.cuh -- the batched receiver: S independent ensemble streams in lock-step, each one
// an instance of the reference's demod_thread_fn loop (dab2eti.c:60-115):
//      sdr_demod -> dab_process_frame -> tuner feedback
// with the sample/bit arithmetic on the GPU and the per-stream control state on the host.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "hostlogic.cuh"
#include "msc.cuh"
#include "ofdm.cuh"
#include "vitbatch.cuh"

namespace dabgpu {

// a few helper threads for the per-stream host state machines (streams are independent)
class HostPool {
 public:
  ~HostPool() { stop(); }
  void start(int n_threads);
  void stop();
  // fn(i) for i in [0, n); returns when all are done.  The calling thread takes part.
  void run(int n, const std::function<void(int)> &fn);

 private:
  void worker();
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)> *fn_ = nullptr;
  int n_ = 0, gen_ = 0, busy_ = 0;
  std::atomic<int> next_{0};
  bool quit_ = false;
};

// derived from a stream's ens_info whenever its sub-channel table changes
struct EnsLayout {
  uint64_t version = 0;
  uint64_t epoch = 0;  // engine-wide counter value of the last re-derivation (job-list cache key)
  int nsub = 0;
  struct Sub {
    uint32_t in_bit0, shape, nbits, row_off, eti_off;
  } sub[64];
  uint32_t rows_bytes = 0;  // step-byte rows of all sub-channels of one CIF
  uint32_t per0 = 0, nper = 0;  // this layout's puncturing periods in the engine's period store
  uint32_t e1 = 0;          // header length
  EnsDev dev;
};

// sdr_state_t's scalar part (input_sdr.h:12-41) + FIFO bookkeeping (sdr_fifo.h:27-33)
struct FrontState {
  uint32_t frequency = 0;
  uint32_t fifo_start = 0, fifo_count = 0;
  int32_t coarse_timeshift = 0, fine_timeshift = 0, coarse_freq_shift = 0;
  double fine_freq_shift = 0;
  int32_t startup_delay = 0;
  uint8_t force_timesync = 0;
  uint64_t samples_in = 0;  // samples ingested so far (virtual-tuner phase reference)
  int last_ok = 0;
  bool pending = false;     // a frame of this stream is in flight in the current step
  bool prev_ring = false;   // the previous frame was used in place from the ring (never copied)
  uint32_t prev_pos = 0;    // ... and started at this ring offset
  GlibcRand rng;
};

struct StreamStats {
  uint64_t frames_demodulated = 0, eti_frames = 0, fib_crc_errors = 0, undecodable_subch = 0;
};

struct Engine {
  int S = 0;
  uint32_t f0 = 0;
  bool quiet = true;
  bool virtual_tuner = false;
  bool follow_reconfig = false;  // DABGPU_ENGINE_FOLLOW_RECONFIG (hostlogic.cuh)
  bool soft = false;  // DABGPU_ENGINE_SOFT: symbols instead of bits from the demapper to the Viterbi decoder
  std::vector<FrontState> front;
  std::vector<BackendState> back;
  std::vector<EnsLayout> layout;
  uint64_t layout_epoch = 0;
  std::vector<uint64_t> subch_mask;  // per stream: SubChIds that are decoded and carried in the ETI
  std::vector<StreamStats> stats;

  // puncturing-period lists of the multiplex layouts seen so far (host copy mirrored in d_periods);
  // streams with the same sub-channel table share one list
  std::vector<PeriodDesc> periods;
  struct LayoutKey {
    std::vector<uint32_t> sig;  // (in_bit0, shape, row_off) per sub-channel
    uint32_t per0, nper;
  };
  std::vector<LayoutKey> layout_keys;
  bool periods_dirty = false;
  DevBuf d_periods;
  // shapes seen so far (host list mirrored in d_shapes)
  std::vector<dabgpu_cw_shape> shapes;
  bool shapes_dirty = false;
  int shape_index(const dabgpu_cw_shape &s);

  // device stores
  DevBuf d_ring, d_frames, d_tails, d_chunk, d_ctl, d_sync;
  DevBuf d_cifs;      // [S][20][CIF_BYTES]
  DevBuf d_cifs_soft; // soft engines: [S][CIF_SLOTS][55296] symbols in logical order
  DevBuf d_fibs;      // [S][5][384]
  DevBuf d_crc;       // [S][5][12]
  DevBuf d_ficbits;   // [S][9216]
  DevBuf d_tfbytes;   // staging of demapped TFs given on the host
  DevBuf d_steps_fic, d_steps_msc;
  DevBuf d_eti;       // [4*S][6144]
  DevBuf d_wf_ring, d_wf_pkts, d_wf_ctl;  // Wavefinder producer: tfs[5] per stream, packet staging, counts/slots/flags
  DevBuf d_consume;   // output of the device-side ETI consumers (extract_subchannel / check_eti)
  DevBuf d_ens;       // EnsDev[S]
  DevBuf d_shapes, d_fic_shape;
  DevBuf d_cifjobs, d_subjobs, d_etijobs, d_planeoff, d_gather_idx, d_gather_out;
  PinBuf h_ctl, h_stepctl[2], h_sync, h_fic_out[2], h_jobs[2], h_msc[2], h_eti, h_chunk;
  // MSC work runs on its own stream so that it overlaps the next frames' front-end kernels
  // ... and the FIC chain (depuncture, Viterbi, CRC, copy back) on another one, so that the host
  // gets the FIBs while the CIF symbols of the same frame are still being demodulated
  cudaStream_t st_fic = nullptr;
  cudaEvent_t ev_fic_ready = nullptr, ev_demod_done[2] = {};
  int demod_ev_next = 0;  // the two events alternate from frame to frame
  cudaEvent_t ev_ctl[2] = {};  // StepCtl staging buffers alternate; each is reusable once uploaded
  int ctl_buf = 0;
  int msc_wait_ev = -1;   // event the next MSC batch has to wait for (-1: its CIFs are in place)
  cudaStream_t st_msc = nullptr;
  cudaEvent_t ev_up[2] = {}, ev_msc_done = nullptr;
  int msc_buf = 0;
  bool msc_inflight = false;
  int join_msc(cudaStream_t user);  // make `user` wait for all MSC work issued so far
  // host IQ is staged through a small ring of device buffers on a copy stream, so that the
  // uploads of the next callbacks overlap the processing of the current one
  enum { N_STAGE = 3 };
  cudaStream_t st_copy = nullptr;
  DevBuf d_stage[N_STAGE];
  cudaEvent_t ev_copied[N_STAGE] = {}, ev_consumed[N_STAGE] = {};
  int stage_len[N_STAGE] = {};
  int stage_head = 0, stage_count = 0;  // FIFO of submitted chunks (at most N_STAGE)
  uint64_t submit_seq = 0, stage_seq[N_STAGE] = {};  // submit_iq calls so far / the call each stage holds
  int uploads_in_flight();  // submitted chunks whose host->device copy has not completed yet
  int consuming_stage = -1;
  int submit_iq(const uint8_t *iq, size_t pitch, int chunk_len);
  int feed_submitted();
  bool bulk_copies = false;  // host-buffer path in use: the copy engines carry bulk transfers
  int ctl_transfer(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st);
  // zero-copy source: one contiguous device-resident capture per stream, consumed in place
  RingGeom rg = {nullptr, 0, IQ_RING_BYTES};
  const uint8_t *capture_base = nullptr;
  uint64_t capture_len = 0, capture_fed = 0;
  bool capture_cyclic = false;  // feed_capture wraps around: the capture is one period of an endless signal
  bool capture_call = false;
  int attach_capture(const uint8_t *iq_device, size_t pitch, size_t len);
  int feed_capture(int chunk_len);
  HostPool pool;
  std::vector<FrameWork> works;
  // Transmission frames whose FIC has been decoded but whose dab_process_frame is still to run.
  // With msc_batch > 1 that happens during the next feed_iq call, after its kernels have been
  // queued, so that the host state machines overlap GPU work instead of stalling it.
  struct LagFrame {
    bool valid = false;
    std::vector<int> active;    // streams with a frame
    std::vector<uint8_t> proc;  // frame passed the synchroniser checks of sdr_demod
    std::vector<int> slot;      // physical TF slot of the frame
    int fic_buf = 0;            // which h_fic_out holds its FIBs
    bool fic_wait = false;      // ... once ev_fic_done[fic_buf] has fired (early-return mode)
    int demod_ev = -1;          // ev_demod_done index covering its CIFs
  } lag;
  int fic_buf = 0;
  // Early return (with trailing back-end logic): feed_iq only waits for the synchroniser outputs,
  // which is all the next FIFO read depends on; the FIC results are awaited where they are used.
  cudaEvent_t ev_sync_done = nullptr, ev_sync_out = nullptr, ev_prep_done = nullptr, ev_fic_done[2] = {},
              ev_jobs[2] = {};
  bool prep_pending = false;
  int jobs_buf = 0;
  std::vector<int> frame_slot;  // per stream: slot of the frame in flight
  int take_slot(int s) {
    const int slot = back[s].phys;
    back[s].phys = (slot + 1) % PHYS_TF_SLOTS;
    return frame_slot[s] = slot;
  }
  int backend_host(cudaStream_t st);
  bool hold_msc_hint = false;  // feed_iq: the next callback is expected to complete frames again
  bool trailing_hint_enabled() const { return msc_batch > 1 && msc_batch < MAX_MSC_BATCH && !timing; }
  // MSC decoding may lag by up to msc_batch calls so that one Viterbi launch covers several
  // transmission frames per stream (more, better balanced work per launch)
  int msc_batch = 1, pend_calls = 0;
  uint64_t row_base = 0;
  std::vector<int32_t> pend_stream;
  std::vector<int> pend_of_stream;   // queued frames per stream
  std::vector<uint64_t> frame_row;   // step-byte row base of every frame of the cached job list
  uint64_t pend_sig = 0xcbf29ce484222325ull, cached_sig = 0;
  int flush_msc(cudaStream_t st);
  VitBatch vb_fic, vb_msc;

  // optional per-kernel device timing (CUDA events on the launch stream), for bench.py's roofline
  enum KernelId { K_INGEST, K_FIFO, K_SYNC, K_DEMOD, K_FIC_PREP, K_FIC_VIT, K_MSC_GATHER, K_MSC_VIT, K_ETI, K_COUNT };
  bool timing = false;
  cudaEvent_t ev[K_COUNT][2] = {};
  bool ev_used[K_COUNT] = {};
  double ms_total[K_COUNT] = {};
  uint64_t n_total[K_COUNT] = {};
  void t0(int k, cudaStream_t st) {
    if (timing) {
      cudaEventRecord(ev[k][0], st);
    }
    if (trace_on) trace_mark(k, 0, st);
  }
  void t1(int k, cudaStream_t st) {
    if (timing) {
      cudaEventRecord(ev[k][1], st);
      ev_used[k] = true;
    }
    if (trace_on) trace_mark(k, 1, st);
  }
  // development aid (DABGPU_TRACE=<file>): GPU timeline of the kernels above without serialising
  // anything -- one event pair per launch, dumped as "kernel start_ms end_ms" when the engine dies
  struct TraceRec {
    int k;
    cudaEvent_t e[2];
    double host_us;
  };
  bool trace_on = false;
  std::vector<TraceRec> trace;
  cudaEvent_t trace_base = nullptr;
  void trace_mark(int k, int which, cudaStream_t st);
  void trace_dump();
  // host-side wall-clock breakdown of a step (always on; microseconds, cumulative)
  enum HostPhase { H_PRE, H_WAIT, H_FSM, H_JOBS, H_COUNT };
  double host_us[H_COUNT] = {};
  int collect_timing(cudaStream_t st);
  int enable_timing(bool on);

  // results of the last step
  int n_eti = 0;
  std::vector<int32_t> eti_stream;
  uint64_t trellis_steps = 0;

  // scratch vectors reused between steps
  std::vector<int> active;
  std::vector<CifJob> cifjobs;
  std::vector<EtiJob> etijobs;

  int init(int n_streams, uint32_t tuner_hz, int flags);
  void destroy();
  int ensure_frontend();
  // one rtlsdr callback worth of IQ for every stream
  int feed_iq(const uint8_t *iq, size_t pitch, int chunk_len, bool on_device);
  // one demapped TF (fic 9216 + msc 221184 bytes of 0/1) for every stream with mask[s] != 0
  int process_demapped(const uint8_t *tfs, size_t pitch, const uint8_t *mask, bool on_device);
  // the Wavefinder packets of one transmission frame per stream (n_packets[s] == 0: no frame)
  int process_wavefinder(const uint8_t *packets, size_t pitch, const int32_t *n_packets);

 private:
  int fic_launch(cudaStream_t st, const uint8_t *d_fic_src, uint64_t fic_stride, bool early);
  int fic_finish(cudaStream_t st, const SyncOut *h_sync_or_null, int demod_ev, bool early);
  int refresh_layout(int s);
  int upload_tables(cudaStream_t st);
};

}  // namespace dabgpu
