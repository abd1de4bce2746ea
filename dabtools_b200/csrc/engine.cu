// engine.cu -- batched receiver (see engine.cuh) and its C ABI (include/dabgpu.h).
#include "engine.cuh"

#include <chrono>
#include <cmath>

namespace dabgpu {

static inline double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum { FIC_ROW = 784 /* vit_row_bytes(774) */, FIBS_PER_TF = 384, TF_SLOTS = PHYS_TF_SLOTS, CIF_SLOTS = 4 * PHYS_TF_SLOTS };

// ---- HostPool ---------------------------------------------------------------------------------
void HostPool::start(int n_threads) {
  for (int i = 0; i < n_threads; i++) th_.emplace_back([this] { worker(); });
}
void HostPool::stop() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    quit_ = true;
  }
  cv_.notify_all();
  for (auto &t : th_) t.join();
  th_.clear();
  quit_ = false;
}
void HostPool::worker() {
  int seen = 0;
  for (;;) {
    const std::function<void(int)> *fn;
    int n;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
      if (quit_) return;
      seen = gen_;
      fn = fn_;
      n = n_;
    }
    for (;;) {
      const int i0 = next_.fetch_add(16);
      if (i0 >= n) break;
      for (int i = i0; i < std::min(n, i0 + 16); i++) (*fn)(i);
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (--busy_ == 0) done_.notify_one();
    }
  }
}
void HostPool::run(int n, const std::function<void(int)> &fn) {
  if (th_.empty() || n < 64) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  {
    std::lock_guard<std::mutex> lk(mu_);
    fn_ = &fn;
    n_ = n;
    next_.store(0);
    busy_ = (int)th_.size();
    gen_++;
  }
  cv_.notify_all();
  for (;;) {
    const int i0 = next_.fetch_add(16);
    if (i0 >= n) break;
    for (int i = i0; i < std::min(n, i0 + 16); i++) fn(i);
  }
  std::unique_lock<std::mutex> lk(mu_);
  done_.wait(lk, [&] { return busy_ == 0; });
}

int Engine::shape_index(const dabgpu_cw_shape &s) {
  for (size_t i = 0; i < shapes.size(); i++)
    if (memcmp(&shapes[i], &s, sizeof s) == 0) return (int)i;
  shapes.push_back(s);
  shapes_dirty = true;
  return (int)shapes.size() - 1;
}

int Engine::init(int n_streams, uint32_t tuner_hz, int flags) {
  int rc;
  if ((rc = ensure_device_ready())) return rc;
  if (n_streams <= 0) {
    set_error(DABGPU_ERR_ARG, "engine: n_streams must be positive");
    return DABGPU_ERR_ARG;
  }
  S = n_streams;
  trace_on = getenv("DABGPU_TRACE") != nullptr;
  f0 = tuner_hz;
  quiet = !(flags & DABGPU_ENGINE_VERBOSE);
  virtual_tuner = flags & DABGPU_ENGINE_VIRTUAL_TUNER;
  soft = flags & DABGPU_ENGINE_SOFT;
  follow_reconfig = flags & DABGPU_ENGINE_FOLLOW_RECONFIG;
  vb_fic.soft = vb_msc.soft = soft;
  subch_mask.assign(S, ~0ull);
  front.assign(S, FrontState());
  back.resize(S);
  layout.assign(S, EnsLayout());
  stats.assign(S, StreamStats());
  for (int s = 0; s < S; s++) {
    front[s].frequency = tuner_hz;
    front[s].rng.seed(1);
    back[s].reset();
    back[s].follow = follow_reconfig;
  }
  {
    const char *env = getenv("DABGPU_HOST_THREADS");
    int nt = env ? atoi(env) : 6;
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0) nt = std::min(nt, std::max(0, hw - 1));
    if (S >= 64 && nt > 0) pool.start(nt);
  }
  if ((rc = d_cifs.reserve((size_t)S * CIF_SLOTS * CIF_BYTES))) return rc;
  if ((rc = d_fibs.reserve((size_t)S * TF_SLOTS * FIBS_PER_TF))) return rc;
  if ((rc = d_ficbits.reserve((size_t)S * 9216))) return rc;
  if ((rc = d_steps_fic.reserve((size_t)S * 4 * FIC_ROW * (soft ? 4 : 1)))) return rc;
  if (soft && (rc = d_cifs_soft.reserve((size_t)S * CIF_SLOTS * 55296))) return rc;
  // ETI output of one call: up to MAX_MSC_BATCH TFs per flush, plus one extra TF for the rare
  // call that has to flush twice
  if ((rc = d_eti.reserve((size_t)S * 4 * (MAX_MSC_BATCH + 1) * DABGPU_ETI_BYTES))) return rc;
  pend_of_stream.assign(S, 0);
  if ((rc = d_ens.reserve((size_t)S * sizeof(EnsDev)))) return rc;
  if ((rc = d_gather_out.reserve((size_t)S * (FIBS_PER_TF + 12)))) return rc;
  for (int i = 0; i < 2; i++)
    if ((rc = h_fic_out[i].reserve((size_t)S * (FIBS_PER_TF + 12)))) return rc;
  frame_slot.assign(S, 0);
  // The FIC chain is small and latency-critical (the host waits for it), the MSC batches are big
  // and nobody waits for them: give the block scheduler that order of preference.
  int prio_lo = 0, prio_hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CUDA_TRY(cudaStreamCreateWithPriority(&st_msc, cudaStreamNonBlocking, prio_lo));
  CUDA_TRY(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithPriority(&st_fic, cudaStreamNonBlocking, prio_hi));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_fic_ready, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_demod_done[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_demod_done[1], cudaEventDisableTiming));
  vb_fic.small_ctas = true;
  for (int i = 0; i < N_STAGE; i++) {
    CUDA_TRY(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_consumed[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&ev_sync_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_sync_out, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_prep_done, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(cudaEventCreateWithFlags(&ev_fic_done[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_jobs[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&ev_ctl[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_ctl[1], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_up[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_up[1], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ev_msc_done, cudaEventDisableTiming));
  CUDA_TRY(cudaMemset(d_cifs.p, 0, (size_t)S * CIF_SLOTS * CIF_BYTES));
  CUDA_TRY(cudaMemset(d_fibs.p, 0, (size_t)S * TF_SLOTS * FIBS_PER_TF));
  dabgpu_cw_shape fs;
  ShapeDev fsd;
  dabgpu_shape_fic(&fs);
  shape_to_dev(fs, &fsd);
  if ((rc = d_fic_shape.reserve(sizeof fsd))) return rc;
  CUDA_TRY(cudaMemcpy(d_fic_shape.p, &fsd, sizeof fsd, cudaMemcpyHostToDevice));
  return DABGPU_OK;
}

// Small control transfers between pinned host memory and the device.  While bulk sample uploads or
// ETI downloads are using the copy engines (host-buffer path) they go by a zero-copy kernel, so
// that they do not queue behind a 268 MB copy; otherwise the copy engines are the cheaper way.
int Engine::ctl_transfer(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  if (bulk_copies) return launch_ctl_copy(dst, src, bytes, st);
  if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, st));
  return DABGPU_OK;
}

void Engine::trace_mark(int k, int which, cudaStream_t st) {
  if (trace.size() > 20000) return;
  if (!trace_base) {
    cudaEventCreate(&trace_base);
    cudaEventRecord(trace_base, st);
  }
  if (which == 0) {
    TraceRec r;
    r.k = k;
    r.host_us = now_us();
    cudaEventCreate(&r.e[0]);
    cudaEventCreate(&r.e[1]);
    cudaEventRecord(r.e[0], st);
    trace.push_back(r);
  } else {
    for (size_t i = trace.size(); i-- > 0;)
      if (trace[i].k == k) {
        cudaEventRecord(trace[i].e[1], st);
        break;
      }
  }
}
void Engine::trace_dump() {
  const char *path = getenv("DABGPU_TRACE");
  if (!path || trace.empty()) return;
  cudaDeviceSynchronize();
  FILE *f = fopen(path, "w");
  if (!f) return;
  static const char *names[K_COUNT] = {"ingest", "fifo", "sync", "demod", "fic_prep", "fic_viterbi",
                                       "msc_gather", "msc_viterbi", "eti_pack"};
  const double h0 = trace[0].host_us;
  for (const TraceRec &r : trace) {
    float a = 0, b = 0;
    if (cudaEventElapsedTime(&a, trace_base, r.e[0]) != cudaSuccess ||
        cudaEventElapsedTime(&b, trace_base, r.e[1]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    fprintf(f, "%-12s %10.4f %10.4f   launched_at_host_ms %10.4f\n", names[r.k], a, b, (r.host_us - h0) / 1e3);
  }
  fclose(f);
}

int Engine::enable_timing(bool on) {
  if (on && !ev[0][0])
    for (int k = 0; k < K_COUNT; k++)
      for (int j = 0; j < 2; j++) CUDA_TRY(cudaEventCreate(&ev[k][j]));
  timing = on;
  for (int k = 0; k < K_COUNT; k++) {
    ms_total[k] = 0;
    n_total[k] = 0;
    ev_used[k] = false;
  }
  return DABGPU_OK;
}

// called at the end of a step when timing is on: wait for the stream, accumulate elapsed times
int Engine::collect_timing(cudaStream_t st) {
  if (!timing) return DABGPU_OK;
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaStreamSynchronize(st_fic));
  CUDA_TRY(cudaStreamSynchronize(st_msc));
  for (int k = 0; k < K_COUNT; k++) {
    if (!ev_used[k]) continue;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev[k][0], ev[k][1]));
    ms_total[k] += ms;
    n_total[k]++;
    ev_used[k] = false;
  }
  return DABGPU_OK;
}

int Engine::join_msc(cudaStream_t user) {
  if (!msc_inflight) return DABGPU_OK;
  CUDA_TRY(cudaStreamWaitEvent(user, ev_msc_done, 0));
  return DABGPU_OK;
}

void Engine::destroy() {
  trace_dump();
  for (TraceRec &r : trace) {
    cudaEventDestroy(r.e[0]);
    cudaEventDestroy(r.e[1]);
  }
  trace.clear();
  pool.stop();
  if (st_fic) {
    cudaStreamSynchronize(st_fic);
    cudaStreamDestroy(st_fic);
    st_fic = nullptr;
    cudaEventDestroy(ev_fic_ready);
    cudaEventDestroy(ev_demod_done[0]);
    cudaEventDestroy(ev_demod_done[1]);
  }
  if (st_copy) {
    cudaStreamSynchronize(st_copy);
    cudaStreamDestroy(st_copy);
    st_copy = nullptr;
    for (int i = 0; i < N_STAGE; i++) {
      cudaEventDestroy(ev_copied[i]);
      cudaEventDestroy(ev_consumed[i]);
      d_stage[i].release();
    }
  }
  if (st_msc) {
    cudaStreamSynchronize(st_msc);
    cudaStreamDestroy(st_msc);
    st_msc = nullptr;
    cudaEventDestroy(ev_up[0]);
    cudaEventDestroy(ev_up[1]);
    cudaEventDestroy(ev_ctl[0]);
    cudaEventDestroy(ev_ctl[1]);
    cudaEventDestroy(ev_sync_done);
    cudaEventDestroy(ev_sync_out);
    cudaEventDestroy(ev_prep_done);
    for (int i = 0; i < 2; i++) {
      cudaEventDestroy(ev_fic_done[i]);
      cudaEventDestroy(ev_jobs[i]);
    }
    cudaEventDestroy(ev_msc_done);
  }
  if (ev[0][0])
    for (int k = 0; k < K_COUNT; k++)
      for (int j = 0; j < 2; j++) cudaEventDestroy(ev[k][j]);
  DevBuf *db[] = {&d_ring, &d_frames, &d_tails, &d_chunk, &d_ctl, &d_sync, &d_cifs, &d_fibs, &d_crc, &d_ficbits,
                  &d_tfbytes, &d_steps_fic, &d_steps_msc, &d_eti, &d_ens, &d_shapes, &d_fic_shape, &d_cifjobs,
                  &d_subjobs, &d_periods, &d_etijobs, &d_planeoff, &d_gather_idx, &d_gather_out, &d_consume, &d_wf_ring, &d_wf_pkts, &d_wf_ctl, &d_cifs_soft};
  for (DevBuf *b : db) b->release();
  PinBuf *pb[] = {&h_ctl, &h_stepctl[0], &h_stepctl[1], &h_sync, &h_fic_out[0], &h_fic_out[1], &h_jobs[0], &h_jobs[1], &h_msc[0], &h_msc[1], &h_eti, &h_chunk};
  for (PinBuf *b : pb) b->release();
  vb_fic.release();
  vb_msc.release();
}

// (re)derive the ETI/MSC layout of stream s from its ens_info (misc.c:153-213, :246-278)
int Engine::refresh_layout(int s) {
  EnsLayout &L = layout[s];
  // (follow mode: the table that was current for the frames' own CIFs, see hostlogic.cuh)
  const bool fol = back[s].follow && back[s].emit_sub;
  const subchannel_info_t *subchans = fol ? back[s].emit_sub : back[s].ens_info.subchans;
  const uint64_t want_version = fol ? back[s].emit_version : back[s].ens_version;
  if (L.version == want_version) return DABGPU_OK;
  L.version = want_version;
  L.epoch = ++layout_epoch;
  L.nsub = 0;
  memset(&L.dev, 0, sizeof L.dev);
  uint32_t row = 0, nst = 0, fl = 0, payload = 0;
  const uint64_t keep = subch_mask[s];  // dabgpu_engine_set_subchannel_mask: all ones by default
  // A sub-channel whose description cannot be decoded (reserved EEP option, size outside the CIF:
  // the reference would index past its tables or read out of bounds here) is left out of this
  // stream's frames and reported through dabgpu_last_error; the other streams of the batched call,
  // and this stream's other sub-channels, carry on.
  dabgpu_cw_shape shp[64];
  bool usable[64];
  for (int j = 0; j < 64; j++) {
    const subchannel_info_t &sc = subchans[j];
    usable[j] = false;
    if (sc.id < 0 || !((keep >> j) & 1)) continue;
    if (host_subch_shape(&sc, &shp[j]) || shp[j].nbits <= 0 || shp[j].nbits > 9216 ||
        sc.start_cu * 64 + shp[j].in_bits > DABGPU_CIF_BITS) {
      set_error(DABGPU_ERR_STATE, "stream %d: sub-channel %d has an undecodable description (skipped)", s, sc.id);
      stats[s].undecodable_subch++;
      continue;
    }
    usable[j] = true;
    nst++;
  }
  L.e1 = 12 + 4 * nst;
  uint32_t e = L.e1 + 96;
  for (int j = 0; j < 64; j++) {
    const subchannel_info_t &sc = subchans[j];
    if (!usable[j]) continue;
    const dabgpu_cw_shape &sh = shp[j];
    EnsLayout::Sub &u = L.sub[L.nsub];
    u.in_bit0 = (uint32_t)sc.start_cu * 64u;
    u.shape = (uint32_t)shape_index(sh);
    u.nbits = (uint32_t)sh.nbits;
    u.row_off = row;
    u.eti_off = e;
    row += vit_row_bytes(u.nbits + 6);
    const int obytes = host_subch_obytes(sh.nbits);
    e += obytes;
    payload += obytes;
    fl += sc.bitrate * 3 / 4;
    const int tpl = sc.slForm == 0 ? (0x10 | (sc.protlev - 1)) : (0x20 | sc.protlev);
    const int stl = sc.bitrate * 3 / 8;
    uint8_t *w = L.dev.stc[L.nsub];
    w[0] = (uint8_t)((sc.id << 2) | ((sc.start_cu >> 8) & 3));
    w[1] = (uint8_t)sc.start_cu;
    w[2] = (uint8_t)((tpl << 2) | ((stl >> 8) & 3));
    w[3] = (uint8_t)stl;
    L.nsub++;
  }
  if (e + 8 > DABGPU_ETI_BYTES) {
    set_error(DABGPU_ERR_STATE, "stream %d: multiplex does not fit an ETI frame", s);
    L.version = 0;
    return DABGPU_ERR_STATE;
  }
  L.rows_bytes = row;
  {
    std::vector<uint32_t> sig;
    for (int u = 0; u < L.nsub; u++) {
      sig.push_back(L.sub[u].in_bit0);
      sig.push_back(L.sub[u].shape);
      sig.push_back(L.sub[u].row_off);
    }
    const LayoutKey *hit = nullptr;
    for (const LayoutKey &k : layout_keys)
      if (k.sig == sig) hit = &k;
    if (!hit) {
      LayoutKey k;
      k.sig = sig;
      k.per0 = (uint32_t)periods.size();
      for (int u = 0; u < L.nsub; u++)
        if (!append_periods(shapes[L.sub[u].shape], L.sub[u].in_bit0, L.sub[u].row_off, periods)) {
          periods.resize(k.per0);
          set_error(DABGPU_ERR_STATE, "stream %d: sub-channel puncturing outside EN 300 401", s);
          L.version = 0;
          return DABGPU_ERR_STATE;
        }
      k.nper = (uint32_t)periods.size() - k.per0;
      layout_keys.push_back(k);
      periods_dirty = true;
      hit = &layout_keys.back();
    }
    L.per0 = hit->per0;
    L.nper = hit->nper;
  }
  L.dev.nst = nst;
  L.dev.fl = fl + nst + 1 + 24;
  L.dev.payload = payload;
  // consumed by eti_pack on the MSC stream; L.dev lives in this object and only changes here
  CUDA_TRY(cudaStreamSynchronize(st_msc));
  CUDA_TRY(cudaMemcpyAsync(d_ens.as<EnsDev>() + s, &L.dev, sizeof(EnsDev), cudaMemcpyHostToDevice, st_msc));
  return DABGPU_OK;
}

int Engine::upload_tables(cudaStream_t st) {
  if (periods_dirty) {
    int rc;
    // (growing the store moves it: the previous batch, which reads it, is complete by now)
    if ((rc = d_periods.reserve(periods.size() * sizeof(PeriodDesc) * 2 + 4096))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_periods.p, periods.data(), periods.size() * sizeof(PeriodDesc), cudaMemcpyHostToDevice,
                             st));
    CUDA_TRY(cudaStreamSynchronize(st));  // `periods` may grow (reallocate) right after
    periods_dirty = false;
  }
  if (!shapes_dirty) return DABGPU_OK;
  std::vector<ShapeDev> sd(shapes.size());
  for (size_t i = 0; i < shapes.size(); i++) shape_to_dev(shapes[i], &sd[i]);
  int rc;
  if ((rc = d_shapes.reserve(sd.size() * sizeof(ShapeDev) + 4096))) return rc;
  CUDA_TRY(cudaMemcpyAsync(d_shapes.p, sd.data(), sd.size() * sizeof(ShapeDev), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // sd is a local
  shapes_dirty = false;
  return DABGPU_OK;
}

// FIC decode of the `active` streams' frames: depuncture, Viterbi, CRC, FIBs into the FIB store,
// compact copy back to the host.  Nothing here waits for the GPU.
// d_fic_src + s*fic_stride holds stream s' 9216 demapped FIC bits (one byte each).
int Engine::fic_launch(cudaStream_t st, const uint8_t *d_fic_src, uint64_t fic_stride, bool early) {
  int rc;
  const int na = (int)active.size();
  if (na == 0) return DABGPU_OK;
  // 4 groups per frame -> compact [na][384] FIBs + [na][12] CRC flags
  uint8_t *d_fib_c = d_gather_out.as<uint8_t>();
  uint8_t *d_crc_c = d_fib_c + (size_t)na * FIBS_PER_TF;
  const size_t idx_bytes = ((size_t)na * 4 + 7) & ~(size_t)7;
  // two staging buffers alternate: with early return the previous frame's upload may not have run yet
  PinBuf &hj = h_jobs[jobs_buf];
  CUDA_TRY(cudaEventSynchronize(ev_jobs[jobs_buf]));
  if ((rc = hj.reserve((size_t)S * 12 + 8))) return rc;
  uint32_t *h_idx = hj.as<uint32_t>();
  uint64_t *h_dst = reinterpret_cast<uint64_t *>(hj.as<uint8_t>() + idx_bytes);
  vb_fic.clear();
  vb_fic.reserve_scale = std::max(1.0, (double)S / na);
  for (int a = 0; a < na; a++) {
    const int s = active[a];
    h_idx[a] = (uint32_t)s;
    h_dst[a] = ((uint64_t)s * TF_SLOTS + (uint64_t)frame_slot[s]) * FIBS_PER_TF;
    for (int k = 0; k < 4; k++)
      vb_fic.add(((uint64_t)a * 4 + k) * FIC_ROW * (soft ? 4u : 1u), (uint64_t)a * FIBS_PER_TF + 96 * k, 768,
                 VIT_DESCRAMBLE);
  }
  if ((rc = d_gather_idx.reserve((size_t)S * 12 + 8))) return rc;
  if ((rc = ctl_transfer(d_gather_idx.p, hj.p, idx_bytes + (size_t)na * 8, cudaMemcpyHostToDevice, st))) return rc;
  CUDA_TRY(cudaEventRecord(ev_jobs[jobs_buf], st));
  jobs_buf ^= 1;
  const uint32_t *d_idx = d_gather_idx.as<uint32_t>();
  const uint64_t *d_dst = reinterpret_cast<const uint64_t *>(d_gather_idx.as<uint8_t>() + idx_bytes);
  t0(K_FIC_PREP, st);
  if (soft) {
    if ((rc = launch_fic_soft_rows(d_fic_src, fic_stride, d_idx, d_steps_fic.as<uint8_t>(), 4 * na,
                                   d_fic_shape.as<ShapeDev>(), st)))
      return rc;
  } else if ((rc = launch_prep_hard(d_fic_src, 2304, 4, fic_stride, d_idx, d_steps_fic.as<uint8_t>(), FIC_ROW, 4 * na,
                                    d_fic_shape.as<ShapeDev>(), 774, st)))
    return rc;
  t1(K_FIC_PREP, st);
  if (early) {  // the next frame's FIC symbols may overwrite d_ficbits from here on
    CUDA_TRY(cudaEventRecord(ev_prep_done, st));
    prep_pending = true;
  }
  t0(K_FIC_VIT, st);
  if ((rc = vb_fic.run(d_steps_fic.as<uint8_t>(), d_fib_c, st))) return rc;
  t1(K_FIC_VIT, st);
  trellis_steps += vb_fic.total_steps;
  if ((rc = launch_fib_crc(d_fib_c, d_crc_c, 12 * na, st))) return rc;
  if ((rc = launch_scatter_rows(d_fib_c, FIBS_PER_TF, d_dst, d_fibs.as<uint8_t>(), na, st))) return rc;
  if ((rc = ctl_transfer(h_fic_out[fic_buf].p, d_fib_c, (size_t)na * (FIBS_PER_TF + 12), cudaMemcpyDeviceToHost, st))) return rc;
  if (early) CUDA_TRY(cudaEventRecord(ev_fic_done[fic_buf], st));
  return DABGPU_OK;
}

// Wait for the FIC results (and the synchroniser outputs) of the frames launched by this call, run
// the part of sdr_demod that feeds back into the next FIFO read, and hand the frames that passed
// to the back-end host logic (`lag`), which runs either right away or during the next call.
// sync: SyncOut per stream (front-end path) or null (demapped path: every active frame is "ok").
int Engine::fic_finish(cudaStream_t st, const SyncOut *sync, int demod_ev, bool early) {
  const int na = (int)active.size();
  if (na == 0) return DABGPU_OK;
  double tw = now_us();
  if (early)
    CUDA_TRY(cudaEventSynchronize(ev_sync_out));  // synchroniser outputs only; the FIC chain runs on
  else
    CUDA_TRY(cudaStreamSynchronize(st));
  host_us[H_WAIT] += now_us() - tw;
  tw = now_us();
  lag.valid = true;
  lag.fic_buf = fic_buf;
  lag.fic_wait = early;
  lag.demod_ev = demod_ev;
  lag.active.swap(active);
  lag.proc.assign(na, 1);
  lag.slot.resize(na);
  for (int a = 0; a < na; a++) {
    const int s = lag.active[a];
    FrontState &fr = front[s];
    fr.pending = false;
    lag.slot[a] = frame_slot[s];
    if (sync) {
      // input_sdr.c:65-112: the order in which sdr_demod updates its state and bails out
      const SyncOut &so = sync[s];
      lag.proc[a] = 0;
      fr.coarse_timeshift = so.coarse_timeshift;
      fr.force_timesync = 0;
      fr.last_ok = 0;
      // A frame that fails here never reaches dab_process_frame: the reference leaves tfidx, the
      // CIF window and its buffers alone (dab2eti.c:68-71, dab.c:97), so the slot taken for it is
      // handed back -- otherwise a run of sync misses would walk the slot ring into CIFs the
      // window (or a queued MSC batch) still refers to.  The demodulator wrote nothing (sync.ok).
      if (so.coarse_timeshift) {
        back[s].phys = frame_slot[s];
        continue;
      }
      fr.fine_timeshift = so.fine_timeshift;
      fr.coarse_freq_shift = so.coarse_freq_shift;
      if (std::abs(so.coarse_freq_shift) > 1) {
        fr.force_timesync = 1;
        back[s].phys = frame_slot[s];
        continue;
      }
      fr.fine_freq_shift = so.fine_freq_shift;
      fr.last_ok = 1;
      lag.proc[a] = 1;
    }
    stats[s].frames_demodulated++;
  }
  fic_buf ^= 1;
  host_us[H_FSM] += now_us() - tw;
  return DABGPU_OK;
}

// dab_process_frame of every frame in `lag` (FIG parsing, lock state machine, CIF window), queueing
// of the resulting ETI frames and, once msc_batch frames per stream are queued, their MSC decode.
int Engine::backend_host(cudaStream_t st) {
  int rc;
  if (!lag.valid) return DABGPU_OK;
  lag.valid = false;
  const int na = (int)lag.active.size();
  double tw = now_us();
  if (lag.fic_wait) {
    CUDA_TRY(cudaEventSynchronize(ev_fic_done[lag.fic_buf]));
    lag.fic_wait = false;
    host_us[H_WAIT] += now_us() - tw;
    tw = now_us();
  }
  // ---- host: per-stream dab_process_frame (streams are independent) ----
  const uint8_t *h_fibs = h_fic_out[lag.fic_buf].as<uint8_t>();
  const uint8_t *h_crc = h_fibs + (size_t)na * FIBS_PER_TF;
  works.resize(na);
  pool.run(na, [&](int a) {
    const int s = lag.active[a];
    FrameWork &work = works[a];
    work.n_eti = 0;
    if (!lag.proc[a]) return;
    for (int i = 0; i < 12; i++) stats[s].fib_crc_errors += h_crc[12 * a + i] ? 0 : 1;
    host_process_frame(back[s], h_fibs + (size_t)a * FIBS_PER_TF, h_crc + 12 * a, lag.slot[a], &work, quiet);
    stats[s].eti_frames += work.n_eti;
  });
  host_us[H_FSM] += now_us() - tw;
  tw = now_us();

  // ---- queue the ETI frames of these transmission frames ----
  bool any = false;
  for (int a = 0; a < na; a++) {
    const FrameWork &work = works[a];
    if (!work.n_eti) continue;
    const int s = lag.active[a];
    if (layout[s].version != (back[s].follow && back[s].emit_sub ? back[s].emit_version : back[s].ens_version)) {
      // the multiplex description changed: frames of this stream that are still queued were
      // produced under the old layout and must be decoded first
      if (pend_of_stream[s]) {
        // jobs of earlier streams from this same frame may already be queued: the flush has to
        // wait for the demodulator that is still writing their newest CIFs
        if (any && lag.demod_ev >= 0) msc_wait_ev = lag.demod_ev;
        if ((rc = flush_msc(st))) return rc;
      }
      if ((rc = refresh_layout(s))) return rc;
    }
    any = true;
    for (int k = 0; k < work.n_eti; k++) {
      pend_stream.push_back(s);
      pend_of_stream[s]++;
      pend_sig = (pend_sig ^ (uint64_t)(uint32_t)s ^ (layout[s].epoch << 32)) * 0x100000001b3ull;
      CifJob cj;
      for (int j = 0; j < 16; j++)
        cj.slot_off[j] = ((uint64_t)s * CIF_SLOTS + (uint64_t)work.win[k][j]) * CIF_BYTES;
      cj.sub0 = cj.nsub = 0;  // filled in at flush time
      cifjobs.push_back(cj);
      EtiJob ej;
      const int w0 = work.win[k][0];
      ej.fib_off = ((uint64_t)s * TF_SLOTS + (uint64_t)(w0 >> 2)) * FIBS_PER_TF + 96u * (uint32_t)(w0 & 3);
      ej.ens = (uint32_t)s;
      ej.cif_hi = work.cif_hi[k];
      ej.cif_lo = work.cif_lo[k];
      ej.pad[0] = ej.pad[1] = 0;
      etijobs.push_back(ej);
    }
  }
  if (any) {
    pend_calls++;
    if (lag.demod_ev >= 0) msc_wait_ev = lag.demod_ev;  // the newest CIFs these frames reference
  }
  host_us[H_JOBS] += now_us() - tw;
  if (pend_calls >= msc_batch) {
    // An MSC batch is best queued right before a stretch in which no front-end kernels will be:
    // it then fills what would be an idle GPU while the host turns around.  If the next callback is
    // expected to complete frames again, the batch may wait for one more transmission frame.
    const bool hold = hold_msc_hint && pend_calls < std::min<int>(msc_batch + 1, MAX_MSC_BATCH);
    if (!hold) return flush_msc(st);
  }
  return DABGPU_OK;
}

// MSC of everything queued: time de-interleave + depuncture gather -> Viterbi + descramble -> ETI
int Engine::flush_msc(cudaStream_t user) {
  int rc;
  (void)user;
  if (etijobs.empty()) {
    pend_calls = 0;
    return DABGPU_OK;
  }
  const double tw = now_us();
  cudaStream_t st = st_msc;
  // bound the lag of the MSC stream: the previous batch must be done before the next one is
  // queued, which keeps every CIF/FIB slot a queued batch references out of the front-end's reach
  if (msc_inflight) CUDA_TRY(cudaEventSynchronize(ev_msc_done));
  // the CIF symbols of the newest queued frame may still be on their way into the CIF store
  if (msc_wait_ev >= 0) {
    CUDA_TRY(cudaStreamWaitEvent(st, ev_demod_done[msc_wait_ev], 0));
    msc_wait_ev = -1;
  }
  // several flushes inside one call (rare: a multiplex change) append to the call's output
  const int base = n_eti, n_new = (int)etijobs.size();
  n_eti = base + n_new;
  eti_stream.insert(eti_stream.end(), pend_stream.begin(), pend_stream.end());
  pend_sig = (pend_sig ^ (uint64_t)base) * 0x100000001b3ull;
  // The per-sub-channel job lists only depend on which streams produced frames and on their
  // multiplex layouts: in the steady state of locked receivers they repeat from flush to flush and
  // the device copies (and the Viterbi plan) are reused as they are.
  const bool reuse = pend_sig == cached_sig && (size_t)n_new == frame_row.size() && n_new > 0;
  if (!reuse) {
    frame_row.clear();
    vb_msc.clear();
    row_base = 0;
    for (int f = 0; f < n_new; f++) {
      const EnsLayout &L = layout[pend_stream[f]];
      frame_row.push_back(row_base);
      for (int u = 0; u < L.nsub; u++)
        vb_msc.add((row_base + L.sub[u].row_off) * (soft ? 4u : 1u),
                   (uint64_t)(base + f) * DABGPU_ETI_BYTES + L.sub[u].eti_off, L.sub[u].nbits, VIT_DESCRAMBLE);
      row_base += L.rows_bytes;
    }
    cached_sig = pend_sig;
  }
  for (int f = 0; f < n_new; f++) {
    const EnsLayout &L = layout[pend_stream[f]];
    cifjobs[f].row_base = frame_row[f];
    cifjobs[f].per0 = L.per0;
    cifjobs[f].nper = L.nper;
  }
  if ((rc = upload_tables(st))) return rc;
  const size_t b_cif = cifjobs.size() * sizeof(CifJob), b_eti = etijobs.size() * sizeof(EtiJob);
  // two pinned staging areas alternate; each is free again once its upload has completed.
  // Every store is sized for a full batch (S streams x msc_batch frames) the first time it is
  // needed, so that no allocation ever happens in the steady state.
  const double scale = std::max(1.0, (double)S * 4.0 * msc_batch / (double)n_new) * 1.02;
  auto full = [scale](size_t bytes) { return (size_t)((double)bytes * scale) + 4096; };
  PinBuf &hm = h_msc[msc_buf];
  CUDA_TRY(cudaEventSynchronize(ev_up[msc_buf]));
  if (hm.cap < b_cif + b_eti && (rc = hm.reserve(full(b_cif + b_eti)))) return rc;
  if (d_cifjobs.cap < b_cif + b_eti && (rc = d_cifjobs.reserve(full(b_cif + b_eti)))) return rc;
  const uint64_t steps_bytes = (row_base + 64) * (soft ? 4u : 1u);
  if (d_steps_msc.cap < steps_bytes && (rc = d_steps_msc.reserve(full(steps_bytes)))) return rc;
  vb_msc.reserve_scale = scale;
  if ((size_t)n_eti * DABGPU_ETI_BYTES > d_eti.cap) {
    set_error(DABGPU_ERR_STATE, "engine: more ETI frames in one call than the output store holds");
    return DABGPU_ERR_STATE;
  }
  uint8_t *hp = hm.as<uint8_t>();
  memcpy(hp, cifjobs.data(), b_cif);
  memcpy(hp + b_cif, etijobs.data(), b_eti);
  if ((rc = ctl_transfer(d_cifjobs.p, hp, b_cif + b_eti, cudaMemcpyHostToDevice, st))) return rc;
  CUDA_TRY(cudaEventRecord(ev_up[msc_buf], st));
  msc_buf ^= 1;
  const CifJob *dj = d_cifjobs.as<CifJob>();
  const EtiJob *de = reinterpret_cast<const EtiJob *>(d_cifjobs.as<uint8_t>() + b_cif);
  host_us[H_JOBS] += now_us() - tw;
  t0(K_MSC_GATHER, st);
  if (soft) {
    if ((rc = launch_msc_soft_gather(d_cifs_soft.as<uint8_t>(), dj, d_periods.as<PeriodDesc>(),
                                     d_steps_msc.as<uint8_t>(), n_new, st)))
      return rc;
  } else if ((rc = launch_msc_gather_periods(d_cifs.as<uint8_t>(), dj, d_periods.as<PeriodDesc>(),
                                             d_steps_msc.as<uint8_t>(), n_new, st)))
    return rc;
  t1(K_MSC_GATHER, st);
  t0(K_MSC_VIT, st);
  if ((rc = reuse ? vb_msc.relaunch(d_steps_msc.as<uint8_t>(), d_eti.as<uint8_t>(), st)
                  : vb_msc.run(d_steps_msc.as<uint8_t>(), d_eti.as<uint8_t>(), st)))
    return rc;
  t1(K_MSC_VIT, st);
  trellis_steps += vb_msc.total_steps;
  t0(K_ETI, st);
  if ((rc = launch_eti_pack(de, d_ens.as<EnsDev>(), d_fibs.as<uint8_t>(),
                            d_eti.as<uint8_t>() + (size_t)base * DABGPU_ETI_BYTES, n_new, st)))
    return rc;
  t1(K_ETI, st);
  CUDA_TRY(cudaEventRecord(ev_msc_done, st));
  msc_inflight = true;
  cifjobs.clear();
  etijobs.clear();
  for (int32_t ps : pend_stream) pend_of_stream[ps] = 0;
  pend_stream.clear();
  pend_sig = 0xcbf29ce484222325ull;
  pend_calls = 0;
  return DABGPU_OK;
}

int Engine::process_demapped(const uint8_t *tfs, size_t pitch, const uint8_t *mask, bool on_device) {
  int rc;
  cudaStream_t st = current_stream();
  if (pitch < 230400) {
    set_error(DABGPU_ERR_ARG, "process_demapped: pitch must be >= 230400");
    return DABGPU_ERR_ARG;
  }
  active.clear();
  for (int s = 0; s < S; s++)
    if (!mask || mask[s]) active.push_back(s);
  const uint8_t *d_tf = tfs;
  if (!on_device) {
    if ((rc = d_tfbytes.reserve((size_t)S * pitch))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_tfbytes.p, tfs, (size_t)S * pitch, cudaMemcpyHostToDevice, st));
    d_tf = d_tfbytes.as<uint8_t>();
  }
  // MSC bytes -> planes in the physical slot that stands for tfs[tfidx] (dab.c:35)
  const int na = (int)active.size();
  n_eti = 0;
  eti_stream.clear();
  hold_msc_hint = false;
  if ((rc = backend_host(st))) return rc;  // a frame left over from a feed_iq call comes first
  if (!na) return DABGPU_OK;
  for (int s : active) take_slot(s);
  if (na == S) {
    if ((rc = h_ctl.reserve((size_t)S * 4 * sizeof(uint64_t)))) return rc;
    if ((rc = d_planeoff.reserve((size_t)S * 4 * sizeof(uint64_t)))) return rc;
    uint64_t *off = h_ctl.as<uint64_t>();
    for (int s = 0; s < S; s++)
      for (int k = 0; k < 4; k++) off[4 * s + k] = ((uint64_t)s * CIF_SLOTS + frame_slot[s] * 4 + k) * CIF_BYTES;
    CUDA_TRY(cudaMemcpyAsync(d_planeoff.p, off, (size_t)S * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if ((rc = soft ? launch_soft_store(d_tf + 9216, pitch, d_planeoff.as<uint64_t>(), d_cifs_soft.as<uint8_t>(), S, st)
                   : launch_pack_planes(d_tf + 9216, pitch, d_planeoff.as<uint64_t>(), d_cifs.as<uint8_t>(), S, st)))
      return rc;
  } else {
    // masked call: pack stream by stream (not a throughput path)
    if ((rc = h_ctl.reserve(4 * sizeof(uint64_t) * (size_t)na))) return rc;
    if ((rc = d_planeoff.reserve(4 * sizeof(uint64_t) * (size_t)na))) return rc;
    uint64_t *off = h_ctl.as<uint64_t>();
    for (int a = 0; a < na; a++)
      for (int k = 0; k < 4; k++)
        off[4 * a + k] = ((uint64_t)active[a] * CIF_SLOTS + frame_slot[active[a]] * 4 + k) * CIF_BYTES;
    CUDA_TRY(cudaMemcpyAsync(d_planeoff.p, off, 4 * sizeof(uint64_t) * (size_t)na, cudaMemcpyHostToDevice, st));
    for (int a = 0; a < na; a++)
      if ((rc = soft ? launch_soft_store(d_tf + (size_t)active[a] * pitch + 9216, pitch,
                                         d_planeoff.as<uint64_t>() + 4 * a, d_cifs_soft.as<uint8_t>(), 1, st)
                     : launch_pack_planes(d_tf + (size_t)active[a] * pitch + 9216, pitch,
                                          d_planeoff.as<uint64_t>() + 4 * a, d_cifs.as<uint8_t>(), 1, st)))
        return rc;
  }
  if ((rc = fic_launch(st, d_tf, pitch, false))) return rc;
  if ((rc = fic_finish(st, nullptr, -1, false))) return rc;  // also covers the pack kernels on `st`
  if ((rc = backend_host(st))) return rc;
  return collect_timing(st);
}

// do_wf_decode's loop body (dab2eti.c:267-270) for every stream with packets: wf_read_frame into
// dab->tfs[dab->tfidx] (input_wf.c:65-115), then dab_process_frame.  Symbols that did not arrive
// keep what the slot held five frames ago, like the reference's ring of five frame buffers.
int Engine::process_wavefinder(const uint8_t *packets, size_t pitch, const int32_t *n_packets) {
  int rc;
  cudaStream_t st = current_stream();
  int maxp = 0;
  for (int s = 0; s < S; s++) {
    if (n_packets[s] < 0 || (size_t)n_packets[s] * 524u > pitch) {
      set_error(DABGPU_ERR_ARG, "process_wavefinder: stream %d: %d packets do not fit the pitch", s, n_packets[s]);
      return DABGPU_ERR_ARG;
    }
    maxp = std::max(maxp, n_packets[s]);
  }
  if (!d_wf_ring.p) {
    if ((rc = d_wf_ring.reserve((size_t)S * 5 * 230400))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_wf_ring.p, 0, (size_t)S * 5 * 230400, st));  // init_dab_state calloc()s tfs
  }
  if ((rc = d_wf_pkts.reserve((size_t)S * pitch + 16))) return rc;
  if ((rc = d_wf_ctl.reserve((size_t)S * 12 + 16))) return rc;
  if ((rc = d_tfbytes.reserve((size_t)S * 230400))) return rc;
  if ((rc = h_chunk.reserve((size_t)S * 9))) return rc;
  int32_t *h = h_chunk.as<int32_t>();
  uint8_t *mask = h_chunk.as<uint8_t>() + (size_t)S * 8;
  CUDA_TRY(cudaStreamSynchronize(st));  // the staging below is reused from call to call
  for (int s = 0; s < S; s++) {
    h[s] = n_packets[s];
    h[S + s] = back[s].tfidx;
    mask[s] = n_packets[s] > 0;
  }
  int32_t *d_n = d_wf_ctl.as<int32_t>(), *d_slot = d_n + S;
  uint32_t *d_seen = reinterpret_cast<uint32_t *>(d_slot + S);
  CUDA_TRY(cudaMemcpyAsync(d_n, h, (size_t)S * 8, cudaMemcpyHostToDevice, st));
  if (maxp) CUDA_TRY(cudaMemcpyAsync(d_wf_pkts.p, packets, (size_t)S * pitch, cudaMemcpyHostToDevice, st));
  if ((rc = launch_wf_demap(d_wf_pkts.as<uint8_t>(), pitch, d_n, maxp, d_slot, d_wf_ring.as<uint8_t>(), d_seen,
                            d_tfbytes.as<uint8_t>(), S, st)))
    return rc;
  return process_demapped(d_tfbytes.as<uint8_t>(), 230400, mask, true);
}

int Engine::ensure_frontend() {
  int rc;
  if (d_ctl.p) return DABGPU_OK;
  if (!capture_base) {
    if ((rc = d_ring.reserve((size_t)S * IQ_RING_BYTES))) return rc;
    CUDA_TRY(cudaMemset(d_ring.p, 0, (size_t)S * IQ_RING_BYTES));
    rg = RingGeom{d_ring.as<uint8_t>(), IQ_RING_BYTES, IQ_RING_BYTES};
  }
  if ((rc = d_frames.reserve((size_t)S * DABGPU_TF_BYTES))) return rc;
  if ((rc = d_tails.reserve((size_t)S * TAIL_BYTES))) return rc;
  if ((rc = d_ctl.reserve((size_t)S * sizeof(StepCtl)))) return rc;
  if ((rc = d_sync.reserve((size_t)S * sizeof(SyncOut)))) return rc;
  if ((rc = h_sync.reserve((size_t)S * sizeof(SyncOut)))) return rc;
  CUDA_TRY(cudaMemset(d_frames.p, 0, (size_t)S * DABGPU_TF_BYTES));
  CUDA_TRY(cudaMemset(d_tails.p, 0, (size_t)S * TAIL_BYTES));
  CUDA_TRY(cudaMemset(d_sync.p, 0, (size_t)S * sizeof(SyncOut)));
  return DABGPU_OK;
}

// dab2eti.c:75-103, applied after every callback
static void tuner_feedback(FrontState &fr) {
  const int cfs = fr.coarse_freq_shift;
  if (std::abs(cfs) > 1) fr.frequency = cfs < 0 ? fr.frequency - 1000u : fr.frequency + 1000u;
  if (std::abs(cfs) == 1) {
    const uint32_t d = (uint32_t)(fr.rng.next() % 1000);
    fr.frequency = cfs < 0 ? fr.frequency - d : fr.frequency + d;
  }
  if (std::abs(cfs) < 1 && std::abs((int)fr.fine_freq_shift) > 50)
    fr.frequency = (uint32_t)((double)fr.frequency + fr.fine_freq_shift / 3);
}

// upload one callback's worth of host IQ for every stream; returns immediately
int Engine::submit_iq(const uint8_t *iq, size_t pitch, int chunk_len) {
  int rc;
  if (chunk_len <= 0 || chunk_len > 262144 || (chunk_len & 15) || pitch < (size_t)chunk_len) {
    set_error(DABGPU_ERR_ARG, "submit_iq: chunk_len must be a multiple of 16 in (0, 262144], pitch >= chunk_len");
    return DABGPU_ERR_ARG;
  }
  if (stage_count >= N_STAGE) {
    set_error(DABGPU_ERR_STATE, "submit_iq: %d chunks are already queued; call dabgpu_engine_feed_submitted", N_STAGE);
    return DABGPU_ERR_STATE;
  }
  bulk_copies = true;
  const int b = (stage_head + stage_count) % N_STAGE;
  if ((rc = d_stage[b].reserve((size_t)S * 262144))) return rc;
  // the buffer is free once the ingest kernel that read it last has run
  CUDA_TRY(cudaStreamWaitEvent(st_copy, ev_consumed[b], 0));
  if (pitch == (size_t)chunk_len)
    CUDA_TRY(cudaMemcpyAsync(d_stage[b].p, iq, (size_t)S * chunk_len, cudaMemcpyHostToDevice, st_copy));
  else
    CUDA_TRY(cudaMemcpy2DAsync(d_stage[b].p, chunk_len, iq, pitch, chunk_len, S, cudaMemcpyHostToDevice, st_copy));
  CUDA_TRY(cudaEventRecord(ev_copied[b], st_copy));
  stage_len[b] = chunk_len;
  stage_seq[b] = ++submit_seq;
  stage_count++;
  return DABGPU_OK;
}

// copies complete in submission order (one copy stream): the newest stage whose event has fired
// bounds what is still on its way
int Engine::uploads_in_flight() {
  uint64_t done = submit_seq > N_STAGE ? submit_seq - N_STAGE : 0;  // older stages were re-used: their copies are done
  for (int b = 0; b < N_STAGE; b++)
    if (stage_seq[b] > done && cudaEventQuery(ev_copied[b]) == cudaSuccess) done = stage_seq[b];
  return (int)(submit_seq - done);
}

int Engine::feed_submitted() {
  if (stage_count <= 0) {
    set_error(DABGPU_ERR_STATE, "feed_submitted: nothing was submitted");
    return DABGPU_ERR_STATE;
  }
  const int b = stage_head;
  CUDA_TRY(cudaStreamWaitEvent(current_stream(), ev_copied[b], 0));
  consuming_stage = b;
  const int rc = feed_iq(d_stage[b].as<uint8_t>(), (size_t)stage_len[b], stage_len[b], true);
  consuming_stage = -1;
  stage_head = (stage_head + 1) % N_STAGE;
  stage_count--;
  return rc;
}

int Engine::attach_capture(const uint8_t *iq_device, size_t pitch, size_t len) {
  if (d_ctl.p || capture_base || stage_count) {
    set_error(DABGPU_ERR_STATE, "attach_capture: must come before the first samples are fed");
    return DABGPU_ERR_STATE;
  }
  if (virtual_tuner) {
    set_error(DABGPU_ERR_STATE, "attach_capture: the virtual tuner rewrites the samples and needs the copying path");
    return DABGPU_ERR_STATE;
  }
  if (!iq_device || ((uintptr_t)iq_device & 15) || (pitch & 15) || (len & 15) || pitch < len ||
      len < 2u * DABGPU_TF_BYTES || len > 0x7fffffffu) {
    set_error(DABGPU_ERR_ARG,
              "attach_capture: pointer, pitch and length must be multiples of 16, pitch >= length, "
              "2 transmission frames <= length < 2 GiB");
    return DABGPU_ERR_ARG;
  }
  capture_base = iq_device;
  capture_len = len;
  capture_fed = 0;
  rg = RingGeom{iq_device, pitch, (uint32_t)len};
  return DABGPU_OK;
}

int Engine::feed_capture(int chunk_len) {
  if (!capture_base) {
    set_error(DABGPU_ERR_STATE, "feed_capture: no capture attached");
    return DABGPU_ERR_STATE;
  }
  if (chunk_len > 0 && !capture_cyclic && capture_fed + (uint64_t)chunk_len > capture_len) {
    set_error(DABGPU_ERR_STATE, "feed_capture: the capture holds %llu more bytes per stream, %d requested",
              (unsigned long long)(capture_len - capture_fed), chunk_len);
    return DABGPU_ERR_STATE;
  }
  capture_call = true;
  const int rc = feed_iq(capture_base, (size_t)rg.pitch, chunk_len, true);
  capture_call = false;
  if (rc == DABGPU_OK) capture_fed += (uint64_t)chunk_len;
  return rc;
}

int Engine::feed_iq(const uint8_t *iq, size_t pitch, int chunk_len, bool on_device) {
  int rc;
  if (capture_base && !capture_call) {
    set_error(DABGPU_ERR_STATE, "feed_iq: this engine consumes an attached capture; use dabgpu_engine_feed_capture");
    return DABGPU_ERR_STATE;
  }
  if (!on_device) {
    if ((rc = submit_iq(iq, pitch, chunk_len))) return rc;
    // chunks submitted earlier are consumed first (FIFO); this call consumes exactly one
    return feed_submitted();
  }
  if (chunk_len <= 0 || chunk_len > 262144 || (chunk_len & 15) || pitch < (size_t)chunk_len) {
    set_error(DABGPU_ERR_ARG, "feed_iq: chunk_len must be a multiple of 16 in (0, 262144], pitch >= chunk_len");
    return DABGPU_ERR_ARG;
  }
  if ((rc = ensure_frontend())) return rc;
  const double t_pre = now_us();
  cudaStream_t st = current_stream();
  // the staging copy of the call before last has certainly been consumed; make sure anyway
  PinBuf &hc = h_stepctl[ctl_buf];
  CUDA_TRY(cudaEventSynchronize(ev_ctl[ctl_buf]));
  if ((rc = hc.reserve((size_t)S * sizeof(StepCtl)))) return rc;
  StepCtl *ctl = hc.as<StepCtl>();
  active.clear();
  bool any_read = false, any_copy = false, any_mat = false;
  for (int s = 0; s < S; s++) {
    FrontState &fr = front[s];
    StepCtl &c = ctl[s];
    memset(&c, 0, sizeof c);
    fr.coarse_freq_shift = 0;  // input_sdr.c:33
    fr.last_ok = 0;
    // cbWrite x chunk_len (input_sdr.c:36-38)
    if (fr.fifo_count + (uint32_t)chunk_len > IQ_RING_BYTES) {
      set_error(DABGPU_ERR_STATE, "stream %d: FIFO overflow", s);
      return DABGPU_ERR_STATE;
    }
    c.wr_pos = (fr.fifo_start + fr.fifo_count) % rg.mod;
    c.nco_hz = virtual_tuner ? (int32_t)(fr.frequency - f0) : 0;
    c.nco_sample0 = fr.samples_in;
    fr.samples_in += (uint64_t)chunk_len / 2;
    fr.fifo_count += (uint32_t)chunk_len;
    if (fr.fifo_count < 196608u * 3u) continue;  // input_sdr.c:41-43
    // sdr_read_fifo(fifo, 393216, coarse+fine shift, buffer) (sdr_fifo.c:43-61)
    const int32_t shift = fr.coarse_timeshift + fr.fine_timeshift;
    const uint32_t bytes = DABGPU_TF_BYTES;
    if (shift > 0) {
      const uint32_t skip = std::min<uint32_t>((uint32_t)shift, fr.fifo_count);
      const uint32_t skip_pos = fr.fifo_start;
      fr.fifo_start = (fr.fifo_start + skip) % rg.mod;
      fr.fifo_count -= skip;
      const uint32_t n = std::min(bytes, fr.fifo_count);
      c.rd_pos[0] = fr.fifo_start;
      c.rd_dst[0] = 0;
      c.rd_bytes[0] = n;
      fr.fifo_start = (fr.fifo_start + n) % rg.mod;
      fr.fifo_count -= n;
      // the skipped bytes were parked in buffer[0..skip) and are only overwritten up to n
      const uint32_t lim = std::min(skip, bytes);
      if (n < lim) {
        c.rd_pos[1] = (skip_pos + n) % rg.mod;
        c.rd_dst[1] = n;
        c.rd_bytes[1] = lim - n;
      }
    } else {
      const uint32_t n = bytes + (uint32_t)shift;  // stale tail of -shift bytes is kept
      c.rd_pos[0] = fr.fifo_start;
      c.rd_dst[0] = 0;
      c.rd_bytes[0] = n;
      fr.fifo_start = (fr.fifo_start + n) % rg.mod;
      fr.fifo_count -= n;
    }
    any_read = true;
    // a plain window is used in place; anything else goes through the frame buffer
    if (c.rd_bytes[1] == 0 && c.rd_dst[0] == 0 && c.rd_bytes[0] >= TAIL_OFF) {
      c.src_ring = 1;
      c.src_pos = c.rd_pos[0];
      fr.prev_ring = true;
      fr.prev_pos = c.rd_pos[0];
    } else {
      c.src_ring = 0;
      if (fr.prev_ring) {
        c.mat = 1;
        c.mat_pos = fr.prev_pos;
        any_mat = true;
      }
      fr.prev_ring = false;
      any_copy = true;
    }
    if (fr.startup_delay <= 0) {  // GAIN_SETTLE_TIME == 0: the first frame is read and dropped
      fr.startup_delay++;
      if (!quiet) fprintf(stderr, "startup_delay=%i\n", fr.startup_delay);
      continue;
    }
    fr.coarse_timeshift = 0;  // input_sdr.c:66
    c.run = 1;
    c.force_timesync = fr.force_timesync;
    const int slot = take_slot(s);
    for (int k = 0; k < 4; k++)
      c.cif_off[k] = ((uint64_t)s * CIF_SLOTS + (uint64_t)slot * 4 + k) * CIF_BYTES;
    fr.pending = true;
    active.push_back(s);
  }
  {
    // will the next callback (same size) complete frames?  (input_sdr.c:41-43)
    int next_frames = 0;
    for (int s = 0; s < S; s++) next_frames += front[s].fifo_count + (uint32_t)chunk_len >= 196608u * 3u;
    hold_msc_hint = trailing_hint_enabled() && 2 * next_frames >= S;
  }
  if ((rc = ctl_transfer(d_ctl.p, ctl, (size_t)S * sizeof(StepCtl), cudaMemcpyHostToDevice, st))) return rc;
  CUDA_TRY(cudaEventRecord(ev_ctl[ctl_buf], st));
  ctl_buf ^= 1;
  host_us[H_PRE] += now_us() - t_pre;
  if (!capture_base) {  // (an attached capture is consumed in place)
    t0(K_INGEST, st);
    if ((rc = launch_ingest(iq, pitch, (uint32_t)chunk_len, d_ring.as<uint8_t>(), d_ctl.as<StepCtl>(), S, st)))
      return rc;
    t1(K_INGEST, st);
  }
  if (consuming_stage >= 0) CUDA_TRY(cudaEventRecord(ev_consumed[consuming_stage], st));
  n_eti = 0;
  eti_stream.clear();
  // With deferred MSC batches the back-end host logic trails the front-end by one frame: it runs
  // below, after this call's kernels have been queued, so the GPU is never waiting for it.
  const bool trailing = msc_batch > 1 && !timing;  // per-kernel timing runs everything serially
  bool launched = false;
  int demod_ev = -1;
  if (any_read) {
    t0(K_FIFO, st);
    if (any_mat && (rc = launch_fifo_read(rg, d_tails.as<uint8_t>(), d_frames.as<uint8_t>(),
                                          d_ctl.as<StepCtl>(), S, true, st)))
      return rc;
    if (any_copy && (rc = launch_fifo_read(rg, d_tails.as<uint8_t>(), d_frames.as<uint8_t>(),
                                           d_ctl.as<StepCtl>(), S, false, st)))
      return rc;
    if ((rc = launch_tail_update(rg, d_frames.as<uint8_t>(), d_tails.as<uint8_t>(),
                                 d_ctl.as<StepCtl>(), S, st)))
      return rc;
    t1(K_FIFO, st);
    if (!active.empty()) {
      t0(K_SYNC, st);
      if ((rc = launch_sync(rg, d_tails.as<uint8_t>(), d_frames.as<uint8_t>(),
                            d_ctl.as<StepCtl>(), d_sync.as<SyncOut>(), S, st)))
        return rc;
      t1(K_SYNC, st);
      // With trailing back-end logic the call returns as soon as the synchroniser outputs are on
      // the host (they are all the next FIFO read depends on): copy them out right after the
      // synchroniser, ahead of the demodulator.
      if (trailing) {
        CUDA_TRY(cudaEventRecord(ev_sync_done, st));
        CUDA_TRY(cudaStreamWaitEvent(st_fic, ev_sync_done, 0));
        if ((rc = ctl_transfer(h_sync.p, d_sync.p, (size_t)S * sizeof(SyncOut), cudaMemcpyDeviceToHost, st_fic)))
          return rc;
        CUDA_TRY(cudaEventRecord(ev_sync_out, st_fic));
        if (prep_pending) {  // the previous frame's FIC depuncture still reads d_ficbits
          CUDA_TRY(cudaStreamWaitEvent(st, ev_prep_done, 0));
          prep_pending = false;
        }
      }
      // FIC symbols first; their decoding then runs on st_fic next to the CIF symbols on `st`
      t0(K_DEMOD, st);
      if ((rc = launch_demod(rg, d_tails.as<uint8_t>(), d_frames.as<uint8_t>(),
                             d_ctl.as<StepCtl>(), d_sync.as<SyncOut>(), d_ficbits.as<uint8_t>(),
                             soft ? d_cifs_soft.as<uint8_t>() : d_cifs.as<uint8_t>(), S, 0, 1, soft, st)))
        return rc;
      CUDA_TRY(cudaEventRecord(ev_fic_ready, st));
      if ((rc = launch_demod(rg, d_tails.as<uint8_t>(), d_frames.as<uint8_t>(),
                             d_ctl.as<StepCtl>(), d_sync.as<SyncOut>(), d_ficbits.as<uint8_t>(),
                             soft ? d_cifs_soft.as<uint8_t>() : d_cifs.as<uint8_t>(), S, 1, 4, soft, st)))
        return rc;
      t1(K_DEMOD, st);
      demod_ev = demod_ev_next;
      demod_ev_next ^= 1;
      CUDA_TRY(cudaEventRecord(ev_demod_done[demod_ev], st));
      // (when kernels are being timed, each one runs alone: the FIC chain starts after the CIFs)
      CUDA_TRY(cudaStreamWaitEvent(st_fic, timing ? ev_demod_done[demod_ev] : ev_fic_ready, 0));
      if (!trailing &&
          (rc = ctl_transfer(h_sync.p, d_sync.p, (size_t)S * sizeof(SyncOut), cudaMemcpyDeviceToHost, st_fic)))
        return rc;
      if ((rc = fic_launch(st_fic, d_ficbits.as<uint8_t>(), 9216, trailing))) return rc;
      launched = true;
    }
  }
  // The previous frame's dab_process_frame runs here, while the kernels queued above execute.  A
  // callback that completed no frame has queued nothing to hide it behind: leave it for the next one.
  if ((launched || !trailing) && (rc = backend_host(st))) return rc;
  if (launched) {
    if ((rc = fic_finish(st_fic, h_sync.as<SyncOut>(), demod_ev, trailing))) return rc;
    if (!trailing && (rc = backend_host(st))) return rc;
  }
  for (int s = 0; s < S; s++) tuner_feedback(front[s]);
  return collect_timing(st);
}

}  // namespace dabgpu

// =============================== C ABI ===========================================================
using namespace dabgpu;

struct dabgpu_engine {
  Engine e;
};

DABGPU_EXPORT int dabgpu_engine_create(dabgpu_engine **out, int n_streams, uint32_t tuner_hz, int flags) {
  if (!out) return DABGPU_ERR_ARG;
  dabgpu_engine *h = new dabgpu_engine();
  int rc = h->e.init(n_streams, tuner_hz, flags);
  if (rc) {
    h->e.destroy();
    delete h;
    *out = nullptr;
    return rc;
  }
  *out = h;
  return DABGPU_OK;
}
DABGPU_EXPORT void dabgpu_engine_destroy(dabgpu_engine *h) {
  if (!h) return;
  cudaDeviceSynchronize();
  h->e.destroy();
  delete h;
}
DABGPU_EXPORT int dabgpu_engine_feed_iq(dabgpu_engine *h, const uint8_t *iq, size_t pitch, int chunk_len,
                                        int on_device) {
  return h->e.feed_iq(iq, pitch, chunk_len, on_device != 0);
}
DABGPU_EXPORT int dabgpu_engine_submit_iq(dabgpu_engine *h, const uint8_t *iq, size_t pitch, int chunk_len) {
  return h->e.submit_iq(iq, pitch, chunk_len);
}
DABGPU_EXPORT int dabgpu_engine_feed_submitted(dabgpu_engine *h) { return h->e.feed_submitted(); }
DABGPU_EXPORT int dabgpu_engine_uploads_in_flight(dabgpu_engine *h) { return h->e.uploads_in_flight(); }
DABGPU_EXPORT int dabgpu_engine_attach_capture(dabgpu_engine *h, const uint8_t *iq_device, size_t pitch, size_t len) {
  return h->e.attach_capture(iq_device, pitch, len);
}
DABGPU_EXPORT int dabgpu_engine_feed_capture(dabgpu_engine *h, int chunk_len) { return h->e.feed_capture(chunk_len); }
DABGPU_EXPORT int dabgpu_engine_set_capture_cyclic(dabgpu_engine *h, int on) {
  h->e.capture_cyclic = on != 0;
  return DABGPU_OK;
}
DABGPU_EXPORT int dabgpu_engine_process_demapped(dabgpu_engine *h, const uint8_t *tfs, size_t pitch,
                                                 const uint8_t *mask, int on_device) {
  return h->e.process_demapped(tfs, pitch, mask, on_device != 0);
}
DABGPU_EXPORT int dabgpu_engine_eti_count(dabgpu_engine *h) { return h->e.n_eti; }
DABGPU_EXPORT const uint8_t *dabgpu_engine_eti_device(dabgpu_engine *h) { return h->e.d_eti.as<uint8_t>(); }
DABGPU_EXPORT int dabgpu_engine_fetch_eti(dabgpu_engine *h, uint8_t *eti, int32_t *stream_ids, int max_frames) {
  Engine &e = h->e;
  const int n = std::min(e.n_eti, max_frames);
  if (n <= 0) return 0;
  cudaStream_t st = e.st_msc;
  if (eti) {
    if ((size_t)n * DABGPU_ETI_BYTES > (4u << 20)) e.bulk_copies = true;  // see Engine::ctl_transfer
    cudaError_t err = cudaMemcpyAsync(eti, e.d_eti.p, (size_t)n * DABGPU_ETI_BYTES, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
      set_error(DABGPU_ERR_CUDA, "fetch_eti: %s", cudaGetErrorString(err));
      return DABGPU_ERR_CUDA;
    }
  }
  if (stream_ids) memcpy(stream_ids, e.eti_stream.data(), (size_t)n * sizeof(int32_t));
  return n;
}
// ETI consumers on the frames of the last call, where they lie in HBM (queued on the MSC stream,
// behind the kernels that produce the frames)
DABGPU_EXPORT int dabgpu_engine_extract_subchannel(dabgpu_engine *h, int subchid, uint8_t *out_host, size_t out_pitch,
                                                   int32_t *out_len_host) {
  Engine &e = h->e;
  int rc;
  const int n = e.n_eti;
  if (n <= 0) return 0;
  if (!out_host || !out_len_host || subchid < 0 || subchid > 63 || out_pitch == 0) {
    set_error(DABGPU_ERR_ARG, "engine_extract_subchannel: null pointer, SubChId outside 0..63 or zero pitch");
    return DABGPU_ERR_ARG;
  }
  cudaStream_t st = e.st_msc;
  const size_t ob = (size_t)n * out_pitch;
  if ((rc = e.d_consume.reserve(ob + (size_t)n * 4 + 64))) return rc;
  uint8_t *d_out = e.d_consume.as<uint8_t>();
  int32_t *d_len = reinterpret_cast<int32_t *>(d_out + ((ob + 15) & ~(size_t)15));
  CUDA_TRY(cudaMemsetAsync(d_out, 0, ob, st));
  if ((rc = launch_eti_extract(e.d_eti.as<uint8_t>(), n, subchid, d_out, out_pitch, d_len, st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_host, d_out, ob, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_len_host, d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return n;
}
DABGPU_EXPORT int dabgpu_engine_check_eti(dabgpu_engine *h, uint32_t *flags_host) {
  Engine &e = h->e;
  int rc;
  const int n = e.n_eti;
  if (n <= 0) return 0;
  if (!flags_host) {
    set_error(DABGPU_ERR_ARG, "engine_check_eti: null pointer");
    return DABGPU_ERR_ARG;
  }
  cudaStream_t st = e.st_msc;
  if ((rc = e.d_consume.reserve((size_t)n * 4 + 64))) return rc;
  if ((rc = launch_eti_check(e.d_eti.as<uint8_t>(), n, e.d_consume.as<uint32_t>(), st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(flags_host, e.d_consume.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return n;
}
DABGPU_EXPORT int dabgpu_engine_status(dabgpu_engine *h, int stream, dabgpu_stream_status *out) {
  Engine &e = h->e;
  if (stream < 0 || stream >= e.S || !out) {
    set_error(DABGPU_ERR_ARG, "engine_status: bad stream index");
    return DABGPU_ERR_ARG;
  }
  const FrontState &fr = e.front[stream];
  const BackendState &bk = e.back[stream];
  out->locked = bk.locked;
  out->okcount = bk.okcount;
  out->ncifs = bk.ncifs;
  out->tfidx = bk.tfidx;
  out->coarse_timeshift = fr.coarse_timeshift;
  out->fine_timeshift = fr.fine_timeshift;
  out->coarse_freq_shift = fr.coarse_freq_shift;
  out->last_ok = fr.last_ok;
  out->fine_freq_shift = fr.fine_freq_shift;
  out->frequency = fr.frequency;
  out->n_subchannels = 0;
  for (int i = 0; i < 64; i++) out->n_subchannels += bk.ens_info.subchans[i].id >= 0;
  out->frames_demodulated = e.stats[stream].frames_demodulated;
  out->eti_frames = e.stats[stream].eti_frames;
  out->fib_crc_errors = e.stats[stream].fib_crc_errors;
  return DABGPU_OK;
}
DABGPU_EXPORT int dabgpu_engine_set_seed(dabgpu_engine *h, int stream, unsigned seed) {
  if (stream < 0 || stream >= h->e.S) return DABGPU_ERR_ARG;
  h->e.front[stream].rng.seed(seed);
  return DABGPU_OK;
}
DABGPU_EXPORT uint64_t dabgpu_engine_trellis_steps(dabgpu_engine *h) { return h->e.trellis_steps; }

DABGPU_EXPORT int dabgpu_engine_enable_timing(dabgpu_engine *h, int on) { return h->e.enable_timing(on != 0); }
DABGPU_EXPORT int dabgpu_engine_kernel_times(dabgpu_engine *h, double *ms_total, uint64_t *launches, int n) {
  for (int k = 0; k < n && k < Engine::K_COUNT; k++) {
    ms_total[k] = h->e.ms_total[k];
    launches[k] = h->e.n_total[k];
  }
  return Engine::K_COUNT;
}

DABGPU_EXPORT void dabgpu_engine_host_times(dabgpu_engine *h, double *us4) {
  for (int i = 0; i < Engine::H_COUNT; i++) us4[i] = h->e.host_us[i];
}

DABGPU_EXPORT int dabgpu_engine_set_subchannel_mask(dabgpu_engine *h, int stream, uint64_t mask) {
  Engine &e = h->e;
  if (stream < -1 || stream >= e.S) {
    set_error(DABGPU_ERR_ARG, "set_subchannel_mask: bad stream index");
    return DABGPU_ERR_ARG;
  }
  for (int s = stream < 0 ? 0 : stream; s < (stream < 0 ? e.S : stream + 1); s++)
    if (e.subch_mask[s] != mask) {
      e.subch_mask[s] = mask;
      e.layout[s].version = 0;  // re-derived (after the frames already queued) with the next frame
    }
  return DABGPU_OK;
}

DABGPU_EXPORT int dabgpu_engine_process_wavefinder(dabgpu_engine *h, const uint8_t *packets, size_t pitch,
                                                   const int32_t *n_packets) {
  if (!packets || !n_packets) {
    set_error(DABGPU_ERR_ARG, "process_wavefinder: null pointer");
    return DABGPU_ERR_ARG;
  }
  return h->e.process_wavefinder(packets, pitch, n_packets);
}
DABGPU_EXPORT int dabgpu_engine_set_msc_batch(dabgpu_engine *h, int calls) {
  if (calls < 1 || calls > MAX_MSC_BATCH) {
    set_error(DABGPU_ERR_ARG, "msc batch depth must be 1..%d transmission frames", MAX_MSC_BATCH);
    return DABGPU_ERR_ARG;
  }
  h->e.msc_batch = calls;
  return DABGPU_OK;
}
DABGPU_EXPORT int dabgpu_engine_flush(dabgpu_engine *h) {
  h->e.n_eti = 0;
  h->e.eti_stream.clear();
  h->e.hold_msc_hint = false;
  int rc = h->e.backend_host(current_stream());  // a frame whose host logic is still outstanding
  if (rc) return rc;
  rc = h->e.flush_msc(current_stream());
  if (rc) return rc;
  return h->e.collect_timing(current_stream());
}

/* make the caller's stream (dabgpu_set_stream) wait for the MSC/ETI work issued so far, e.g.
 * before recording an event that should cover it or before reading dabgpu_engine_eti_device() */
DABGPU_EXPORT int dabgpu_engine_join(dabgpu_engine *h) { return h->e.join_msc(current_stream()); }
