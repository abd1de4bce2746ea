// pump.cu -- streaming ingest front: the batched counterpart of dab2eti's
// rtlsdr_read_async + demod_thread_fn pair (dab2eti.c:60-135, 237-239) for S live or recorded sources.
//
// dabgpu_engine_pump() reads 262144-byte callbacks' worth of uint8 I/Q from one file descriptor per
// stream (a recording, a FIFO, the stdout of `rtl_sdr -`), keeps the engine fed through
// dabgpu_engine_submit_iq / feed_submitted with the uploads two callbacks ahead, and writes every
// ETI frame to the stream's output descriptor, as dab2eti's eti_callback does (dab2eti.c:132-135).
// A reader thread fills pinned host buffers while the GPU works on the previous ones (the role of
// librtlsdr's asynchronous transfer thread); blocking reads pace a live source by themselves.
// Only the public engine API of include/dabgpu.h is used here.
#include <errno.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

using namespace dabgpu;

namespace {

enum { CHUNK = 262144, N_BUF = 5 };  // 3 may be in flight inside the engine, 1 being read, 1 spare

bool read_full(int fd, uint8_t *dst, size_t n) {
  size_t got = 0;
  while (got < n) {
    const ssize_t r = read(fd, dst + got, n - got);
    if (r == 0) return false;  // end of the source
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    got += (size_t)r;
  }
  return true;
}
bool write_full(int fd, const uint8_t *src, size_t n) {
  size_t put = 0;
  while (put < n) {
    const ssize_t r = write(fd, src + put, n - put);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    put += (size_t)r;
  }
  return true;
}

}  // namespace

DABGPU_EXPORT long long dabgpu_engine_pump(dabgpu_engine *e, int n_streams, const int *in_fds, const int *out_fds,
                                           long long max_callbacks) {
  if (!e || n_streams <= 0 || !in_fds) {
    set_error(DABGPU_ERR_ARG, "engine_pump: null engine / descriptors or no streams");
    return DABGPU_ERR_ARG;
  }
  const int S = n_streams;
  PinBuf in[N_BUF], out;
  for (int b = 0; b < N_BUF; b++)
    if (in[b].reserve((size_t)S * CHUNK)) return DABGPU_ERR_CUDA;
  size_t out_frames = (size_t)S * 4 * 5;  // msc_batch <= 4 frame-producing calls + slack
  if (out.reserve(out_frames * DABGPU_ETI_BYTES)) return DABGPU_ERR_CUDA;
  std::vector<int32_t> ids(out_frames);

  // ---- reader thread: buffer b holds callback k = b mod N_BUF once filled[k] ----
  std::mutex mu;
  std::condition_variable cv;
  long long filled = 0, released = 0;  // callbacks read so far / callbacks whose buffer is free again
  bool eof = false;
  std::thread reader([&] {
    for (long long k = 0; max_callbacks < 0 || k < max_callbacks; k++) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return k - released < N_BUF; });
      }
      uint8_t *dst = in[k % N_BUF].as<uint8_t>();
      bool ok = true;
      for (int s = 0; s < S && ok; s++) ok = read_full(in_fds[s], dst + (size_t)s * CHUNK, CHUNK);
      std::lock_guard<std::mutex> lk(mu);
      if (!ok) break;
      filled = k + 1;
      cv.notify_all();
    }
    std::lock_guard<std::mutex> lk(mu);
    eof = true;
    cv.notify_all();
  });
  auto wait_filled = [&](long long k) {  // true once callback k has been read, false at the end
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return filled > k || eof; });
    return filled > k;
  };

  long long total = 0, submitted = 0, fed = 0;
  int rc = DABGPU_OK;
  auto drain = [&]() -> int {  // the frames of the last feed/flush call -> the output descriptors
    const int n = dabgpu_engine_eti_count(e);
    if (n <= 0) return DABGPU_OK;
    if ((size_t)n > out_frames) {
      out_frames = (size_t)n;
      if (out.reserve(out_frames * DABGPU_ETI_BYTES)) return DABGPU_ERR_CUDA;
      ids.resize(out_frames);
    }
    const int got = dabgpu_engine_fetch_eti(e, out.as<uint8_t>(), ids.data(), n);
    if (got < 0) return got;
    for (int f = 0; f < got; f++) {
      const int s = ids[f];
      if (out_fds && out_fds[s] >= 0 && !write_full(out_fds[s], out.as<uint8_t>() + (size_t)f * DABGPU_ETI_BYTES, DABGPU_ETI_BYTES)) {
        set_error(DABGPU_ERR_STATE, "engine_pump: write to the output of stream %d failed (%s)", s, strerror(errno));
        return DABGPU_ERR_STATE;
      }
    }
    total += got;
    return DABGPU_OK;
  };
  const int AHEAD = 2;
  while (rc == DABGPU_OK) {
    // keep the uploads AHEAD callbacks in front of the one being processed
    while (submitted < fed + 1 + AHEAD && wait_filled(submitted)) {
      if ((rc = dabgpu_engine_submit_iq(e, in[submitted % N_BUF].as<uint8_t>(), CHUNK, CHUNK))) break;
      submitted++;
      // a pinned buffer is free again once its host->device copy has completed (at most 3 are on
      // their way, so the reader always finds room among the N_BUF buffers)
      std::lock_guard<std::mutex> lk(mu);
      released = submitted - dabgpu_engine_uploads_in_flight(e);
      cv.notify_all();
    }
    if (rc || fed >= submitted) break;
    if ((rc = dabgpu_engine_feed_submitted(e))) break;
    fed++;
    rc = drain();
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    released = 1LL << 60;  // let the reader run to its end
    cv.notify_all();
  }
  reader.join();
  if (rc == DABGPU_OK && !(rc = dabgpu_engine_flush(e))) rc = drain();
  for (int b = 0; b < N_BUF; b++) in[b].release();
  out.release();
  return rc ? rc : total;
}
