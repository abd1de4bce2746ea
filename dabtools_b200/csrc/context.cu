// context.cu -- error state, stream selection, scratch buffers, one-time device setup.
#include <cstdarg>
#include <atomic>
#include <mutex>

#include "common.cuh"

#include <algorithm>

namespace dabgpu {

int viterbi_init_constants();
int msc_init_constants();
int ofdm_init_constants();

static thread_local int t_err_code = 0;
static thread_local char t_err_msg[512] = "";
static thread_local cudaStream_t t_stream = nullptr;

void set_error(int code, const char *fmt, ...) {
  t_err_code = code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err_msg, sizeof t_err_msg, fmt, ap);
  va_end(ap);
  if (getenv("DABGPU_VERBOSE")) fprintf(stderr, "libdabgpu: %s\n", t_err_msg);
}
int last_error_code() { return t_err_code; }
cudaStream_t current_stream() { return t_stream; }

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return DABGPU_OK;
  if (p) CUDA_TRY(cudaFree(p));
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CUDA_TRY(cudaMalloc(&p, want));
  cap = want;
  return DABGPU_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}
int PinBuf::reserve(size_t bytes) {
  if (bytes <= cap) return DABGPU_OK;
  if (p) CUDA_TRY(cudaFreeHost(p));
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CUDA_TRY(cudaMallocHost(&p, want));
  cap = want;
  return DABGPU_OK;
}
void PinBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

__global__ void ctl_copy_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t words) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) dst[i] = src[i];
}
int launch_ctl_copy(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  if (!bytes) return DABGPU_OK;
  const size_t words = bytes / 4;
  const unsigned blocks = (unsigned)std::min<size_t>((words + 255) / 256, 592);
  ctl_copy_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t *>(dst), static_cast<const uint32_t *>(src), words);
  LAUNCH_CHECK();
  return DABGPU_OK;
}

static std::atomic<uint64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

static std::mutex g_init_mu;
static bool g_ready[64];

int ensure_device_ready() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0) {
    set_error(DABGPU_ERR_NO_DEVICE, "no CUDA device available (%s); libdabgpu has no CPU fallback",
              cudaGetErrorString(e));
    cudaGetLastError();
    return DABGPU_ERR_NO_DEVICE;
  }
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (dev < 64 && g_ready[dev]) return DABGPU_OK;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error(DABGPU_ERR_NO_DEVICE, "device %d is sm_%d%d; libdabgpu is built for sm_100a only", dev,
              prop.major, prop.minor);
    return DABGPU_ERR_NO_DEVICE;
  }
  int rc;
  if ((rc = viterbi_init_constants())) return rc;
  if ((rc = msc_init_constants())) return rc;
  if ((rc = ofdm_init_constants())) return rc;
  if (dev < 64) g_ready[dev] = true;
  return DABGPU_OK;
}

}  // namespace dabgpu

using namespace dabgpu;

DABGPU_EXPORT int dabgpu_last_error(void) { return t_err_code; }
DABGPU_EXPORT const char *dabgpu_last_error_string(void) { return t_err_msg; }
DABGPU_EXPORT void dabgpu_clear_error(void) {
  t_err_code = 0;
  t_err_msg[0] = 0;
}
DABGPU_EXPORT int dabgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
DABGPU_EXPORT int dabgpu_set_device(int dev) {
  CUDA_TRY(cudaSetDevice(dev));
  return ensure_device_ready();
}
DABGPU_EXPORT void dabgpu_set_stream(void *cuda_stream) { t_stream = (cudaStream_t)cuda_stream; }
DABGPU_EXPORT uint64_t dabgpu_launch_count(void) { return launch_count(); }
DABGPU_EXPORT int dabgpu_synchronize(void) {
  CUDA_TRY(cudaStreamSynchronize(t_stream));
  return DABGPU_OK;
}
