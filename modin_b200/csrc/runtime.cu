// runtime.cu — device discovery, memory, transfers and the error channel of the C ABI.
#include <mutex>

#include "common.cuh"

namespace mb200 {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

static DevProps g_props[64];
static bool g_props_ok[64] = {false};
static std::mutex g_props_mu;

int dev_props(DevProps* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail("cudaGetDevice (no CUDA device: there is no CPU fallback)", e);
  if (dev < 0 || dev >= 64) return fail("dev_props", "device ordinal out of range");
  std::lock_guard<std::mutex> lk(g_props_mu);
  if (!g_props_ok[dev]) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cuda_fail("cudaGetDeviceProperties", e);
    if (p.major != 10) {
      char msg[128];
      snprintf(msg, sizeof(msg), "device %d is sm_%d%d; libmodin_b200 carries sm_100a code only", dev,
               p.major, p.minor);
      return fail("dev_props", msg);
    }
    g_props[dev].sm_count = p.multiProcessorCount;
    g_props[dev].l2_bytes = (size_t)p.l2CacheSize;
    g_props[dev].cc_major = p.major;
    g_props[dev].cc_minor = p.minor;
    g_props[dev].smem_optin = p.sharedMemPerBlockOptin;
    // Stream-ordered allocations (group / join tables) come from the device's default pool; keep freed
    // memory in the pool instead of returning it to the OS at every synchronisation (default threshold 0),
    // otherwise each groupby step pays for re-mapping ~100 MB of table memory.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_props_ok[dev] = true;
  }
  *out = g_props[dev];
  return 0;
}
}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_abi_version(void) { return MB200_ABI_VERSION; }
const char* mb200_last_error(void) { return g_err; }
int64_t mb200_launch_count(void) { return (int64_t)g_launches.load(); }

int mb200_set_device(int device) {
  MB_CUDA(cudaSetDevice(device));
  return 0;
}

int mb200_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount (no CUDA device: there is no CPU fallback)", e);
  if (device < 0 || device >= n) return fail("mb200_device_check", "no such device");
  cudaDeviceProp p;
  MB_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return fail("mb200_device_check", "device is not sm_100 (Blackwell B200)");
  return 0;
}

int mb200_device_info(int device, int* sm_count, size_t* l2_bytes, size_t* total_mem, int* cc_major,
                      int* cc_minor) {
  cudaDeviceProp p;
  MB_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

int mb200_alloc(void** ptr, size_t bytes, mb200_stream_t stream) {
  if (!ptr) return fail("mb200_alloc", "null out pointer");
  if (bytes == 0) {
    *ptr = nullptr;
    return 0;
  }
  MB_CUDA(cudaMallocAsync(ptr, bytes, (cudaStream_t)stream));
  return 0;
}
int mb200_free(void* ptr, mb200_stream_t stream) {
  if (!ptr) return 0;
  MB_CUDA(cudaFreeAsync(ptr, (cudaStream_t)stream));
  return 0;
}
int mb200_alloc_host(void** ptr, size_t bytes) {
  if (!ptr) return fail("mb200_alloc_host", "null out pointer");
  MB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
  return 0;
}
int mb200_free_host(void* ptr) {
  if (!ptr) return 0;
  MB_CUDA(cudaFreeHost(ptr));
  return 0;
}
int mb200_h2d(void* dst, const void* src, size_t bytes, mb200_stream_t stream) {
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
int mb200_d2h(void* dst, const void* src, size_t bytes, mb200_stream_t stream) {
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
int mb200_d2d(void* dst, const void* src, size_t bytes, mb200_stream_t stream) {
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int mb200_memset(void* dst, int byte, size_t bytes, mb200_stream_t stream) {
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t)stream));
  return 0;
}
int mb200_stream_sync(mb200_stream_t stream) {
  MB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- persisting-L2 carve-out
// The carve-out (cudaLimitPersistingL2CacheSize) is taken out of the normally managed L2 for as long as the
// limit is set, whether or not any line is pinned: measured, a hash-table groupby that ran after a dense one
// left the limit at its 82.9 MB maximum dropped from 30.8 to 11.4 G rows/s.  So tables hold a reference
// while they need it and the last release gives the whole L2 back.
namespace mb200 {
static std::mutex g_carve_mu;
static int g_carve_refs = 0;
static bool g_carve_set = false;  // the device limit is currently raised

// cudaDeviceSetLimit / cudaCtxResetPersistingL2Cache are context-wide calls that wait for the device, so the carve-out
// is NOT handed back the moment the last table dies: a stream of groupby queries would pay two device synchronisations
// per query (measured, round 2: the whole host side of the next query serialised behind the kernel -- 20.7 ms per step
// instead of 16.8).  It is dropped lazily instead, by the next kernel family that wants the whole L2 for its own
// working set (l2_carveout_drop_idle: the join build / probes), or when the process ends.
size_t l2_carveout_acquire(size_t* max_window_bytes) {
  int dev = 0, max_persist = 0, max_window = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  if (max_persist <= 0 || max_window <= 0) return 0;
  std::lock_guard<std::mutex> lk(g_carve_mu);
  if (!g_carve_set) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    g_carve_set = true;
  }
  ++g_carve_refs;
  if (max_window_bytes) *max_window_bytes = (size_t)max_window;
  return (size_t)max_persist;
}

void l2_carveout_release() {
  std::lock_guard<std::mutex> lk(g_carve_mu);
  if (g_carve_refs > 0) --g_carve_refs;
}

void l2_carveout_drop_idle() {
  std::lock_guard<std::mutex> lk(g_carve_mu);
  if (g_carve_set && g_carve_refs == 0) {
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    g_carve_set = false;
  }
}
}  // namespace mb200
