// Error plumbing + version / device probes of the C ABI (include/magat_gat.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace magat {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting + optional per-kernel CUDA-event timing ----------------------------
// Every kernel launch of the library goes through check_launch(), which counts it and, when
// profiling is on, records one event behind it; entry points record a begin event.  A kernel's
// time is the gap to the previous event on the same stream (launches are serialised there).
static std::atomic<long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfRec { const char* name; cudaEvent_t ev; };
static std::vector<ProfRec> g_prof;
static constexpr size_t kProfCap = 1 << 16;

static void prof_record(const char* name, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof.size() >= kProfCap) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_prof.push_back({name, ev});
}

void prof_begin(cudaStream_t st) {
  if (g_prof_on.load(std::memory_order_relaxed)) prof_record(nullptr, st);
}

// ---- per-device caches (one process may drive several devices) -----------------------------------
static std::atomic<int> g_sm_count[64];
static std::atomic<unsigned long long> g_attr_done[KID_MAX];     // [kernel id] -> bit per device

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

int device_sm_count() {
  const int dev = current_device();
  int n = g_sm_count[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 0;
  }
  g_sm_count[dev].store(n, std::memory_order_relaxed);
  return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: do it once per (kernel, device)
int ensure_dyn_smem(int kernel_id, const void* func, size_t bytes, const char* name) {
  if (kernel_id < 0 || kernel_id >= KID_MAX) {
    set_error("ensure_dyn_smem(%s): kernel id %d outside [0, %d)", name, kernel_id, (int)KID_MAX);
    return MAGAT_E_CUDA;
  }
  const int dev = current_device();
  const unsigned long long bit = 1ull << dev;
  if (g_attr_done[kernel_id].load(std::memory_order_acquire) & bit) return MAGAT_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%s, %zu B): %s", name, bytes, cudaGetErrorString(e));
    return MAGAT_E_CUDA;
  }
  g_attr_done[kernel_id].fetch_or(bit, std::memory_order_release);
  return MAGAT_OK;
}

int check_launch(const char* what, cudaStream_t st) {
  static const bool sync_check = getenv("MAGAT_SYNC_CHECK") != nullptr;     // debugging: pin an asynchronous fault to its kernel
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && sync_check) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MAGAT_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) prof_record(what, st);
  return MAGAT_OK;
}

}  // namespace magat

extern "C" int magat_abi_version(void) { return MAGAT_ABI_VERSION; }

extern "C" const char* magat_last_error(void) { return magat::g_err; }

extern "C" int magat_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    magat::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return MAGAT_E_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    magat::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return MAGAT_E_DEVICE;
  }
  return MAGAT_OK;
}

extern "C" long magat_launch_count(void) { return magat::g_launches.load(); }

extern "C" void magat_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(magat::g_prof_mu);
  for (auto& r : magat::g_prof) cudaEventDestroy(r.ev);
  magat::g_prof.clear();
  magat::g_prof_on.store(on ? 1 : 0);
}

// Synchronises, then writes "name,launches,total_ms\n" lines (aggregated by kernel name, in first-use
// order) into buf; returns the number of bytes needed (excluding the terminator).  Clears the records.
extern "C" long magat_profile_collect(char* buf, long cap) {
  std::lock_guard<std::mutex> lk(magat::g_prof_mu);
  std::vector<std::string> order;
  std::map<std::string, std::pair<long, double>> agg;
  cudaEvent_t prev = nullptr;
  for (auto& r : magat::g_prof) {
    cudaEventSynchronize(r.ev);
    if (r.name && prev) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, prev, r.ev) == cudaSuccess) {
        auto it = agg.find(r.name);
        if (it == agg.end()) {
          order.push_back(r.name);
          agg[r.name] = {1, ms};
        } else {
          it->second.first += 1;
          it->second.second += ms;
        }
      }
    }
    prev = r.ev;
  }
  std::string out;
  char line[256];
  for (auto& n : order) {
    snprintf(line, sizeof(line), "%s,%ld,%.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  for (auto& r : magat::g_prof) cudaEventDestroy(r.ev);
  magat::g_prof.clear();
  if (buf && cap > 0) {
    const long n = (long)out.size() < cap - 1 ? (long)out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (long)out.size();
}
