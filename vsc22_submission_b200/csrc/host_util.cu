#include <stdlib.h>

#include "host_util.h"

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace vscb200 {

static thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_last_error(const std::string& msg) { g_last_error = msg; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// Encoded descriptors are cached per thread, keyed by every argument: an encoder plan launches the same few dozen
// (buffer, shape, box) combinations thousands of times per step, and cuTensorMapEncodeTiled costs ~1 us a call.
namespace {
struct TmapKey {
  const void* base; uint64_t rows, cols, ld; uint32_t box_rows, box_cols; int dtype, elem_bytes, swz;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols &&
           dtype == o.dtype && elem_bytes == o.elem_bytes && swz == o.swz;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.rows, k.cols, k.ld, (static_cast<uint64_t>(k.box_rows) << 32) | k.box_cols,
                       (static_cast<uint64_t>(k.dtype) << 16) ^ (static_cast<uint64_t>(k.elem_bytes) << 8) ^ static_cast<uint64_t>(k.swz)})
      h = (h ^ v) * 0xBF58476D1CE4E5B9ull + (h >> 29);
    return static_cast<size_t>(h);
  }
};
thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> t_tmaps;
}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
  const TmapKey key{base, rows, cols, ld, box_rows, box_cols, static_cast<int>(dtype), elem_bytes, swizzle128 ? 1 : 0};
  auto hit = t_tmaps.find(key);
  if (hit != t_tmaps.end()) { *out = hit->second; return VSCB200_OK; }
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled driver entry point not available");
    return VSCB200_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                   " (rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld) +
                   " box=" + std::to_string(box_rows) + "x" + std::to_string(box_cols) + ")");
    return VSCB200_ERR_CUDA;
  }
  if (t_tmaps.size() >= 8192) t_tmaps.clear();
  t_tmaps.emplace(key, *out);
  return VSCB200_OK;
}

bool stream_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VSCB200_STREAM_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VSCB200_PDL"); return e ? atoi(e) != 0 : false; }();
  return on;
}

int device_sm_count() {
  static std::atomic<int> sms[64];                 // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// ---------------------------------------------------------------- caching allocator
struct PoolBlock { void* p; size_t bytes; int device; cudaEvent_t ev; };
static std::mutex g_pool_mu;
static std::vector<PoolBlock> g_pool_free;
static std::vector<PoolBlock> g_pool_live;

static size_t pool_round(size_t b) {
  const size_t g = b < (1u << 20) ? 512 : (2u << 20);
  return (b + g - 1) / g * g;
}

int pool_alloc(void** p, size_t bytes, cudaStream_t stream) {
  bytes = pool_round(bytes ? bytes : 1);
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < static_cast<int>(g_pool_free.size()); ++i) {
      const PoolBlock& b = g_pool_free[i];
      if (b.device == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + (1u << 20) &&
          (best < 0 || b.bytes < g_pool_free[best].bytes))
        best = i;
    }
    if (best >= 0) {
      PoolBlock b = g_pool_free[best];
      g_pool_free.erase(g_pool_free.begin() + best);
      if (b.ev) cudaStreamWaitEvent(stream, b.ev, 0);
      g_pool_live.push_back(b);
      *p = b.p;
      return VSCB200_OK;
    }
  }
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) {
    pool_trim();                      // give cached blocks back and retry once
    e = cudaMalloc(&q, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("cudaMalloc(" + std::to_string(bytes) + ") failed: " + cudaGetErrorString(e));
    return VSCB200_ERR_NOMEM;
  }
  PoolBlock b{q, bytes, dev, nullptr};
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pool_live.push_back(b);
  *p = q;
  return VSCB200_OK;
}

void pool_free(void* p, cudaStream_t stream) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (size_t i = 0; i < g_pool_live.size(); ++i) {
    if (g_pool_live[i].p == p) {
      PoolBlock b = g_pool_live[i];
      g_pool_live.erase(g_pool_live.begin() + i);
      // The guarding event belongs to the block's device, which need not be the caller's current one (an index
      // destroyed from a Python finaliser while another GPU is current): switch for the create / record, and if the
      // record fails (`stream` of a different device) order the reuse by draining the stream instead.
      DeviceGuard guard(b.device);
      if (!b.ev && cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) b.ev = nullptr;
      if (!b.ev || cudaEventRecord(b.ev, stream) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(stream);
        if (b.ev) { cudaEventDestroy(b.ev); b.ev = nullptr; }
      }
      g_pool_free.push_back(b);
      return;
    }
  }
  cudaFree(p);   // not ours
}

void pool_trim() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (PoolBlock& b : g_pool_free) {
    if (b.ev) { cudaEventSynchronize(b.ev); cudaEventDestroy(b.ev); }
    cudaFree(b.p);
  }
  g_pool_free.clear();
  cudaGetLastError();
}

// ---------------------------------------------------------------- event profiler
struct ProfRec { cudaEvent_t a = nullptr, b = nullptr; int kind = 0; double work = 0; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;       // records of the current collection window
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int kind, cudaStream_t stream, double work) : slot_(-1), stream_(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.a = get_event();
  r.b = get_event();
  r.kind = kind;
  r.work = work;
  cudaEventRecord(r.a, stream);
  g_prof.push_back(r);
  slot_ = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot_ < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[slot_].b, stream_);
}

}  // namespace vscb200

extern "C" {
int vscb200_prof_enable(int on) {
  vscb200::g_prof_on.store(on ? 1 : 0);
  return VSCB200_OK;
}
// Sum the device time (ms), launch count and algorithmic work of the records of `kind` collected
// since the last call, then drop ALL records.  Synchronises on the recorded events.
int vscb200_prof_collect(double* ms_by_kind, int64_t* launches_by_kind, double* work_by_kind, int nkinds) {
  using namespace vscb200;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < nkinds; ++k) { ms_by_kind[k] = 0; launches_by_kind[k] = 0; work_by_kind[k] = 0; }
  for (ProfRec& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess &&
        r.kind < nkinds) {
      ms_by_kind[r.kind] += ms;
      launches_by_kind[r.kind] += 1;
      work_by_kind[r.kind] += r.work;
    }
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  cudaGetLastError();
  return VSCB200_OK;
}
const char* vscb200_last_error(void) { return vscb200::g_last_error.c_str(); }
int vscb200_version(void) { return 100; }
int64_t vscb200_launch_count(void) { return vscb200::g_launch_count.load(); }
void vscb200_free(void* p) { free(p); }
int vscb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int vscb200_trim(void) {
  vscb200::pool_trim();
  return VSCB200_OK;
}
int vscb200_set_device(int device) {
  VSCB_CUDA_OK(cudaSetDevice(device));
  return VSCB200_OK;
}
}
