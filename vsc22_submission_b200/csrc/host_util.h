// Host-side helpers shared by the C-ABI translation units: error codes, CUDA checks,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/vscb200.h"

namespace vscb200 {

void set_last_error(const std::string& msg);
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define VSCB_CUDA_OK(expr)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::vscb200::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                std::to_string(__LINE__) + ")");                                  \
      return VSCB200_ERR_CUDA;                                                                    \
    }                                                                                             \
  } while (0)

#define VSCB_REQUIRE(cond, msg)                                                                   \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::vscb200::set_last_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return VSCB200_ERR_INVALID;                                                                 \
    }                                                                                             \
  } while (0)

// 2-D row-major tensor [rows, cols] of `elem_bytes` elements with `ld` elements between rows;
// box = [box_rows, box_cols]; 128B swizzle when box_cols*elem_bytes == 128, else none.
int make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dtype, int elem_bytes, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols, bool swizzle128);

int device_sm_count();     // of the current device (cached per device)

// Switch to `dev` for the lifetime of the guard (several GPUs driven by one process: faiss_compat.index_cpu_to_all_gpus)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
bool pdl_enabled();     // VSCB200_PDL=1: launch with programmatic stream serialization (measured: no gain, off by default)

// <<<>>> with the programmatic-stream-serialization attribute: the kernel may start while its predecessor in the
// stream drains; it must execute griddepcontrol.wait (ptx.cuh pdl_wait) before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// the encoder's launches: measured no gain, behind VSCB200_PDL=1
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_pdl_if(pdl_enabled(), kernel, grid, block, smem, stream, args...);
}
// the streaming search's chain of small launches: on unless VSCB200_STREAM_PDL=0
bool stream_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_pdl_if(stream_pdl_enabled(), kernel, grid, block, smem, stream, args...);
}

// Caching device allocator for the index's transient buffers (score workspace, staging, operand
// planes).  The reference builds a fresh faiss index per score_normalize call
// (score_normalization.py:87); without caching every such call would pay cudaMalloc/cudaFree
// (milliseconds + device syncs).  Freed blocks keep an event recorded on the freeing stream; a block is
// handed out again only after the new stream waits on it.  Memory returns to the driver at vscb200_trim().
int pool_alloc(void** p, size_t bytes, cudaStream_t stream);
void pool_free(void* p, cudaStream_t stream);
void pool_trim();

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's live roofline
// numbers).  Disabled by default: zero cost beyond one relaxed atomic load per launch.
enum ProfKind { kProfGemm = 0, kProfAttention = 1, kProfLayerNorm = 2, kProfVitOther = 3, kProfScores = 4,
                kProfSelect = 5, kProfKinds = 6 };
struct ProfScope {
  ProfScope(int kind, cudaStream_t stream, double work);
  ~ProfScope();
  int slot_;
  cudaStream_t stream_;
};

}  // namespace vscb200
