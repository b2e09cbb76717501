// Host-side helpers: error reporting across the C ABI, the TMA descriptor encoder, device queries.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace gsd {

inline std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define GSD_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::gsd::fail(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define GSD_CHECK(cond, ...)                          \
  do {                                                \
    if (!(cond)) return ::gsd::fail(-1, __VA_ARGS__); \
  } while (0)

#define GSD_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// gsd_debug_plan_* (CPU tests of the launch rules): plan a launch without a GPU or driver -- tensor maps are skipped
inline bool& plan_only_mode() {
  static thread_local bool on = false;
  return on;
}

// bf16 tensor map, `rank` dims (innermost first), byte strides for dims 1..rank-1, all-ones element strides,
// zero fill out of bounds.
inline int encode_bf16_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, CUtensorMapSwizzle swz, bool weights) {
  if (plan_only_mode()) return 0;
  PFN_encodeTiled fn = get_encode_fn();
  GSD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gd[5] = {1, 1, 1, 1, 1};
  cuuint64_t gs[4] = {0, 0, 0, 0};
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  GSD_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  weights ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GSD_CHECK(r == CUDA_SUCCESS,
            "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r,
            rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
            (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0],
            rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
  return 0;
}

// Every ABI entry that launches work makes `device` current for the duration of the call and restores the caller's
// current device on return (a torch process may drive cuda:1 while its current device is 0).
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) {
    if (device < 0) return;          // host-only call (dry-run plans of the CPU tests)
    err_ = cudaGetDevice(&prev_);
    if (err_ == cudaSuccess && prev_ != device) {
      err_ = cudaSetDevice(device);
      switched_ = err_ == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(prev_);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
  int status() const { return err_ == cudaSuccess ? 0 : fail(-2, "cudaSetDevice failed: %s", cudaGetErrorString(err_)); }

 private:
  int prev_ = 0;
  bool switched_ = false;
  cudaError_t err_ = cudaSuccess;
};

#define GSD_DEVICE(dev)                  \
  ::gsd::DeviceGuard _gsd_guard(dev);    \
  GSD_TRY(_gsd_guard.status())

// device that owns a pointer (ops that take no device argument launch where their first tensor lives)
inline int device_of(const void* ptr, int* device) {
  cudaPointerAttributes a;
  GSD_CUDA(cudaPointerGetAttributes(&a, ptr));
  GSD_CHECK(a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged, "expected a device pointer");
  *device = a.device;
  return 0;
}
// launch on the device that owns `ptr`, restoring the caller's current device afterwards
#define GSD_DEVICE_OF(ptr)               \
  int _gsd_dev = 0;                      \
  GSD_TRY(::gsd::device_of(ptr, &_gsd_dev)); \
  GSD_DEVICE(_gsd_dev)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gsd
