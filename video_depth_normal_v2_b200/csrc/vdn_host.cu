// Library state + tensor-map encoding.  cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint), so the library links against libcudart only.
#include <atomic>
#include <mutex>

#include "../../include/vdn_b200.h"
#include "vdn_host.h"

namespace vdn {

static thread_local std::string g_error;
static std::atomic<int> g_fmt{0};  // 0 = fp16 (default), 1 = bf16
static std::atomic<long long> g_launches{0};

int set_error(const std::string& msg) {
  g_error = msg;
  return 1;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}
int num_sms() {
  static std::atomic<int> n[kMaxDevices];  // zero-initialised; one process may drive several GPUs
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int get_operand_format() { return g_fmt.load(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int fmt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return make_tensor_map_ex(out, base, fmt, 128, rank, dims, strides_bytes, box);
}

int make_tensor_map_ex(CUtensorMap* out, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error("cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error("tensor map: box dim out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) return set_error("tensor map: global stride not a multiple of 16 bytes");
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return set_error("tensor map: base not 16-byte aligned");
  const CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

}  // namespace vdn

extern "C" {
const char* vdn_last_error(void) { return vdn::g_error.c_str(); }
int vdn_version(void) { return 100; }
int vdn_set_operand_format(int fmt) {
  if (fmt != 0 && fmt != 1) return vdn::set_error("operand format must be 0 (fp16) or 1 (bf16)");
  vdn::g_fmt.store(fmt);
  return 0;
}
int vdn_get_operand_format(void) { return vdn::g_fmt.load(); }
int64_t vdn_launch_count(void) { return vdn::g_launches.load(); }
void vdn_reset_launch_count(void) { vdn::g_launches.store(0); }
void vdn_add_launch_count(int64_t n) { vdn::g_launches.fetch_add(n, std::memory_order_relaxed); }
}
