// Host-side helpers shared by the launchers (error state, SM count, tensor-map encoding, launch counter).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace vdn {
int set_error(const std::string& msg);  // records msg, returns 1
int check_launch(const char* what);     // cudaGetLastError() -> 0 / set_error
constexpr int kMaxDevices = 64;
int current_device();                    // cudaGetDevice(), clamped to [0, kMaxDevices)
int num_sms();                           // of the current device
void count_launch();
int get_operand_format();
// Encode a tiled, 128B-swizzled tensor map for a 16-bit tensor. dims/box innermost first; strides (bytes) for dims 1..rank-1.
int make_tensor_map(CUtensorMap* out, const void* base, int fmt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
// General form: dtype 0 = fp16, 1 = bf16, 2 = fp32; swizzle_bytes 0 / 32 / 64 / 128.
int make_tensor_map_ex(CUtensorMap* out, const void* base, int dtype, int swizzle_bytes, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box);
}  // namespace vdn
