// Host-side helpers shared by the .cu files of libb200.so.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200 {

struct DeviceInfo {
  int valid;
  int sm_count;
  int cc_major, cc_minor;
  int l2_bytes;
  int smem_optin;
};
const DeviceInfo& device_info();

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension ld_elems,
// box = box_rows x box_cols elements, 128-byte swizzle (box_cols * 2 bytes must be 128).
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);

}  // namespace b200
