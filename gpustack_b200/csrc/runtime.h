// Host-side helpers shared by the .cu files of libb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
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

// Same over a byte matrix (fp8 operands): box_cols bytes must be 128.
int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                    uint32_t box_rows, uint32_t box_cols);

// 3-D, un-swizzled map used by the GEMM epilogues' TMA STORES: tensor [planes][rows][cols] of
// elem_bytes-wide elements (2 = bf16, 4 = fp32), row pitch ld_elems, plane pitch plane_elems
// (split-K partial planes; 1 plane otherwise); box = box_cols x box_rows x 1.  The hardware clips
// the part of a box that lies outside [cols, rows, planes], so tail tiles need no predicates.
int make_tmap_out_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                     uint64_t planes, uint64_t ld_elems, uint64_t plane_elems, uint32_t box_cols,
                     uint32_t box_rows);

// B200_PDL=0 disables programmatic dependent launch (default on).
bool pdl_enabled();

// Kernel classes whose launches can be taken out of programmatic dependent launch individually
// (B200_PDL_OFF_CLASSES = bitmask, default 2): 1 = step-final sampler, 2 = MoE routing / grouped GEMMs.
enum PdlClass { kPdlDefault = 0, kPdlSampler = 1, kPdlMoe = 2 };
bool pdl_class_enabled(int cls);

// Launch with the programmatic-stream-serialization attribute (and an optional cluster size).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cls(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                  cudaStream_t stream, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled() && pdl_class_enabled(cls)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, int cluster, Args&&... args) {
  return launch_pdl_cls(kPdlDefault, kernel, grid, block, smem, stream, cluster, static_cast<Args&&>(args)...);
}

}  // namespace b200
