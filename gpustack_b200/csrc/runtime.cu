// Runtime glue of libb200.so: error reporting, device queries, TMA descriptor creation and
// CUDA IPC handles.  No kernels here.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

static thread_local char g_err[512] = "";

void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b200_check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    b200_set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

namespace b200 {

static std::once_flag g_dev_once;
static DeviceInfo g_dev;

const DeviceInfo& device_info() {
  std::call_once(g_dev_once, [] {
    int dev = 0;
    cudaDeviceProp p;
    memset(&g_dev, 0, sizeof(g_dev));
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      g_dev.sm_count = p.multiProcessorCount;
      g_dev.cc_major = p.major;
      g_dev.cc_minor = p.minor;
      g_dev.l2_bytes = p.l2CacheSize;
      g_dev.smem_optin = (int)p.sharedMemPerBlockOptin;
      g_dev.valid = 1;
    }
  });
  return g_dev;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

bool pdl_class_enabled(int cls) {
  static int off = -1;
  if (off < 0) {
    // Every class is a programmatic dependent by default.  (Round 1 had to serialise the MoE class:
    // with the entry-time trigger, CUDA-graph replays of the tiny Mixtral decode step computed wrong
    // logits -- DESIGN.md section 9.  The late trigger in common.cuh removes the cause; the knob
    // B200_PDL_OFF_CLASSES=<bitmask> stays for experiments.)
    const char* e = getenv("B200_PDL_OFF_CLASSES");
    off = e ? atoi(e) : 0;
  }
  return cls == 0 || (off & cls) == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                        uint32_t box_rows, uint32_t box_cols, int elem_bytes);

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, base, rows, cols, ld_elems, box_rows, box_cols, 2);
}

int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                    uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, base, rows, cols, ld_elems, box_rows, box_cols, 1);
}

static int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                        uint32_t box_rows, uint32_t box_cols, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    b200_set_error("cuTensorMapEncodeTiled entry point not available");
    return B200_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld_elems * elem_bytes) & 15)) {
    b200_set_error("TMA operand must be 16-byte aligned (base %p, ld %llu elems)", base,
                   (unsigned long long)ld_elems);
    return B200_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                  const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %llu cols %llu ld %llu box %ux%u)",
                   (int)r, (unsigned long long)rows, (unsigned long long)cols,
                   (unsigned long long)ld_elems, box_rows, box_cols);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int make_tmap_out_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                     uint64_t planes, uint64_t ld_elems, uint64_t plane_elems, uint32_t box_cols,
                     uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    b200_set_error("cuTensorMapEncodeTiled entry point not available");
    return B200_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld_elems * elem_bytes) & 15) ||
      ((box_cols * elem_bytes) & 15)) {
    b200_set_error("TMA store target must be 16-byte aligned (base %p, row pitch %llu B, box %u B)", base,
                   (unsigned long long)(ld_elems * elem_bytes), box_cols * elem_bytes);
    return B200_ERR_INVALID;
  }
  if (planes < 1) planes = 1;
  if (plane_elems == 0) plane_elems = rows * ld_elems;
  cuuint64_t dims[3] = {cols, rows, planes};
  cuuint64_t strides[2] = {ld_elems * elem_bytes, plane_elems * elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled(out) failed: CUresult %d (cols %llu rows %llu planes %llu ld %llu box %ux%u)",
                   (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)planes,
                   (unsigned long long)ld_elems, box_cols, box_rows);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // namespace b200

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

const char* b200_last_error(void) { return g_err; }

int b200_device_info(int* sm_count, int* cc_major, int* cc_minor, int* l2_bytes) {
  const b200::DeviceInfo& d = b200::device_info();
  if (!d.valid) {
    b200_set_error("no CUDA device");
    return B200_ERR_CUDA;
  }
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  if (l2_bytes) *l2_bytes = d.l2_bytes;
  return B200_OK;
}

int b200_dev_malloc(size_t bytes, void** dev_ptr_out) {
  cudaError_t e = cudaMalloc(dev_ptr_out, bytes);
  if (e == cudaSuccess) e = cudaMemset(*dev_ptr_out, 0, bytes);
  if (e != cudaSuccess) {
    b200_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int b200_dev_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) {
    b200_set_error("cudaFree: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int b200_ipc_get_handle(void* dev_ptr, void* handle_out_host) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e != cudaSuccess) {
    b200_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle_out_host, &h, sizeof(h));
  return B200_OK;
}

int b200_ipc_open_handle(const void* handle_host, void** dev_ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    b200_set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

int b200_ipc_close_handle(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) {
    b200_set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // extern "C"
