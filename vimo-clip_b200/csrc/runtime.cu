// Library runtime: error strings, device queries, TMA tensor-map encoding, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "vimoclip_b200.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void vmc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void vmc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int vmc_num_sms() {
  static std::mutex mu;
  static int cached[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
    have[dev] = true;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static std::mutex mu;
  static PFN_encodeTiled fn = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int vmc_encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  VMC_CHECK_ARG(fn != nullptr, VMC_ERR_DRIVER,
                "cuTensorMapEncodeTiled not available from the CUDA driver");
  VMC_CHECK_ARG(rank >= 2 && rank <= 3, VMC_ERR_ARG, "tensor map rank %d unsupported", rank);
  VMC_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, VMC_ERR_ALIGN,
                "TMA base pointer must be 16-byte aligned");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    VMC_CHECK_ARG((gstr[i] & 15) == 0, VMC_ERR_ALIGN, "TMA stride %llu not a multiple of 16 bytes",
                  (unsigned long long)gstr[i]);
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VMC_CHECK_ARG(r == CUDA_SUCCESS, VMC_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d",
                (int)r);
  return VMC_OK;
}

extern "C" {

const char* vmc_last_error(void) { return g_err; }

int vmc_abi_version(void) { return VMC_ABI_VERSION; }

long long vmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void vmc_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

int vmc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VMC_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VMC_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return VMC_OK;
}

}  // extern "C"
