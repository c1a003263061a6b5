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

// ---- per-kernel-class event profile -------------------------------------------------------
#include <vector>
namespace {
struct ProfRec {
  int cat;
  cudaEvent_t e0, e1;
  double flops, bytes;
};
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
std::vector<ProfRec*> g_prof;
}  // namespace

VmcProfScope::VmcProfScope(int c, cudaStream_t s, double flops, double bytes) : cat(c), stream(s), rec(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec* r = new ProfRec;
  r->cat = c;
  r->flops = flops;
  r->bytes = bytes;
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, s);
  rec = r;
}
VmcProfScope::~VmcProfScope() {
  if (!rec) return;
  ProfRec* r = static_cast<ProfRec*>(rec);
  cudaEventRecord(r->e1, stream);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back(r);
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
  return vmc_encode_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, 128);
}

int vmc_encode_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
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
                  gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VMC_CHECK_ARG(r == CUDA_SUCCESS, VMC_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d",
                (int)r);
  return VMC_OK;
}

static std::atomic<long long> g_options[8];

int vmc_get_option(int option) {
  return (option >= 0 && option < 8) ? (int)g_options[option].load(std::memory_order_relaxed) : 0;
}
long long vmc_get_option64(int option) {
  return (option >= 0 && option < 8) ? g_options[option].load(std::memory_order_relaxed) : 0;
}

extern "C" {

int vmc_set_option(int option, long long value) {
  VMC_CHECK_ARG(option >= 0 && option < 8, VMC_ERR_ARG, "vmc_set_option: unknown option %d", option);
  g_options[option].store(value, std::memory_order_relaxed);
  return VMC_OK;
}

const char* vmc_last_error(void) { return g_err; }

int vmc_abi_version(void) { return VMC_ABI_VERSION; }

long long vmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void vmc_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

void vmc_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (ProfRec* r : g_prof) {
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  g_prof.clear();
  g_prof_on.store(true);
}

int vmc_profile_end(double* ms, double* flops, double* bytes, long long* launches, int ncat) {
  g_prof_on.store(false);
  VMC_CHECK_ARG(ms && flops && bytes && launches && ncat >= VMC_K_COUNT, VMC_ERR_ARG,
                "vmc_profile_end: need %d-entry arrays", (int)VMC_K_COUNT);
  VMC_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < ncat; ++i) {
    ms[i] = 0;
    flops[i] = 0;
    bytes[i] = 0;
    launches[i] = 0;
  }
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (ProfRec* r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r->e0, r->e1) == cudaSuccess) {
      ms[r->cat] += t;
      flops[r->cat] += r->flops;
      bytes[r->cat] += r->bytes;
      launches[r->cat] += 1;
    }
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  g_prof.clear();
  return VMC_OK;
}

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
