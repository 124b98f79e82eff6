// Host-side utilities: thread-local error message, device queries, TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace mimo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// which tensor-core kernel the last conv launcher of this thread dispatched to (read by the executor's per-launch profiler)
static thread_local int g_last_kernel = 0;
void note_kernel(int id) { g_last_kernel = id; }
int last_kernel() { return g_last_kernel; }
const char* conv_kernel_name(int id) {
  static const char* names[] = {"", "conv3x3_flat_kernel", "conv3x3_flatk_kernel", "conv3x3_igemm_kernel", "conv3x3_wgrad_flat_kernel",
                                "conv3x3_wgrad_flatk_kernel", "conv3x3_wgrad_kernel", "conv3x3_c2_kernel", "conv3x3_wgrad_c2_kernel", "conv3x3_flat2_kernel", "conv3x3_wgrad_flat2_kernel",
                                "conv3x3_thin_kernel"};
  return (id >= 0 && id < 12) ? names[id] : "";
}

int num_sms() {
  static int sms = 0;   // every GPU of a node is the same part: the count of the first device queried serves them all
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Encoded maps are a pure function of (base, rank, dims, strides, box, swizzle): eager launches (no CUDA graph: profiling, dropout
// masks, the component-level entry points) re-use them instead of paying five driver calls per convolution launch.
namespace {
struct TmapCache {
  std::mutex mu;
  std::unordered_map<std::string, CUtensorMap> map;
};
TmapCache& tmap_cache() {
  static TmapCache c;
  return c;
}
}  // namespace

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  MIMO_CHECK(fn != nullptr, MIMO_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point not available");
  MIMO_CHECK(rank >= 1 && rank <= 5, MIMO_ERR_ARG, "encode_tmap: rank %d", rank);
  struct Key { const void* base; int rank, swz; uint64_t dims[5], strides[4]; uint32_t box[5]; } key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.swz = swizzle128;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; if (i + 1 < rank) key.strides[i] = strides_bytes[i]; }
  const std::string k(reinterpret_cast<const char*>(&key), sizeof(key));
  {
    TmapCache& c = tmap_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.map.find(k);
    if (it != c.map.end()) { *out = it->second; return MIMO_OK; }
    if (c.map.size() > 8192) c.map.clear();   // bounded: a plan uses a few hundred maps
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u, stride0 %llu)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, (unsigned long long)strides_bytes[0]);
    return MIMO_ERR_CUDA;
  }
  {
    TmapCache& c = tmap_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    c.map.emplace(k, *out);
  }
  return MIMO_OK;
}

}  // namespace mimo
