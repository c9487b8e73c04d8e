// Library-level entry points: version, error string, launch counter, device probe, TMA
// descriptor encoding through the driver entry point (no link-time libcuda dependency).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "vy_common.cuh"

namespace vy {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Programmatic dependent launch: VY_PDL in the environment is the process default (off: no gain inside the captured TRAINING
// step, whose kernels are long), vy_set_pdl() overrides it at run time (the captured DECODE step turns it on: ~30 kernels of
// 4-20 us each, -21 us per step measured, profiles/r02_decode_attn_ab.txt).
static std::atomic<int> g_pdl_override{-1};
bool pdl_enabled() {
  static const bool env_on = getenv("VY_PDL") && atoi(getenv("VY_PDL")) != 0;
  const int o = g_pdl_override.load(std::memory_order_relaxed);
  return o < 0 ? env_on : o != 0;
}

}  // namespace vy
extern "C" int vy_set_pdl(int on) {
  const int prev = vy::g_pdl_override.exchange(on < 0 ? -1 : (on ? 1 : 0));
  return prev;
}
namespace vy {

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
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
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VY_ERR_CUDA;
  }
  // The driver entry point needs the primary context bound to the CALLING thread. A thread whose first CUDA
  // work is this call (autograd's backward worker, when a backward pass opens with a GEMM) has none yet:
  // cudaFree(nullptr) binds it (a no-op afterwards).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUtensorMapDataType dt =
      dtype == VY_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 == 3   ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                  : swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d base %p dims [%llu,%llu,%llu,%llu] "
        "stride1 %llu box [%u,%u,%u,%u]",
        static_cast<int>(r), rank, base, (unsigned long long)dims[0],
        (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
        (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 1 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return VY_ERR_CUDA;
  }
  return VY_OK;
}

struct TmapKey {
  const void* base;
  uint64_t d[5], s[5];
  uint32_t b[5];
  int dtype, swz, rank;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
  }
};

// A descriptor depends only on its key (pointer, geometry, box, swizzle), so reuse across calls is
// safe even when the caching allocator hands the same address to a different tensor.
int get_tensor_map_cached(CUtensorMap* out, int dtype, int rank, const void* base,
                          const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                          int swizzle) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.dtype = dtype;
  key.swz = swizzle;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) {
    key.d[i] = dims[i];
    key.s[i] = i > 0 ? strides_bytes[i] : 0;
    key.b[i] = box[i];
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return VY_OK;
    }
  }
  int rc = make_tensor_map(out, dtype, rank, base, dims, strides_bytes, box, swizzle);
  if (rc != VY_OK) return rc;
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return VY_OK;
}

}  // namespace vy

extern "C" {

int vy_version(void) { return VY_ABI_VERSION; }

const char* vy_last_error(void) { return vy::g_err; }

int64_t vy_launch_count(void) { return vy::g_launches.load(); }

int vy_abi_sizeof(const char* name) {
  if (!name) return -1;
#define VY_SZ(T) if (strcmp(name, #T) == 0) return static_cast<int>(sizeof(T))
  VY_SZ(VyGemm);
  VY_SZ(VyNorm);
  VY_SZ(VyAttn);
  VY_SZ(VyAttnBwd);
  VY_SZ(VyDecode);
  VY_SZ(VyEmbed);
  VY_SZ(VyPatchify);
  VY_SZ(VyCast4d);
  VY_SZ(VyXent);
  VY_SZ(VyAdamW);
  VY_SZ(VyRope);
  VY_SZ(VyRopeAppend);
  VY_SZ(VyDecodeLayer);
  VY_SZ(VyDecodeStep);
  VY_SZ(VyDpGroup);
  VY_SZ(VyDpReduce);
  VY_SZ(VyDpAdamW);
#undef VY_SZ
  return -1;
}

int vy_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return 0;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
