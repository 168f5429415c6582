// Error plumbing and version of the C ABI (include/dropdec.h).
#include <cstdlib>
#include <string>

#include "dd_common.cuh"

namespace dd {
static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const char* what) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  if (!fn) {
    set_error("%s: cuTensorMapEncodeTiled unavailable (driver too old?)", what);
    return DD_ERR_CUDA;
  }
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
    return DD_ERR_CUDA;
  }
  return DD_OK;
}

// 0 / 1 once decided; -1 = not yet (then the DD_PDL environment variable, else off)
static int g_pdl = -1;
bool pdl_enabled(bool weight_gemm) {
  if (g_pdl < 0) {
    const char* e = getenv("DD_PDL");
    g_pdl = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return g_pdl == 1 || (g_pdl == 2 && weight_gemm) || (g_pdl == 3 && !weight_gemm);
}

static thread_local L2Prefetch g_l2_prefetch{};
L2Prefetch take_l2_prefetch() {
  const L2Prefetch pf = g_l2_prefetch;
  g_l2_prefetch = L2Prefetch{};
  return pf;
}
}  // namespace dd

extern "C" int dd_hint_l2_prefetch(const void* p0, size_t bytes0, const void* p1, size_t bytes1) {
  using namespace dd;
  DD_CHECK_ARG((bytes0 == 0 || p0) && (bytes1 == 0 || p1), "dd_hint_l2_prefetch: null range");
  DD_CHECK_ARG(((uintptr_t)p0 & 15) == 0 && ((uintptr_t)p1 & 15) == 0 && bytes0 % 16 == 0 && bytes1 % 16 == 0,
               "dd_hint_l2_prefetch: ranges must be 16-byte aligned and a multiple of 16 bytes long");
  g_l2_prefetch.ptr[0] = (const char*)p0; g_l2_prefetch.bytes[0] = bytes0;
  g_l2_prefetch.ptr[1] = (const char*)p1; g_l2_prefetch.bytes[1] = bytes1;
  return DD_OK;
}

extern "C" int dd_version(void) { return 100; }  // 0.1.0
extern "C" int dd_set_chain_launch(int on) {
  dd::pdl_enabled();   // (resolves the environment default)
  const int was = dd::g_pdl;
  dd::g_pdl = (on >= 1 && on <= 3) ? on : 0;
  return was;
}
extern "C" const char* dd_last_error(void) { return dd::g_last_error.c_str(); }
