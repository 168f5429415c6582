// TMA ingest probe (sm_100a): how many bytes per cycle does ONE SM take in through cp.async.bulk[.tensor], and what
// does it depend on?  Round 1 measured 34-36 B/clk for [128 x 64] bf16 SWIZZLE_128B boxes of an L2-resident matrix
// (tools/umma_probe.cu) -- the number both GEMM designs are built around.  This probe varies one thing at a time:
//   boxes in flight, box height, number of independent producer warps, swizzle / inner extent, plain 1-D bulk
//   copies, shared vs per-CTA-distinct L2-resident regions, and HBM streaming.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_probe.cu -o <bin> -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

constexpr int kMaxSlots = 12;
constexpr int kWarps = 4;

struct Cfg {
  int kind;         // 0: 2-D tensor boxes, 1: 1-D bulk copies
  int warps;        // independent producer warps (each its own ring)
  int slots;        // boxes in flight per warp
  int box_bytes;    // bytes per box
  int box_rows;     // rows per box (kind 0)
  int inner;        // elements per box row (kind 0)
  int iters;        // boxes per warp
  int nbox_k;       // boxes along the inner dimension
  int row_tiles;    // box rows along the outer dimension (per region)
  long long cta_stride_rows;   // kind 0: outer-dimension offset of CTA c = c * cta_stride_rows (0 = every CTA the same region)
  long long cta_stride_bytes;  // kind 1
};

__global__ void __launch_bounds__(32 * kWarps, 1)
probe(const __grid_constant__ CUtensorMap tmap, const uint8_t* base, Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[kWarps * kMaxSlots];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWarps * kMaxSlots; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= c.warps) return;
  const uint32_t ring = smem_u32(smem) + warp * c.slots * c.box_bytes;
  const uint32_t bar0 = smem_u32(bars + warp * kMaxSlots);
  const long long t0 = clock64();
  int slot = 0;
  uint32_t phase = 0;
  for (int it = 0; it < c.iters + c.slots; ++it) {
    if (it >= c.slots) mbar_wait(bar0 + 8 * slot, phase);
    if (it < c.iters && elect_one()) {
      const int t = it * c.warps + warp;
      mbar_expect_tx(bar0 + 8 * slot, c.box_bytes);
      if (c.kind == 0) {
        const long long row = (long long)blockIdx.x * c.cta_stride_rows + (long long)((t / c.nbox_k) % c.row_tiles) * c.box_rows;
        tma_load_2d(ring + slot * c.box_bytes, &tmap, bar0 + 8 * slot, (t % c.nbox_k) * c.inner, (int)row);
      } else {
        const uint8_t* src = base + (long long)blockIdx.x * c.cta_stride_bytes +
                             (long long)(t % (c.nbox_k * c.row_tiles)) * c.box_bytes;
        bulk_load_1d(ring + slot * c.box_bytes, src, c.box_bytes, bar0 + 8 * slot);
      }
    }
    __syncwarp();
    if (++slot == c.slots) { slot = 0; if (it >= c.slots) phase ^= 1; }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * kWarps + warp] = (unsigned long long)(t1 - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static void run(const char* label, void* a, long long rows_total, int d, Cfg c, int ctas, CUtensorMapSwizzle sw,
                unsigned long long* d_out) {
  CUtensorMap tm{};
  if (c.kind == 0) {
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {(cuuint32_t)c.inner, (cuuint32_t)c.box_rows}, es[2] = {1, 1};
    CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("{\"probe\": \"%s\", \"error\": \"encode failed %d\"}\n", label, (int)r);
      return;
    }
  }
  const size_t smem = 1024 + (size_t)c.warps * c.slots * c.box_bytes;
  if (smem > 227 * 1024) {
    printf("{\"probe\": \"%s\", \"error\": \"smem %zu\"}\n", label, smem);
    return;
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(unsigned long long) * 256 * kWarps));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  probe<<<ctas, 32 * kWarps, smem>>>(tm, (const uint8_t*)a, c, d_out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<<<ctas, 32 * kWarps, smem>>>(tm, (const uint8_t*)a, c, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> h(ctas * kWarps);
  CK(cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * ctas * kWarps, cudaMemcpyDeviceToHost));
  double cyc = 0;
  for (int i = 0; i < ctas; ++i) {
    unsigned long long m = 0;
    for (int w = 0; w < c.warps; ++w) m = h[i * kWarps + w] > m ? h[i * kWarps + w] : m;
    cyc += (double)m;
  }
  cyc /= ctas;
  const double bytes = (double)c.iters * c.warps * c.box_bytes;
  printf("{\"probe\": \"%s\", \"kind\": \"%s\", \"ctas\": %d, \"warps\": %d, \"slots\": %d, \"box_bytes\": %d, \"box\": \"%dx%d\", "
         "\"ms\": %.4f, \"bytes_per_cycle_per_sm\": %.1f, \"chip_TBps\": %.2f}\n",
         label, c.kind ? "bulk1d" : "tensor2d", ctas, c.warps, c.slots, c.box_bytes, c.box_rows, c.inner, ms, bytes / cyc,
         bytes * ctas / (ms * 1e-3) / 1e12);
  fflush(stdout);
}

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  g_encode = (EncodeTiledFn)fnp;
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(unsigned long long) * 256 * kWarps));
  const int d = 4096;
  const int iters = 4096;
  // ---- (A) one L2-resident 5 MB matrix shared by every SM (the activation-tile pattern) ----
  {
    const long long rows = 640;
    void* a;
    CK(cudaMalloc(&a, rows * d * 2));
    CK(cudaMemset(a, 0, rows * d * 2));
    auto T = [&](const char* label, int warps, int slots, int box_rows, int inner, CUtensorMapSwizzle sw, int ctas) {
      Cfg c{0, warps, slots, box_rows * inner * 2, box_rows, inner, iters / warps, d / inner, (int)(rows / box_rows), 0, 0};
      run(label, a, rows, d, c, ctas, sw, d_out);
    };
    T("shared5MB_base", 1, 6, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_base_1cta", 1, 6, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, 1);
    T("shared5MB_slots12", 1, 12, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_slots3", 1, 3, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_2warps", 2, 6, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_4warps", 4, 3, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_4warps_1cta", 4, 3, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B, 1);
    T("shared5MB_box256rows", 1, 6, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_box64rows", 1, 12, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, sms);
    T("shared5MB_noswz_64x128", 1, 6, 128, 64, CU_TENSOR_MAP_SWIZZLE_NONE, sms);
    T("shared5MB_noswz_256x32", 1, 6, 32, 256, CU_TENSOR_MAP_SWIZZLE_NONE, sms);
    T("shared5MB_swz64_32x128", 1, 12, 128, 32, CU_TENSOR_MAP_SWIZZLE_64B, sms);
    // 1-D bulk copies of the same bytes (16 KB / 32 KB contiguous chunks)
    {
      Cfg c{1, 1, 6, 16384, 0, 0, iters, 64, 5, 0, 0};
      run("shared5MB_bulk1d_16K", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    {
      Cfg c{1, 1, 6, 32768, 0, 0, iters / 2, 32, 5, 0, 0};
      run("shared5MB_bulk1d_32K", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    {
      Cfg c{1, 2, 6, 16384, 0, 0, iters / 2, 64, 5, 0, 0};
      run("shared5MB_bulk1d_16K_2warps", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    {
      Cfg c{1, 1, 6, 16384, 0, 0, iters, 64, 5, 0, 0};
      run("shared5MB_bulk1d_16K_1cta", a, rows, d, c, 1, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    CK(cudaFree(a));
  }
  // ---- (B) every CTA its own L2-resident region: 148 x 512 KB = 74 MB (< 126 MB L2), read 128 times over ----
  {
    const long long rows_per = 64;  // 64 rows x 8 KB = 512 KB per CTA
    const long long rows = rows_per * sms;
    void* a;
    CK(cudaMalloc(&a, rows * d * 2));
    CK(cudaMemset(a, 0, rows * d * 2));
    {
      Cfg c{0, 1, 6, 16384, 64, 128 / 1, iters, 0, 0, 0, 0};
      c.box_rows = 64; c.inner = 64; c.box_bytes = 64 * 64 * 2; c.slots = 12; c.nbox_k = d / 64; c.row_tiles = 1;
      c.cta_stride_rows = rows_per;
      run("distinctL2_box64x64rows", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_128B, d_out);
    }
    {
      Cfg c{1, 1, 6, 16384, 0, 0, iters, 32, 1, 0, rows_per * d * 2};
      run("distinctL2_bulk1d_16K", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    {
      Cfg c{1, 2, 6, 16384, 0, 0, iters / 2, 32, 1, 0, rows_per * d * 2};
      run("distinctL2_bulk1d_16K_2warps", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    CK(cudaFree(a));
  }
  // ---- (C) HBM streaming: every CTA a distinct 7 x 128-row region per pass (1 GB total) ----
  {
    const long long rows_per = 7 * 128;
    const long long rows = rows_per * sms;
    void* a;
    CK(cudaMalloc(&a, rows * d * 2));
    CK(cudaMemset(a, 0, rows * d * 2));
    auto T = [&](const char* label, int warps, int slots, int box_rows) {
      Cfg c{0, warps, slots, box_rows * 64 * 2, box_rows, 64, 7 * 128 / box_rows * 64 / warps, d / 64, (int)(rows_per / box_rows),
            rows_per, 0};
      run(label, a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_128B, d_out);
    };
    T("hbm_tensor_base", 1, 6, 128);
    T("hbm_tensor_slots12", 1, 12, 128);
    T("hbm_tensor_2warps", 2, 6, 128);
    {
      Cfg c{1, 1, 12, 16384, 0, 0, 7 * 64, 7 * 64, 1, 0, rows_per * d * 2};
      run("hbm_bulk1d_16K", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    {
      Cfg c{1, 1, 6, 32768, 0, 0, 7 * 32, 7 * 32, 1, 0, rows_per * d * 2};
      run("hbm_bulk1d_32K", a, rows, d, c, sms, CU_TENSOR_MAP_SWIZZLE_NONE, d_out);
    }
    CK(cudaFree(a));
  }
  return 0;
}
