// TMA issue probe 2 (sm_100a): WHAT serialises bulk-tensor copies issued by one warp?  tools/tma_probe.cu found a fixed
// ~600-670 cycles per copy per issuing warp, independent of the box size and of the number of ring slots, and linear
// scaling over warps 0..3.  Hypotheses separated here, all on [128 x 64] bf16 SWIZZLE_128B boxes of one L2-resident
// 5 MB matrix:
//   lanes   : L lanes of ONE warp issue, each into its own ring            (per-thread or per-warp limit?)
//   spread  : issuing warps {0,1,2,3} vs {0,4,8,12} (same SM sub-partition) (per-scheduler limit?)
//   batch   : one thread issues B copies back to back, then waits for all B (does the wait->issue pattern matter?)
//   nowait  : one thread issues without ever waiting on a barrier (fire and forget; slots reused blindly)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_probe2.cu -o <bin> -lcuda
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

constexpr int kBox = 16384;
constexpr int kMaxIssuers = 16;

struct Cfg {
  int mode;      // 0: wait->issue ring per issuer, 1: batch (issue B, wait B), 2: no wait at all
  int warps;     // issuing warps
  int warp_stride;  // issuing warp ids = w * warp_stride
  int lanes;     // issuing lanes per warp (lane ids 0..lanes-1), each an independent issuer
  int slots;     // ring slots per issuer
  int iters;     // copies per issuer
};

__global__ void __launch_bounds__(512, 1)
probe(const __grid_constant__ CUtensorMap tmap, Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[kMaxIssuers * 12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxIssuers * 12; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp % c.warp_stride != 0 || warp / c.warp_stride >= c.warps || lane >= c.lanes) return;
  const int issuer = (warp / c.warp_stride) * c.lanes + lane;
  const int n_issuers = c.warps * c.lanes;
  const uint32_t ring = smem_u32(smem) + issuer * c.slots * kBox;
  const uint32_t bar0 = smem_u32(bars + issuer * 12);
  const int nbox_k = 64, row_tiles = 5;
  const long long t0 = clock64();
  if (c.mode == 0) {
    int slot = 0;
    uint32_t phase = 0;
    for (int it = 0; it < c.iters + c.slots; ++it) {
      if (it >= c.slots) mbar_wait(bar0 + 8 * slot, phase);
      if (it < c.iters) {
        const int t = it * n_issuers + issuer;
        mbar_expect_tx(bar0 + 8 * slot, kBox);
        tma_load_2d(ring + slot * kBox, &tmap, bar0 + 8 * slot, (t % nbox_k) * 64, ((t / nbox_k) % row_tiles) * 128);
      }
      if (++slot == c.slots) { slot = 0; if (it >= c.slots) phase ^= 1; }
    }
  } else if (c.mode == 1) {
    uint32_t phase = 0;
    for (int it = 0; it < c.iters; it += c.slots) {
      for (int s = 0; s < c.slots; ++s) {
        const int t = (it + s) * n_issuers + issuer;
        mbar_expect_tx(bar0 + 8 * s, kBox);
        tma_load_2d(ring + s * kBox, &tmap, bar0 + 8 * s, (t % nbox_k) * 64, ((t / nbox_k) % row_tiles) * 128);
      }
      for (int s = 0; s < c.slots; ++s) mbar_wait(bar0 + 8 * s, phase);
      phase ^= 1;
    }
  } else {
    // fire and forget: every copy arms slot (it % slots) again without waiting; the final wait covers only the tail
    for (int it = 0; it < c.iters; ++it) {
      const int s = it % c.slots;
      const int t = it * n_issuers + issuer;
      mbar_expect_tx(bar0 + 8 * s, kBox);
      tma_load_2d(ring + s * kBox, &tmap, bar0 + 8 * s, (t % nbox_k) * 64, ((t / nbox_k) % row_tiles) * 128);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * kMaxIssuers + issuer] = (unsigned long long)(t1 - t0);
  if (c.mode == 2) {  // let the in-flight copies land before the CTA (and its shared memory) goes away
    const long long t2 = clock64();
    while (clock64() - t2 < 200000) { }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fnp;
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(unsigned long long) * 256 * kMaxIssuers));
  const int d = 4096, rows = 640;
  void* a;
  CK(cudaMalloc(&a, (size_t)rows * d * 2));
  CK(cudaMemset(a, 0, (size_t)rows * d * 2));
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    fprintf(stderr, "encode failed\n");
    return 1;
  }
  auto run = [&](const char* label, Cfg c, int ctas) {
    const int n_issuers = c.warps * c.lanes;
    const size_t smem = 1024 + (size_t)n_issuers * c.slots * kBox;
    if (smem > 227 * 1024 || n_issuers > kMaxIssuers) { printf("{\"probe\": \"%s\", \"error\": \"config\"}\n", label); return; }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(d_out, 0, sizeof(unsigned long long) * 256 * kMaxIssuers));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<<<ctas, 512, smem>>>(tm, c, d_out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<<<ctas, 512, smem>>>(tm, c, d_out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> h(ctas * kMaxIssuers);
    CK(cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * ctas * kMaxIssuers, cudaMemcpyDeviceToHost));
    double cyc = 0;
    for (int i = 0; i < ctas; ++i) {
      unsigned long long m = 0;
      for (int w = 0; w < n_issuers; ++w) m = h[i * kMaxIssuers + w] > m ? h[i * kMaxIssuers + w] : m;
      cyc += (double)m;
    }
    cyc /= ctas;
    const double bytes = (double)c.iters * n_issuers * kBox;
    printf("{\"probe\": \"%s\", \"mode\": %d, \"ctas\": %d, \"warps\": %d, \"warp_stride\": %d, \"lanes\": %d, \"slots\": %d, "
           "\"cycles_per_copy_per_issuer\": %.0f, \"bytes_per_cycle_per_sm\": %.1f, \"ms\": %.4f}\n",
           label, c.mode, ctas, c.warps, c.warp_stride, c.lanes, c.slots, cyc / c.iters, bytes / cyc, ms);
    fflush(stdout);
  };
  const int it = 2048;
  run("ring_1warp_1lane", Cfg{0, 1, 1, 1, 6, it}, sms);
  run("ring_1warp_2lanes", Cfg{0, 1, 1, 2, 6, it}, sms);
  run("ring_1warp_4lanes", Cfg{0, 1, 1, 4, 3, it}, sms);
  run("ring_1warp_8lanes", Cfg{0, 1, 1, 8, 1, it}, sms);
  run("ring_4warps_adjacent", Cfg{0, 4, 1, 1, 3, it}, sms);
  run("ring_4warps_same_partition", Cfg{0, 4, 4, 1, 3, it}, sms);
  run("ring_2warps_adjacent", Cfg{0, 2, 1, 1, 6, it}, sms);
  run("ring_2warps_same_partition", Cfg{0, 2, 4, 1, 6, it}, sms);
  run("ring_8warps", Cfg{0, 8, 1, 1, 1, it}, sms);
  run("ring_12warps", Cfg{0, 12, 1, 1, 1, it}, sms);
  run("batch6_1warp_1lane", Cfg{1, 1, 1, 1, 6, it - it % 6}, sms);
  run("batch12_1warp_1lane", Cfg{1, 1, 1, 1, 12, it - it % 12}, sms);
  run("nowait_1warp_1lane", Cfg{2, 1, 1, 1, 12, it}, sms);
  run("nowait_1warp_1lane_1cta", Cfg{2, 1, 1, 1, 12, it}, 1);
  run("ring_4warps_adjacent_1cta", Cfg{0, 4, 1, 1, 3, it}, 1);
  return 0;
}
