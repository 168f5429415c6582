// Micro-probe for the uncertainty-stage design (sm_100a): what do tcgen05.mma (SS operands, SWIZZLE_128B,
// K-major, cta_group::1) and TMA tile loads actually sustain when all 148 SMs do the same thing?
//   mode 0: MMA only        -- every CTA re-issues M=128 x N x K=16 MMAs on resident smem tiles
//   mode 1: TMA only        -- every CTA streams [128 x 64] bf16 boxes of one L2-resident matrix (same
//                              addresses on every SM, like the activation tiles of the uncertainty kernel)
//   mode 2: both at once    -- interference between the operand reads of the tensor core and TMA writes
//   mode 3: TMA only, every CTA streams a DIFFERENT region (no sharing)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/umma_probe.cu -o <bin> -lcuda
// Prints one JSON line per configuration (cycles per MMA, bytes per cycle per SM, chip TB/s).
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

constexpr int kSlotBytes = 16384;  // one [128 x 64] bf16 box
constexpr int kTmaSlots = 6;

struct ProbeOut {
  unsigned long long mma_cycles, tma_cycles;
};

// smem: [0,16K) A tile | [16K,48K) B tile (up to 256 rows) | TMA ring
template <int N>
__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, int mode, int mma_iters, int tma_iters, int rows_total,
             ProbeOut* out, int commit_every, int n_acc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[kTmaSlots + 2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i <= kTmaSlots + 1; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand tiles (finite numbers)
  for (int i = threadIdx.x; i < 3 * kSlotBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 && (mode == 1 || mode == 2 || mode == 3)) {
    // ---- TMA streamer: ring of kTmaSlots boxes, wait for landing then immediately reissue ----
    const uint32_t ring = base + 3 * kSlotBytes;
    const int nbox_k = 64;  // 4096 / 64
    const int row_tiles = rows_total / 128;
    const long long t0 = clock64();
    int slot = 0;
    uint32_t phase = 0;
    const int cta_off = mode == 3 ? blockIdx.x * 7 * 64 : 0;
    for (int it = 0; it < tma_iters + kTmaSlots; ++it) {
      if (it >= kTmaSlots) mbar_wait(bar0 + 8 * slot, phase);  // previous use of this slot landed
      if (it < tma_iters && elect_one()) {
        const int t = it + cta_off;
        mbar_expect_tx(bar0 + 8 * slot, kSlotBytes);
        tma_load_2d(ring + slot * kSlotBytes, &tmap, bar0 + 8 * slot, (t % nbox_k) * 64, ((t / nbox_k) % row_tiles) * 128);
      }
      __syncwarp();
      if (++slot == kTmaSlots) { slot = 0; if (it >= kTmaSlots) phase ^= 1; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].tma_cycles = (unsigned long long)(t1 - t0);
  } else if (warp == 1 && (mode == 0 || mode == 2)) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_desc = make_smem_desc(base), b_desc = make_smem_desc(base + kSlotBytes);
    const uint32_t bar_done = bar0 + 8 * kTmaSlots;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < mma_iters; it += 64) {
      if (elect_one()) {
        for (int j = 0; j < 64; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + ((j % n_acc) * N), a_desc + 2 * k, b_desc + 2 * k, idesc, 1);
          // a commit to a barrier nobody waits on: does the arrive-on-completion stall the tensor pipe?
          if (commit_every && (j % commit_every) == commit_every - 1) umma_commit(bar_done + 8);
        }
        umma_commit(bar_done);
      }
      __syncwarp();
      mbar_wait(bar_done, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x].mma_cycles = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- mode 4: cluster multicast.  Each CTA of a cluster of C issues 1/C of every [128 x 64] box (a
// [128/C x 64] slice) with .multicast::cluster to all C CTAs; a slot is rewritten only after all C CTAs
// released it (remote mbarrier arrives), i.e. the real producer/consumer protocol of a multicast GEMM.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

constexpr int kMcSlots = 12;
template <int C>
__global__ void __launch_bounds__(128, 1)
probe_mc_kernel(const __grid_constant__ CUtensorMap tmap_slice, int tma_iters, int rows_total, ProbeOut* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[kMcSlots];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t ring = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMcSlots; ++i) mbar_init(full0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int nbox_k = 64, row_tiles = rows_total / 128;
  const long long t0 = clock64();
  uint32_t phase = 0;
  // rounds of kMcSlots boxes: every CTA issues its 1/C slice of each box (multicast to the whole cluster),
  // waits until all kMcSlots boxes have fully landed in ITS smem, then the cluster synchronises (so that
  // no CTA re-arms a slot while a peer's slice of the previous round is still in flight)
  for (int it0 = 0; it0 < tma_iters; it0 += kMcSlots) {
    if (threadIdx.x < 32) {
      for (int sl = 0; sl < kMcSlots; ++sl) {
        const int it = it0 + sl;
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * sl, kSlotBytes);
          tma_load_2d_mc(ring + sl * kSlotBytes + rank * (kSlotBytes / C), &tmap_slice, full0 + 8 * sl, (it % nbox_k) * 64,
                         ((it / nbox_k) % row_tiles) * 128 + rank * (128 / C), (uint16_t)((1u << C) - 1));
        }
        __syncwarp();
      }
      for (int sl = 0; sl < kMcSlots; ++sl) mbar_wait(full0 + 8 * sl, phase);
    }
    phase ^= 1;
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x].tma_cycles = (unsigned long long)(t1 - t0);
}

// ---- mode 5: the LSU path.  NT threads stream the same [128 x 64] boxes with cp.async (16 B per thread per
// request, manual XOR swizzle like SWIZZLE_128B), 4 groups in flight.
template <int NT>
__global__ void __launch_bounds__(NT, 1) probe_cpasync_kernel(const __nv_bfloat16* a, int d, int iters, int rows_total,
                                                              ProbeOut* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t ring = smem_u32(smem);
  const int nbox_k = d / 64, row_tiles = rows_total / 128;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int kb = it % nbox_k, rt = (it / nbox_k) % row_tiles;
    const uint32_t dst = ring + (it % kTmaSlots) * kSlotBytes;
#pragma unroll
    for (int i = 0; i < 1024 / NT; ++i) {
      const int idx = i * NT + threadIdx.x, row = idx >> 3, c = idx & 7;
      const __nv_bfloat16* src = a + (size_t)(rt * 128 + row) * d + kb * 64 + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + row * 128 + ((c ^ (row & 7)) << 4)), "l"(src));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group %0;" ::"n"(kTmaSlots - 2));
  }
  asm volatile("cp.async.wait_group 0;");
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x].tma_cycles = (unsigned long long)(t1 - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N>
static void run(const CUtensorMap& tm, int mode, int sms, int rows_total, ProbeOut* d_out, const char* label,
                int commit_every = 0, int n_acc = 2) {
  const int mma_iters = 4096, tma_iters = 4096;
  const size_t smem = 1024 + (3 + kTmaSlots) * kSlotBytes;
  CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * sms));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  probe_kernel<N><<<sms, 192, smem>>>(tm, mode, mma_iters, tma_iters, rows_total, d_out, commit_every, n_acc);  // warm
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe_kernel<N><<<sms, 192, smem>>>(tm, mode, mma_iters, tma_iters, rows_total, d_out, commit_every, n_acc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<ProbeOut> h(sms);
  CK(cudaMemcpy(h.data(), d_out, sizeof(ProbeOut) * sms, cudaMemcpyDeviceToHost));
  double mc = 0, tc = 0;
  for (auto& o : h) { mc += o.mma_cycles; tc += o.tma_cycles; }
  mc /= sms; tc /= sms;
  const double cyc_per_mma = mc / (mma_iters * 4.0);
  const double bpc = tc > 0 ? (double)tma_iters * kSlotBytes / tc : 0;
  const double tflops = mc > 0 ? 2.0 * 128 * N * 16 * mma_iters * 4.0 * sms / (ms * 1e-3) / 1e12 : 0;
  const double tbs = tc > 0 ? (double)tma_iters * kSlotBytes * sms / (ms * 1e-3) / 1e12 : 0;
  printf("{\"probe\": \"%s\", \"commit_every_4mma_groups\": %d, \"accumulators\": %d, \"N\": %d, \"mode\": %d, \"ctas\": %d, \"ms\": %.4f, \"cycles_per_mma\": %.1f, "
         "\"ideal_cycles_per_mma\": %.1f, \"mma_tflops_if_alone\": %.1f, \"tma_bytes_per_cycle_per_sm\": %.1f, "
         "\"tma_chip_TBps\": %.2f}\n",
         label, commit_every, n_acc, N, mode, sms, ms, cyc_per_mma, N / 2.0, tflops, bpc, tbs);
  fflush(stdout);
}

template <int C>
static void run_mc(EncodeTiledFn encode, void* a, int d, int rows_total, int sms, ProbeOut* d_out) {
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  const cuuint32_t box[2] = {64, 128 / C}, es[2] = {1, 1};
  if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
      CUDA_SUCCESS) {
    fprintf(stderr, "encode (mc) failed\n");
    return;
  }
  const int tma_iters = 4092;  // multiple of kMcSlots
  const size_t smem = 1024 + kMcSlots * kSlotBytes;
  CK(cudaFuncSetAttribute(probe_mc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sms / C * C;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, probe_mc_kernel<C>, &cfg);
  if (max_clusters * C < grid) { grid = max_clusters * C; cfg.gridDim = dim3(grid); }
  CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * 256));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, probe_mc_kernel<C>, tm, tma_iters, rows_total, d_out));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, probe_mc_kernel<C>, tm, tma_iters, rows_total, d_out));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<ProbeOut> h(grid);
  CK(cudaMemcpy(h.data(), d_out, sizeof(ProbeOut) * grid, cudaMemcpyDeviceToHost));
  double tc = 0;
  for (auto& o : h) tc += o.tma_cycles;
  tc /= grid;
  printf("{\"probe\": \"tma_multicast_shared\", \"cluster\": %d, \"ctas\": %d, \"max_active_clusters\": %d, \"ms\": %.4f, "
         "\"landed_bytes_per_cycle_per_sm\": %.1f, \"issued_bytes_per_cycle_per_sm\": %.1f, \"landed_chip_TBps\": %.2f}\n",
         C, grid, max_clusters, ms, (double)tma_iters * kSlotBytes / tc, (double)tma_iters * kSlotBytes / C / tc,
         (double)tma_iters * kSlotBytes * grid / (ms * 1e-3) / 1e12);
  fflush(stdout);
}

template <int NT>
static void run_cpasync(const __nv_bfloat16* a, int d, int rows_total, int sms, ProbeOut* d_out) {
  const int iters = 4096;
  const size_t smem = 1024 + kTmaSlots * kSlotBytes;
  CK(cudaFuncSetAttribute(probe_cpasync_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe_cpasync_kernel<NT><<<sms, NT, smem>>>(a, d, iters, rows_total, d_out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe_cpasync_kernel<NT><<<sms, NT, smem>>>(a, d, iters, rows_total, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<ProbeOut> h(sms);
  CK(cudaMemcpy(h.data(), d_out, sizeof(ProbeOut) * sms, cudaMemcpyDeviceToHost));
  double tc = 0;
  for (auto& o : h) tc += o.tma_cycles;
  tc /= sms;
  printf("{\"probe\": \"cp_async_shared_5MB\", \"threads\": %d, \"ctas\": %d, \"ms\": %.4f, \"bytes_per_cycle_per_sm\": %.1f, "
         "\"chip_TBps\": %.2f}\n", NT, sms, ms, (double)iters * kSlotBytes / tc, (double)iters * kSlotBytes * sms / (ms * 1e-3) / 1e12);
  fflush(stdout);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fnp;
  ProbeOut* d_out;
  CK(cudaMalloc(&d_out, sizeof(ProbeOut) * 256));
  for (int rows_total : {640, 148 * 7 * 128}) {  // 5 MB (L2-resident, shared) | 1 GB-ish distinct regions
    const int d = 4096;
    __nv_bfloat16* a;
    const size_t bytes = (size_t)rows_total * d * 2;
    if (rows_total > 640 && bytes > (size_t)2 << 30) continue;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMemset(a, 0, bytes));
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
        CUDA_SUCCESS) {
      fprintf(stderr, "encode failed\n");
      return 1;
    }
    if (rows_total == 640) {
      run<96>(tm, 0, sms, rows_total, d_out, "mma_only");
      run<128>(tm, 0, sms, rows_total, d_out, "mma_only");
      run<192>(tm, 0, sms, rows_total, d_out, "mma_only");
      run<256>(tm, 0, sms, rows_total, d_out, "mma_only");
      run<96>(tm, 0, 1, rows_total, d_out, "mma_only_1cta");
      run<96>(tm, 0, sms, rows_total, d_out, "mma_commit_each_tile", 1, 2);
      run<96>(tm, 0, sms, rows_total, d_out, "mma_5_accumulators", 0, 5);
      run<96>(tm, 0, sms, rows_total, d_out, "mma_commit_each_tile_5acc", 1, 5);
      run<96>(tm, 0, sms, rows_total, d_out, "mma_commit_every_5_tiles_5acc", 5, 5);
      run<256>(tm, 0, sms, rows_total, d_out, "mma_commit_each_tile", 1, 2);
      run<96>(tm, 1, sms, rows_total, d_out, "tma_shared_5MB");
      run<96>(tm, 1, 1, rows_total, d_out, "tma_shared_5MB_1cta");
      run<96>(tm, 1, 16, rows_total, d_out, "tma_shared_5MB_16cta");
      run<96>(tm, 2, sms, rows_total, d_out, "mma+tma_shared");
      run<256>(tm, 2, sms, rows_total, d_out, "mma+tma_shared");
      run_cpasync<128>(a, d, rows_total, sms, d_out);
      run_cpasync<256>(a, d, rows_total, sms, d_out);
      run_cpasync<512>(a, d, rows_total, sms, d_out);
      run_mc<1>(encode, a, d, rows_total, sms, d_out);
      run_mc<2>(encode, a, d, rows_total, sms, d_out);
      run_mc<4>(encode, a, d, rows_total, sms, d_out);
      run_mc<8>(encode, a, d, rows_total, sms, d_out);
    } else {
      run<96>(tm, 3, sms, rows_total, d_out, "tma_distinct_hbm");
    }
    CK(cudaFree(a));
  }
  return 0;
}
