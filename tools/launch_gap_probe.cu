// What does one kernel node of a CUDA graph cost on B200 when the kernel itself does nothing?  Variants: grid size,
// dynamic shared memory, TMEM alloc/dealloc, 128-byte __grid_constant__ tensor maps, thread-block clusters, and a
// kernel that ends with outstanding global stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/launch_gap_probe.cu -o tools/_bin/launch_gap_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

struct Big { unsigned char b[256]; };

__global__ void k_empty(int* sink) {
  if (sink && threadIdx.x == 1000) *sink = 1;
}
__global__ void k_params(const __grid_constant__ Big a, int* sink) {
  if (sink && threadIdx.x == 1000) *sink = a.b[3];
}
__global__ void k_smem(int* sink) {
  extern __shared__ unsigned char sm[];
  if (sink && threadIdx.x == 1000) *sink = sm[5];
}
__global__ void k_tmem(int* sink, int cols) {
  __shared__ uint32_t slot;
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)),
                 "r"((uint32_t)cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"((uint32_t)cols) : "memory");
  if (sink && threadIdx.x == 1000) *sink = 1;
}
// a kernel that is busy for `ns` nanoseconds (globaltimer) and then exits: is the boundary between two LONG kernels as
// cheap as between two empty ones?  out[2*blockIdx.x] = entry time, out[2*blockIdx.x+1] = exit time
__global__ void k_spin(long long ns, long long* out, int slot) {
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0) :: "memory");
  long long t = t0;
  while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
  if (threadIdx.x == 0 && out) { out[(size_t)slot * 512 + 2 * blockIdx.x] = t0; out[(size_t)slot * 512 + 2 * blockIdx.x + 1] = t; }
}
// the same busy kernel as a programmatic dependent launch: entry stamp, griddepcontrol.wait, post-wait stamp, spin, exit
// stamp.  trigger: 0 = never (implicit at exit), 1 = at entry, 2 = after the spin
__global__ void k_spin_pdl(long long ns, long long* out, int slot, int trigger) {
  long long te, t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te) :: "memory");
  if (trigger == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0) :: "memory");
  long long t = t0;
  while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
  if (trigger == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0 && out) {
    out[(size_t)slot * 512 + 3 * blockIdx.x] = te; out[(size_t)slot * 512 + 3 * blockIdx.x + 1] = t0; out[(size_t)slot * 512 + 3 * blockIdx.x + 2] = t;
  }
}
__global__ void k_store(float* out, int n_per_thread) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < n_per_thread; ++j) out[i + (size_t)j * gridDim.x * blockDim.x] = (float)j;
}

template <typename F>
static float time_graph(cudaStream_t st, int n, F&& launch) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int r = 0; r < n; ++r) launch();
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGraphExecDestroy(ge));
  CK(cudaGraphDestroy(g));
  return best * 1e3f / n;
}

int main() {
  CK(cudaSetDevice(0));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const int n = 200;
  float* out;
  CK(cudaMalloc(&out, 256 << 20));
  Big big{};
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  auto report = [&](const char* name, float us) { printf("{\"probe\": \"%s\", \"us_per_node\": %.2f}\n", name, us); fflush(stdout); };
  report("empty_1x32", time_graph(st, n, [&] { k_empty<<<1, 32, 0, st>>>(nullptr); }));
  report("empty_148x416", time_graph(st, n, [&] { k_empty<<<148, 416, 0, st>>>(nullptr); }));
  report("empty_148x1024", time_graph(st, n, [&] { k_empty<<<148, 1024, 0, st>>>(nullptr); }));
  report("params256B_148x416", time_graph(st, n, [&] { k_params<<<148, 416, 0, st>>>(big, nullptr); }));
  report("smem200KB_148x416", time_graph(st, n, [&] { k_smem<<<148, 416, 200 * 1024, st>>>(nullptr); }));
  report("tmem512_148x416", time_graph(st, n, [&] { k_tmem<<<148, 416, 0, st>>>(nullptr, 512); }));
  report("tmem128_148x416", time_graph(st, n, [&] { k_tmem<<<148, 416, 0, st>>>(nullptr, 128); }));
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(416); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    report("cluster2_148x416", time_graph(st, n, [&] { CK(cudaLaunchKernelEx(&cfg, k_empty, (int*)nullptr)); }));
    cfg.gridDim = dim3(128);
    attr[0].val.clusterDim.x = 4;
    report("cluster4_128x416", time_graph(st, n, [&] { CK(cudaLaunchKernelEx(&cfg, k_empty, (int*)nullptr)); }));
  }
  {
    long long* stamps;
    CK(cudaMalloc(&stamps, 64 * 512 * 8));
    for (long long ns : {2000ll, 20000ll}) {
      for (int smem : {0, 200 * 1024}) {
        CK(cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CK(cudaMemset(stamps, 0, 64 * 512 * 8));
        int slot = 0;
        const float us = time_graph(st, 32, [&] { k_spin<<<148, 416, smem, st>>>(ns, stamps, slot); slot = (slot + 1) % 32; });
        std::vector<long long> h(32 * 512);
        CK(cudaMemcpy(h.data(), stamps, 32 * 512 * 8, cudaMemcpyDeviceToHost));
        double gap = 0;
        for (int k = 1; k < 32; ++k) {
          long long last_exit = 0, first_entry = h[k * 512];
          for (int c = 0; c < 148; ++c) { last_exit = std::max(last_exit, h[(k - 1) * 512 + 2 * c + 1]); first_entry = std::min(first_entry, h[k * 512 + 2 * c]); }
          gap += (first_entry - last_exit) * 1e-3;
        }
        printf("{\"probe\": \"spin_%lldus_smem%dKB\", \"us_per_node\": %.2f, \"gap_last_exit_to_first_entry_us\": %.2f}\n", ns / 1000, smem / 1024, us, gap / 31);
      }
    }
  }
  {
    long long* stamps;
    CK(cudaMalloc(&stamps, 64 * 512 * 8));
    CK(cudaFuncSetAttribute(k_spin_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (long long ns : {2000ll, 20000ll}) {
      for (int smem : {0, 200 * 1024}) {
        for (int trigger : {0, 1, 2}) {
          CK(cudaMemset(stamps, 0, 64 * 512 * 8));
          int slot = 0;
          cudaLaunchConfig_t cfg{};
          cfg.gridDim = dim3(148); cfg.blockDim = dim3(416); cfg.dynamicSmemBytes = smem; cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          const float us = time_graph(st, 32, [&] { CK(cudaLaunchKernelEx(&cfg, k_spin_pdl, ns, stamps, slot, trigger)); slot = (slot + 1) % 32; });
          std::vector<long long> h(32 * 512);
          CK(cudaMemcpy(h.data(), stamps, 32 * 512 * 8, cudaMemcpyDeviceToHost));
          double gap = 0, early = 0;
          for (int k = 1; k < 32; ++k) {
            long long last_exit = 0, first_go = h[k * 512 + 1], first_entry = h[k * 512];
            for (int c = 0; c < 148; ++c) {
              last_exit = std::max(last_exit, h[(k - 1) * 512 + 3 * c + 2]);
              first_go = std::min(first_go, h[k * 512 + 3 * c + 1]);
              first_entry = std::min(first_entry, h[k * 512 + 3 * c]);
            }
            gap += (first_go - last_exit) * 1e-3;
            early += (last_exit - first_entry) * 1e-3;
          }
          printf("{\"probe\": \"pdl_spin_%lldus_smem%dKB_trigger%d\", \"us_per_node\": %.2f, \"gap_last_exit_to_first_go_us\": %.2f, "
                 "\"entry_before_prev_exit_us\": %.2f}\n", ns / 1000, smem / 1024, trigger, us, gap / 31, early / 31);
          fflush(stdout);
        }
      }
    }
  }
  report("store_148x416_x1", time_graph(st, n, [&] { k_store<<<148, 416, 0, st>>>(out, 1); }));
  report("store_148x416_x64(16MB)", time_graph(st, n, [&] { k_store<<<148, 416, 0, st>>>(out, 64); }));
  return 0;
}
