// Stand-alone timing + sanity harness for dd_stream_gemm (through the C ABI of libdropdec.so), next to cuBLAS.
//   tools/_bin/sgemm_bench [M ...]          (default M = 24 64 128 192 256)
// For every decoder GEMM shape of LLaVA-1.5-7B and every M: (1) the result is compared element by element with a
// cuBLAS fp32 GEMM pushed through the same epilogue (differences of one bf16 ulp come from the summation order);
// (2) `reps` launches over rotating weight copies (> L2) are captured in a CUDA graph and timed with CUDA events.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/sgemm_bench.cu -o tools/_bin/sgemm_bench
//        -Ldropoutdecoding_b200/_build -ldropdec -lcublas -Xlinker -rpath -Xlinker '$ORIGIN/../../dropoutdecoding_b200/_build'
#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "../include/dropdec.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)
#define CB(x)                                                             \
  do {                                                                    \
    cublasStatus_t s_ = (x);                                              \
    if (s_ != CUBLAS_STATUS_SUCCESS) {                                    \
      fprintf(stderr, "%s failed: %d (line %d)\n", #x, (int)s_, __LINE__); \
      exit(1);                                                            \
    }                                                                     \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn(((int)(h & 0xffff) - 32768) / 32768.0f * scale);
  }
}

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// expected output from the fp32 GEMM c[M, w_rows] with the library's epilogue; counts elements further than
// `ulps` bf16 ulps (of the larger magnitude) + 1e-6 apart
__global__ void check(const float* c, const void* y, const __nv_bfloat16* res, int M, int N, int epi, int w_rows,
                      unsigned long long* bad, float* maxerr) {
  const size_t n = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), f = (int)(i % N);
    float want, got;
    if (epi == 3) {
      const float g = bf16r(c[(size_t)m * w_rows + f]), u = bf16r(c[(size_t)m * w_rows + N + f]);
      want = bf16r(bf16r(g / (1.f + __expf(-g))) * u);
    } else {
      const float r0 = bf16r(c[(size_t)m * w_rows + f]);
      want = epi == 1 ? bf16r(r0 + __bfloat162float(res[i])) : r0;
    }
    got = epi == 2 ? reinterpret_cast<const float*>(y)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(y)[i]);
    const float d = fabsf(want - got);
    const float mag = fmaxf(fabsf(want), fabsf(got));
    const float ulp = exp2f(floorf(log2f(fmaxf(mag, 1e-20f))) - 7.f);
    if (!(d <= 2.f * ulp + 1e-6f)) atomicAdd(bad, 1ull);
    atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(d));
  }
}

struct Shape { const char* name; int N, K, epi; };

#ifdef DD_SG_EXPERIMENT
extern "C" int dd_stream_gemm_debug_ctas(int M, int N, int K, int epilogue);
// per-CTA %globaltimer stamps of the last launch (DD_SG_DBG=16): min / median / max per phase, us since the first entry
static void dump_timeline(const void* ws, size_t wsb, int G, const char* name, int M) {
  std::vector<long long> h((size_t)G * 8);
  CK(cudaMemcpy(h.data(), ws, (size_t)G * 64, cudaMemcpyDeviceToHost)); return;
  const char* names[5] = {"prod0_loop", "prod0_wait_empty", "mma_loop", "mma_wait_full", "epi_work"};
  printf("%s M=%d G=%d  cycles (min / median / max over CTAs)\n", name, M, G);
  for (int s = 0; s < 5; ++s) {
    std::vector<long long> v;
    for (int c = 0; c < G; ++c) v.push_back(h[c * 8 + s]);
    std::sort(v.begin(), v.end());
    printf("  %-18s %9lld %9lld %9lld\n", names[s], v.front(), v[v.size() / 2], v.back());
  }
  long long t0 = h[6];
  for (int c = 0; c < G; ++c) t0 = std::min(t0, h[c * 8 + 6]);
  std::vector<double> e, m, d;
  for (int c = 0; c < G; ++c) { e.push_back((h[c * 8 + 6] - t0) * 1e-3); m.push_back((h[c * 8 + 5] - t0) * 1e-3); d.push_back((h[c * 8 + 7] - t0) * 1e-3); }
  std::sort(e.begin(), e.end()); std::sort(m.begin(), m.end()); std::sort(d.begin(), d.end());
  printf("  us since first entry: entry %.1f/%.1f/%.1f  mma_issued %.1f/%.1f/%.1f  done %.1f/%.1f/%.1f  (mma_loop cycles / mma_issued us => %.2f GHz)\n",
         e.front(), e[G / 2], e.back(), m.front(), m[G / 2], m.back(), d.front(), d[G / 2], d.back(),
         (double)h[(G / 2) * 8 + 2] / ((h[(G / 2) * 8 + 5] - h[(G / 2) * 8 + 6]) * 1.0));
}
#endif

int main(int argc, char** argv) {
  std::vector<int> Ms;
  for (int i = 1; i < argc; ++i) Ms.push_back(atoi(argv[i]));
  if (Ms.empty()) Ms = {24, 64, 128, 192, 256};
  const Shape shapes[] = {{"qkv", 12288, 4096, 0}, {"o_proj", 4096, 4096, 1}, {"gate_up", 11008, 4096, 3},
                          {"down", 4096, 11008, 1}, {"lm_head", 32064, 4096, 2}};
  CK(cudaSetDevice(0));
  cublasHandle_t hb;
  CB(cublasCreate(&hb));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CB(cublasSetStream(hb, st));
  void* cublas_ws;
  CK(cudaMalloc(&cublas_ws, 64 << 20));
  CB(cublasSetWorkspace(hb, cublas_ws, 64 << 20));
  const double hbm_peak = 6536.4;  // MEASURED_PEAKS.json of this pool (GB/s)
  unsigned long long* d_bad;
  float* d_maxerr;
  CK(cudaMalloc(&d_bad, 8));
  CK(cudaMalloc(&d_maxerr, 4));
  for (const Shape& sh : shapes) {
    const int w_rows = sh.epi == 3 ? 2 * sh.N : sh.N;
    const size_t wbytes = (size_t)w_rows * sh.K * 2;
    const int nrot = (int)((600ull << 20) / wbytes) + 2;  // > 600 MB of distinct weights per rotation
    __nv_bfloat16* w;
    CK(cudaMalloc(&w, wbytes * nrot));
    fill_bf16<<<2048, 256, 0, st>>>(w, (size_t)w_rows * sh.K * nrot, 0x1234u + sh.N, 0.02f);
    for (int M : Ms) {
      __nv_bfloat16 *x, *res;
      void* y;
      float* c32;
      void* yb;
      CK(cudaMalloc(&x, (size_t)M * sh.K * 2));
      CK(cudaMalloc(&res, (size_t)M * sh.N * 2));
      CK(cudaMalloc(&y, (size_t)M * sh.N * 4));
      CK(cudaMalloc(&c32, (size_t)M * w_rows * 4));
      CK(cudaMalloc(&yb, (size_t)M * w_rows * 2));
      fill_bf16<<<256, 256, 0, st>>>(x, (size_t)M * sh.K, 77u + M, 1.0f);
      fill_bf16<<<256, 256, 0, st>>>(res, (size_t)M * sh.N, 99u + M, 1.0f);
      const size_t wsb = dd_stream_gemm_workspace_bytes(M, sh.N, sh.K, sh.epi);
      void* ws;
      CK(cudaMalloc(&ws, wsb));
      CK(cudaMemsetAsync(ws, 0, wsb, st));
      // ---- sanity: twice on the same weights (the second launch also proves the counters re-arm)
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemsetAsync(y, 0xff, (size_t)M * sh.N * (sh.epi == 2 ? 4 : 2), st));
        int rc = dd_stream_gemm(x, w, res, M, sh.N, sh.K, sh.epi, y, ws, wsb, st);
        if (rc) { fprintf(stderr, "dd_stream_gemm rc=%d: %s\n", rc, dd_last_error()); return 1; }
      }
      {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { fprintf(stderr, "dd_stream_gemm %s M=%d: %s\n", sh.name, M, cudaGetErrorString(e)); return 1; }
      }
#ifdef DD_SG_EXPERIMENT
      if (dd_stream_gemm(x, w + (size_t)(nrot - 1) * w_rows * sh.K, res, M, sh.N, sh.K, sh.epi, y, ws, wsb, st)) return 1;  // cold copy
      CK(cudaStreamSynchronize(st));
#endif
      const float one = 1.f, zero = 0.f;
      CB(cublasGemmEx(hb, CUBLAS_OP_T, CUBLAS_OP_N, w_rows, M, sh.K, &one, w, CUDA_R_16BF, sh.K, x, CUDA_R_16BF, sh.K,
                      &zero, c32, CUDA_R_32F, w_rows, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
      CK(cudaMemsetAsync(d_bad, 0, 8, st));
      CK(cudaMemsetAsync(d_maxerr, 0, 4, st));
      check<<<512, 256, 0, st>>>(c32, y, res, M, sh.N, sh.epi, w_rows, d_bad, d_maxerr);
      unsigned long long bad = 0;
      float maxerr = 0;
      CK(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(&maxerr, d_maxerr, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      // ---- timing: graph of nrot launches over rotating weights
      const int passes = (200 + nrot - 1) / nrot;   // >= 200 kernel nodes per graph launch (its own latency is ~tens of us)
      auto time_graph = [&](auto&& launch) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int pass = 0; pass < passes; ++pass)
          for (int r = 0; r < nrot; ++r) launch(w + (size_t)r * w_rows * sh.K);
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
        return best * 1e3f / (nrot * passes);  // us per launch
      };
#ifdef DD_SG_EXPERIMENT
      uint8_t* wsr;
      CK(cudaMalloc(&wsr, wsb * nrot));
      CK(cudaMemsetAsync(wsr, 0, wsb * nrot, st));
#endif
      const float us = time_graph([&](const __nv_bfloat16* wr) {
        void* wsl = ws;
#ifdef DD_SG_EXPERIMENT
        wsl = wsr + (size_t)((wr - w) / ((size_t)w_rows * sh.K)) * wsb;
#endif
        if (dd_stream_gemm(x, wr, res, M, sh.N, sh.K, sh.epi, y, wsl, wsb, st)) { fprintf(stderr, "%s\n", dd_last_error()); exit(1); }
      });
#ifdef DD_SG_EXPERIMENT
      if (getenv("DD_SG_DBG")) {   // consecutive launches of the last pass: per-phase stamps (us since the launch's first CTA entry)
        const int G = dd_stream_gemm_debug_ctas(M, sh.N, sh.K, sh.epi);
        std::vector<long long> h((size_t)G * 16);
        long long prev_entry = 0;
        printf("%s M=%d G=%d phases (median over CTAs, us since first entry): entry->entry | mma_issued | acc_complete | sync1 | xchg_written | sync2 | stored | exit\n", sh.name, M, G);
        for (int r = 0; r < nrot && r < 6; ++r) {
          CK(cudaMemcpy(h.data(), wsr + (size_t)r * wsb, (size_t)G * 128, cudaMemcpyDeviceToHost));
          long long e0 = h[6];
          for (int c = 0; c < G; ++c) e0 = std::min(e0, h[c * 16 + 6]);
          auto med = [&](int slot) { std::vector<long long> v; for (int c = 0; c < G; ++c) if (h[c * 16 + slot]) v.push_back(h[c * 16 + slot] - e0); if (v.empty()) return -1.0; std::sort(v.begin(), v.end()); return v[v.size() / 2] * 1e-3; };
          auto mx = [&](int slot) { long long m = 0; for (int c = 0; c < G; ++c) m = std::max(m, h[c * 16 + slot] - e0); return m * 1e-3; };
          printf("   [%5.1f | %5.1f | %5.1f | %5.1f | %5.1f | %5.1f | %5.1f | %5.1f (max %.1f)]\n", r ? (e0 - prev_entry) * 1e-3 : 0.0, med(5), med(8), med(9), med(10), med(11), med(12), med(7), mx(7));
          prev_entry = e0;
        }
      }
      CK(cudaFree(wsr));
#endif
#ifdef DD_SG_EXPERIMENT
      if (getenv("DD_SG_DBG")) dump_timeline(ws, wsb, dd_stream_gemm_debug_ctas(M, sh.N, sh.K, sh.epi), sh.name, M);
#endif
      const float us_cublas = time_graph([&](const __nv_bfloat16* wr) {
        CB(cublasGemmEx(hb, CUBLAS_OP_T, CUBLAS_OP_N, w_rows, M, sh.K, &one, wr, CUDA_R_16BF, sh.K, x, CUDA_R_16BF,
                        sh.K, &zero, yb, CUDA_R_16BF, w_rows, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT));
      });
      const double gbs = wbytes / (us * 1e-6) / 1e9;
      printf("{\"kernel\": \"stream_gemm\", \"shape\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"epi\": %d, \"us\": %.2f, "
             "\"GBps\": %.1f, \"frac_hbm\": %.3f, \"cublas_us\": %.2f, \"vs_cublas\": %.3f, \"bad\": %llu, \"max_abs_err\": %.4g, "
             "\"timing\": \"CUDA-graph replay over %d rotating weight copies\"}\n",
             sh.name, M, sh.N, sh.K, sh.epi, us, gbs, gbs / hbm_peak, us_cublas, us_cublas / us, bad, maxerr, nrot);
      fflush(stdout);
      CK(cudaFree(x)); CK(cudaFree(res)); CK(cudaFree(y)); CK(cudaFree(c32)); CK(cudaFree(yb)); CK(cudaFree(ws));
    }
    CK(cudaFree(w));
  }
  return 0;
}
