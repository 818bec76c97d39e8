// Device peak probes used by bench.py for the roofline denominators that MEASURED_PEAKS.json does not hold:
// sfem_measure_fp64_peak = sustained DFMA rate of this GPU (the "fp64 roofline" of BASELINE.json's north_star).
#include <algorithm>

#include "sfem_internal.h"

namespace sfem {
namespace {

// 8 independent accumulator chains per thread, 16 warps per CTA: enough to keep the fp64 pipe full at its issue rate
__global__ void __launch_bounds__(512) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace
}  // namespace sfem

using namespace sfem;

extern "C" int sfem_measure_fp64_peak(double* tflops, void* stream) {
  SFEM_CHECK(tflops, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  SFEM_CUDA(cudaGetDevice(&dev));
  SFEM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  SFEM_CUDA(dev_malloc(&d, 8, st));
  cudaEvent_t e0, e1;
  SFEM_CUDA(cudaEventCreate(&e0));
  SFEM_CUDA(cudaEventCreate(&e1));
  const int iters = 2048, grid = sms * 4;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    SFEM_CUDA(cudaEventRecord(e0, st));
    k_dfma_peak<<<grid, 512, 0, st>>>(d, iters, 0.999999, 1e-9);
    SFEM_CUDA(cudaEventRecord(e1, st));
    SFEM_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SFEM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * 512.0 * grid;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  dev_free(d, st);
  *tflops = best;
  return 0;
}
