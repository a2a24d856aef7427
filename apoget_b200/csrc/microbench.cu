// Pipe-throughput micro-benchmarks: the roofline denominators that MEASURED_PEAKS.json does not
// hold (SURVEY.md 8d asks for a measured __popc rate on sm_100 instead of the cc<=9.0 table value).
#include <cuda_runtime.h>
#include <stdint.h>

#include "apt_b200.h"

namespace {

template <int WHAT>
__global__ void __launch_bounds__(256) pipe_kernel (uint32_t iters, uint32_t seed, uint32_t *sink) {
  // 8 independent dependency chains per thread keep the pipe's issue slots full
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 2u, a3 = a0 * 7u + 3u;
  uint32_t a4 = a0 * 11u + 4u, a5 = a0 * 13u + 5u, a6 = a0 * 17u + 6u, a7 = a0 * 19u + 7u;
  double d0 = a0, d1 = a1, d2 = a2, d3 = a3;
  float f0 = a0, f1 = a1, f2 = a2, f3 = a3, f4 = a4, f5 = a5, f6 = a6, f7 = a7;
  for (uint32_t i = 0; i < iters; i++) {
    if (WHAT == 0) {          // POPC (+ the XOR that feeds it, as in a Hamming distance)
      a0 = __popc(a0 ^ seed) + i; a1 = __popc(a1 ^ seed) + i; a2 = __popc(a2 ^ seed) + i; a3 = __popc(a3 ^ seed) + i;
      a4 = __popc(a4 ^ seed) + i; a5 = __popc(a5 ^ seed) + i; a6 = __popc(a6 ^ seed) + i; a7 = __popc(a7 ^ seed) + i;
    } else if (WHAT == 1) {   // LOP3 only (one three-input instruction per operation: a ^ b ^ c)
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(a1), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(a2), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(a3), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(a4), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a4) : "r"(a5), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a5) : "r"(a6), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a6) : "r"(a7), "r"(i));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a7) : "r"(a0), "r"(i));
    } else if (WHAT == 2) {   // DFMA
      d0 = fma(d0, 1.0000001, 0.5); d1 = fma(d1, 1.0000001, 0.5); d2 = fma(d2, 1.0000001, 0.5); d3 = fma(d3, 1.0000001, 0.5);
    } else if (WHAT == 4) {   // F2F: float -> double -> float round trips (the `float avg += double c` of the scoring loop)
      f0 = float(double(f0) + 0.1); f1 = float(double(f1) + 0.1); f2 = float(double(f2) + 0.1); f3 = float(double(f3) + 0.1);   // 0.1 is no float: the sum cannot be folded into an FADD
    } else {                  // FFMA
      f0 = fmaf(f0, 1.0001f, f1); f1 = fmaf(f1, 1.0001f, f2); f2 = fmaf(f2, 1.0001f, f3); f3 = fmaf(f3, 1.0001f, f4);
      f4 = fmaf(f4, 1.0001f, f5); f5 = fmaf(f5, 1.0001f, f6); f6 = fmaf(f6, 1.0001f, f7); f7 = fmaf(f7, 1.0001f, f0);
    }
  }
  uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ uint32_t(d0 + d1 + d2 + d3) ^ uint32_t(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
  if (r == 0x12345678u) sink[0] = r;      // practically never: keeps the loop alive
}

template <int WHAT>
double run (uint32_t opsPerIter) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t *sink = nullptr;
  cudaMalloc(&sink, 4);
  const uint32_t iters = 1u << 16, grid = uint32_t(sms) * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);  cudaEventCreate(&e1);
  pipe_kernel<WHAT><<<grid, block>>>(1024, 1, sink);        // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    pipe_kernel<WHAT><<<grid, block>>>(iters, 1 + rep, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = double(grid) * block * iters * opsPerIter;
    best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
  }
  cudaEventDestroy(e0);  cudaEventDestroy(e1);  cudaFree(sink);
  return cudaGetLastError() == cudaSuccess ? best : -1.;
}

} // namespace

extern "C" int apt_microbench (int device, double out[4]) {
  if (cudaSetDevice(device) != cudaSuccess) return APT_ERR_CUDA;
  out[0] = run<0>(8);      // popc32 / s (each with its XOR)
  out[1] = run<1>(8);      // LOP3 / s
  out[2] = run<2>(4);      // DFMA / s
  out[3] = run<3>(8);      // FFMA / s
  return (out[0] > 0 && out[1] > 0 && out[2] > 0 && out[3] > 0) ? APT_OK : APT_ERR_CUDA;
}

extern "C" int apt_microbench_cvt (int device, double out[1]) {
  if (cudaSetDevice(device) != cudaSuccess) return APT_ERR_CUDA;
  out[0] = run<4>(8);      // F2F conversions / s (4 chains x {f32->f64, f64->f32}; the DADD between them is on another pipe)
  return out[0] > 0 ? APT_OK : APT_ERR_CUDA;
}
