// Measures the vector-fp64 FMA ceiling and an HBM read+write copy ceiling of the
// current GPU (SURVEY.md section 7 hard part 2: no fp64 peak is in MEASURED_PEAKS.json).
// Prints one JSON line.  Usage: fp64_peak [device]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                   \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = double(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ in, double2* out, size_t n) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    out[i] = in[i];
}

int main(int argc, char** argv) {
  const int dev = argc > 1 ? atoi(argv[1]) : 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = p.multiProcessorCount * 8, iters = 4096;
  constexpr int ILP = 8;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * size_t(blocks) * 256));
  double best_tf = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    k_dfma<ILP><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * double(blocks) * 256.0 * double(iters) * ILP;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best_tf) best_tf = tf;
  }
  const size_t n = size_t(1) << 27;  // 2 GiB per buffer of double2
  double2 *a, *b;
  CK(cudaMalloc(&a, n * sizeof(double2)));
  CK(cudaMalloc(&b, n * sizeof(double2)));
  CK(cudaMemset(a, 1, n * sizeof(double2)));
  double best_gbs = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    k_copy<<<p.multiProcessorCount * 16, 256>>>(a, b, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double gbs = 2.0 * double(n) * sizeof(double2) / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbs > best_gbs) best_gbs = gbs;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.3f, \"copy_gbs\": %.1f}\n", p.name,
         p.multiProcessorCount, best_tf, best_gbs);
  return 0;
}
