// Latency microbenchmarks for the building blocks of the search kernel (one warp unless noted):
// dependent-chain cycles per op.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include "../automatedvaletparking_b200/csrc/avp_dev.cuh"

#define N 512
template <class F> __device__ long long chain(F f, double &x) {
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) x = f(x);
  long long t1 = clock64();
  return (t1 - t0) / N;
}

__global__ void k_lat(long long *out, double seed, const double *g, double *sink) {
  double x = seed + threadIdx.x * 1e-3;
  int k = 0;
  out[k++] = chain([](double v) { return __fma_rn(v, 1.0000001, 1e-9); }, x);                 // 0 DFMA
  out[k++] = chain([](double v) { return __dadd_rn(v, 1e-9); }, x);                           // 1 DADD
  out[k++] = chain([](double v) { return __dmul_rn(v, 1.0000001); }, x);                      // 2 DMUL
  out[k++] = chain([](double v) { return 1.0 / (v + 1.5); }, x);                              // 3 DDIV (+add)
  out[k++] = chain([](double v) { return sqrt(v + 2.0); }, x);                                // 4 DSQRT (+add)
  out[k++] = chain([](double v) { return d_sin(v) + 0.7; }, x);                               // 5 d_sin
  out[k++] = chain([](double v) { return d_cos(v) + 0.3; }, x);                               // 6 d_cos
  out[k++] = chain([](double v) { return d_atan2(v, 0.7) + 0.4; }, x);                        // 7 d_atan2
  out[k++] = chain([](double v) { return py_hypot(v, 0.9) - 0.3; }, x);                       // 8 py_hypot
  out[k++] = chain([](double v) { return d_pow2(v) * 0.5 + 0.4; }, x);                        // 9 d_pow2
  out[k++] = chain([](double v) { return d_asin(v * 0.3) + 0.5; }, x);                        // 10 d_asin
  out[k++] = chain([](double v) { return d_acos(v * 0.3) * 0.5; }, x);                        // 11 d_acos
  out[k++] = chain([](double v) { return d_tan(v * 0.5) * 0.3 + 0.2; }, x);                   // 12 d_tan
  out[k++] = chain([](double v) { return rs_M(v * 3.0 + 1.0); }, x);                          // 13 rs_M
  out[k++] = chain([](double v) { return floor(v * 7.3) * 0.01 + 0.5; }, x);                  // 14 floor chain
  // inlined sin (no call overhead)
  out[k++] = chain([](double v) { return avp_sin(v) + 0.7; }, x);                             // 15 avp_sin inlined
  // global load dependent chain (L1 hits after the first pass): pointer chase over 32 doubles
  {
    int idx = threadIdx.x & 31; long long t0 = clock64();
    for (int i = 0; i < N; ++i) idx = (int)g[idx];
    out[k++] = (clock64() - t0) / N; x += idx;                                                // 16 LDG L1-hit chase
  }
  {
    __shared__ double sm[32]; sm[threadIdx.x & 31] = (double)((threadIdx.x + 1) & 31); __syncwarp();
    int idx = threadIdx.x & 31; long long t0 = clock64();
    for (int i = 0; i < N; ++i) idx = (int)sm[idx];
    out[k++] = (clock64() - t0) / N; x += idx;                                                // 17 LDS chase
  }
  {
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { x = shfl_dbl(x, (threadIdx.x + 1) & 31); }
    out[k++] = (clock64() - t0) / N;                                                          // 18 64-bit shuffle
  }
  sink[threadIdx.x] = x;
}

// barrier latency: BLOCK threads all arriving together
__global__ void k_bar(long long *out) {
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / 256;
  if (threadIdx.x >= 32) {
    long long t2 = clock64();
    for (int i = 0; i < 256; ++i) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 32) : "memory");
    if (threadIdx.x == 32) out[1] = (clock64() - t2) / 256;
  }
}

// divergence cost: lanes take 1, 2 or 4 different transcendental paths
__global__ void k_div(long long *out, double seed, double *sink) {
  double x = seed + threadIdx.x * 1e-3;
  const int l = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < 128; ++i) { if (l & 1) x = d_sin(x) + 0.7; else x = d_atan2(x, 0.7) + 0.4; }
  out[0] = (clock64() - t0) / 128;
  t0 = clock64();
  for (int i = 0; i < 128; ++i) { switch (l & 3) { case 0: x = d_sin(x) + 0.7; break; case 1: x = d_atan2(x, 0.7) + 0.4; break; case 2: x = py_hypot(x, 0.9) - 0.3; break; default: x = d_acos(x * 0.3) * 0.5; } }
  out[1] = (clock64() - t0) / 128;
  sink[threadIdx.x] = x;
}

int main() {
  long long *d, h[32]; double *g, *sink; double hg[32];
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&g, sizeof(hg)); cudaMalloc(&sink, 8 * 1024);
  for (int i = 0; i < 32; ++i) hg[i] = (double)((i * 7 + 3) & 31);
  cudaMemcpy(g, hg, sizeof(hg), cudaMemcpyHostToDevice);
  const char *names[] = {"DFMA", "DADD", "DMUL", "DDIV+add", "DSQRT+add", "d_sin(call)", "d_cos(call)", "d_atan2", "py_hypot", "d_pow2", "d_asin", "d_acos", "d_tan", "rs_M", "floor chain", "avp_sin inlined", "LDG chase (L1)", "LDS chase", "shfl64"};
  for (int threads : {32, 1}) {
    k_lat<<<1, threads>>>(d, 0.4, g, sink); cudaDeviceSynchronize();
    k_lat<<<1, threads>>>(d, 0.4, g, sink); cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("--- %d thread(s), dependent chain, cycles per op (2nd run: warm i-cache/tables)\n", threads);
    for (int i = 0; i < 19; ++i) printf("%-18s %6lld\n", names[i], h[i]);
  }
  for (int b : {128, 256, 512}) { k_bar<<<1, b>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); printf("block %d: __syncthreads %lld cycles, bar.sync 1 (block-32 threads) %lld cycles\n", b, h[0], h[1]); }
  k_div<<<1, 32>>>(d, 0.4, sink); cudaDeviceSynchronize(); k_div<<<1, 32>>>(d, 0.4, sink); cudaDeviceSynchronize();
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("divergent 2 paths (sin|atan2): %lld cycles/iter; 4 paths (sin|atan2|hypot|acos): %lld cycles/iter\n", h[0], h[1]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
