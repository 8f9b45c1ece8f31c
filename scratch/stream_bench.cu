// What does a pure streaming kernel with this path's traffic mix (2 arrays read, 7 written, + 3 row vectors) reach on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(const float4* __restrict__ x, const float4* __restrict__ y, float4* o0, float4* o1, float4* o2, float4* o3, long n4, int nout) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 a = x[i], b = y[i];
    float4 d = make_float4(fabsf(a.x - b.x), fabsf(a.y - b.y), fabsf(a.z - b.z), fabsf(a.w - b.w));
    o0[i] = a;
    if (nout > 1) o1[i] = d;
    if (nout > 2) o2[i] = make_float4(d.x * 2, d.y * 2, d.z * 2, d.w * 2);
    if (nout > 3) o3[i] = make_float4(d.x * 3, d.y * 3, d.z * 3, d.w * 3);
  }
}
__global__ void copyk(const float4* __restrict__ x, float4* o, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) o[i] = x[i];
}
__global__ void fillk(float4* o, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) o[i] = make_float4(1, 2, 3, 4);
}
int main() {
  const long rows = 10000000, n4 = rows * 16;
  float4 *x, *y, *o[4];
  cudaMalloc(&x, n4 * 16); cudaMalloc(&y, n4 * 16);
  for (int k = 0; k < 4; ++k) cudaMalloc(&o[k], n4 * 16);
  cudaMemset(x, 0, n4 * 16); cudaMemset(y, 0, n4 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto f, const char* name, double bytes) {
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-28s %.3f ms  %.0f GB/s\n", name, ms, bytes / ms / 1e6);
  };
  const int g = 148 * 8, b = 512;
  double B = (double)n4 * 16;
  time([&] { copyk<<<g, b>>>(x, o[0], n4); }, "copy 1r:1w", 2 * B);
  time([&] { fillk<<<g, b>>>(o[0], n4); }, "fill 0r:1w", B);
  time([&] { mix<<<g, b>>>(x, y, o[0], o[1], o[2], o[3], n4, 1); }, "mix 2r:1w", 3 * B);
  time([&] { mix<<<g, b>>>(x, y, o[0], o[1], o[2], o[3], n4, 2); }, "mix 2r:2w", 4 * B);
  time([&] { mix<<<g, b>>>(x, y, o[0], o[1], o[2], o[3], n4, 4); }, "mix 2r:4w", 6 * B);
  return 0;
}
