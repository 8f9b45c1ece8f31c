// Throughput / latency of the legacy warp-level tensor path (mma.sync -> HMMA) on sm_100a, to decide whether the 32-row training step
// can use it.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (KIND == 0)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 1)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 2)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 3)
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
  if (KIND == 4) {  // 32 FFMA per lane = the same 1024 MACs of a m16n8k8 tile on the CUDA cores
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c[0] = fmaf(__uint_as_float(a[i & 3]), __uint_as_float(b[i & 1]), c[0]);
      c[1] = fmaf(__uint_as_float(a[(i + 1) & 3]), __uint_as_float(b[i & 1]), c[1]);
      c[2] = fmaf(__uint_as_float(a[(i + 2) & 3]), __uint_as_float(b[(i + 1) & 1]), c[2]);
      c[3] = fmaf(__uint_as_float(a[(i + 3) & 3]), __uint_as_float(b[(i + 1) & 1]), c[3]);
    }
  }
}

template <int KIND, int CHAINS>
__global__ void bench(float* out, long long* cycles, int iters) {
  uint32_t a[4] = {threadIdx.x + 1u, threadIdx.x * 3u, 7u, 11u}, b[2] = {threadIdx.x ^ 5u, 13u};
  float c[CHAINS][4];
  for (int j = 0; j < CHAINS; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma<KIND>(c[j], a, b);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int CHAINS>
void run(const char* name, int threads, double macs) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  bench<KIND, CHAINS><<<148, threads>>>(out, cyc, 10);
  bench<KIND, CHAINS><<<148, threads>>>(out, cyc, iters);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per_warp = avg / (double(iters) * CHAINS);
  const int warps = threads / 32;
  printf("%-28s warps/SM %2d chains %d : %7.1f cycles per instr per warp, %6.2f instr/clk/SM, %7.1f MAC/clk/SM\n", name, warps, CHAINS, per_warp,
         warps / per_warp, macs * warps / per_warp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 1>("tf32 m16n8k8 (latency)", 32, 1024);
  run<0, 4>("tf32 m16n8k8", 128, 1024);
  run<0, 4>("tf32 m16n8k8", 512, 1024);
  run<3, 4>("tf32 m16n8k4", 512, 512);
  run<1, 1>("bf16 m16n8k16 (latency)", 32, 2048);
  run<1, 4>("bf16 m16n8k16", 128, 2048);
  run<1, 4>("bf16 m16n8k16", 512, 2048);
  run<2, 4>("f16 m16n8k16", 512, 2048);
  run<4, 1>("32 FFMA (latency)", 32, 1024);
  run<4, 4>("32 FFMA", 512, 1024);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
