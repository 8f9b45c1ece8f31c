// Micro-benchmark: how fast can ONE thread issue small tcgen05.mma (M=128, N, K=8 tf32 / K=16 bf16, A from TMEM)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mma_issue_bench scratch/mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_bdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int fmt, int n) { return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int MODE>
__global__ void bench(int n_mma, int N, long long* out, int n_threads_issuing) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (warp < n_threads_issuing && lane == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t id32 = make_idesc(2, N), id16 = make_idesc(1, N);
    const uint64_t d0 = make_bdesc(sb, N * 16, 128);
    const uint32_t dcol = tm + warp * 128;          // each issuer its own accumulator + A region
    const uint32_t acol = tm + warp * 128 + 64;
    const long long t0 = clock64();
    if (MODE == 0) {  // runtime loop, tf32
      for (int i = 0; i < n_mma; ++i) mma_tf32_ts(dcol, acol + (i & 7) * 8, d0 + (uint64_t)((i & 7) * 2 * N), id32, i > 0);
    } else if (MODE == 1) {  // unrolled x8, tf32
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_tf32_ts(dcol, acol + j * 8, d0 + (uint64_t)(j * 2 * N), id32, (i + j) > 0);
      }
    } else {  // unrolled x8, bf16 kind
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_f16_ts(dcol, acol + j * 8, d0 + (uint64_t)(j * 2 * N), id16, (i + j) > 0);
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int MODE>
void run(const char* name, int n_mma, int N, int issuers) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  bench<MODE><<<1, 128, 64 * 1024>>>(n_mma, N, d, issuers);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d n=%4d issuers=%d : issue %6.1f cyc/mma, issue+complete %6.1f cyc/mma  (%s)\n", name, N, n_mma, issuers, (double)h[0] / n_mma,
         (double)h[1] / n_mma, cudaGetErrorString(e));
  if (issuers > 1) printf("%-28s   second issuer: issue %6.1f, complete %6.1f\n", "", (double)h[2] / n_mma, (double)h[3] / n_mma);
  cudaFree(d);
}

int main() {
  for (int N : {32, 64, 128, 256}) {
    run<0>("tf32 runtime loop", 256, N, 1);
    run<1>("tf32 unrolled x8", 256, N, 1);
    run<2>("bf16 unrolled x8", 256, N, 1);
  }
  run<1>("tf32 unrolled, 2 issuers", 256, 64, 2);
  run<1>("tf32 unrolled, 4 issuers", 256, 64, 4);
  run<1>("tf32 unrolled short", 16, 64, 1);
  run<1>("tf32 unrolled short", 24, 64, 1);
  return 0;
}
