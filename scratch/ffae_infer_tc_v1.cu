// K1+K4, variant 2: fused Dense-stack forward + anomaly score on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// Covers 64-tag autoencoders with hidden widths <= 64 (feedforward_hourglass(64) = 64-53-43-32-32-43-53-64 is the
// BASELINE workload).  The path is HBM-bound (1 548 algorithmic bytes and 30 236 FLOP per window => 128 TFLOP/s at
// the measured 6.58 TB/s): fp32 CUDA cores (74 TFLOP/s peak) cannot keep up, the tensor cores can.
//
// Numerics: 1e-4 parity with the float32 reference forbids plain TF32 (2^-11 per operand).  Every layer is computed as
//        D  =  A_lo*W_hi  +  A_hi*W_hi            (kind::tf32, A = A_hi + A_lo exactly, W_hi = W rounded to TF32)
//           +  bf16(A)*bf16(W - W_hi)             (kind::f16, the 2^-11-sized correction needs only 8 bits)
// accumulated in fp32 in TMEM: error ~2^-20 relative per product, while the bf16 correction image costs half the
// shared memory of a TF32 one (the budget that lets x/y/out staging fit beside the weights).
//
// One persistent CTA per SM: 8 epilogue warps + 1 control warp.  Per work item (job chunk) the slot's weights are
// split and laid out once in shared memory as UMMA K-major operands ([K/4][N][4] TF32, [K/8][N][8] BF16).  Per
// 128-row tile:  TMA (SWIZZLE_128B boxes) brings x and y into shared memory; the epilogue warps split x into the A
// operand held in TMEM (lane = row); for every layer the control thread issues tcgen05.mma (A from TMEM, B from the
// resident weight image, D in TMEM) and commits to an mbarrier; the epilogue warps tcgen05.ld the accumulator, add
// bias, apply tanh, split and tcgen05.st the next layer's A operand.  The last layer's epilogue forms every anomaly
// column against the y tile and leaves through swizzled shared-memory boxes + TMA tensor stores (full-line writes).
//
// Reference arithmetic replaced: keras Dense under Model.predict (gordo/machine/model/models.py:289-300) and
// DiffBasedAnomalyDetector.anomaly (gordo/machine/model/anomaly/diff.py:350-385, 420-444).
#include <cuda.h>
#include <cuda_bf16.h>
#include "gb_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr int NTHREADS = 288;  // warps 0-7: epilogue (warp%4 = TMEM lane quadrant, warp/4 = column half); warp 8: control
constexpr int EPI_THREADS = 256;
constexpr int MAXL = 8;
constexpr int BOX_BYTES = TILE * 128;  // 128 rows x 32 fp32, one SWIZZLE_128B box
constexpr int W = 64;                  // feature width this kernel is specialised for

// TMEM column map (fp32 columns)
constexpr uint32_t COL_D = 0, COL_AHI = 64, COL_ALO = 128, COL_ABF = 192, TMEM_COLS = 256;

struct TcArgs {
  int n_layers, last_layer;  // layers actually evaluated: 0..last_layer (debug aid; == n_layers-1 in production)
  int K[MAXL], N[MAXL], Np[MAXL], k8[MAXL], k16[MAXL], act[MAXL];
  int whi_ofs[MAXL], wlo_ofs[MAXL], bias_ofs[MAXL];  // byte offsets into dynamic smem
  int pofs[MAXL];                                    // float offsets of W_l in the canonical parameter vector
  int w_bytes;                                       // bytes of the weight+bias region (zero-filled before staging)
  int vec_ofs, xbox_ofs, ybox_ofs, stage_ofs, pair_ofs, bar_ofs;
  int n_jobs, chunks_per_job, rows_per_chunk, flags;
  long pstride;
  const float* params;
  const gb_job* jobs;
  const float *y, *scale, *feat_thr, *agg_thr;
  float *o_model, *o_ts, *o_tu, *o_conf, *o_tots, *o_totu, *o_totconf;
};

enum { FLAG_SWAP_BF16 = 1 };

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, no-swizzle UMMA shared-memory descriptor: core matrix = 8 rows x 16 B contiguous;
// SBO = byte distance between 8-row groups (along N), LBO = byte distance between 16-byte K chunks.
__device__ __forceinline__ uint64_t make_bdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type (bits 61-63) = 0: SWIZZLE_NONE
}
// instruction descriptor: D fp32, A/B format fmt (2 = TF32, 1 = BF16), both K-major, M = 128, N = n
__host__ __device__ __forceinline__ uint32_t make_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  // the wait is part of the same statement so no consumer of r0..r7 can be scheduled ahead of it
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
      : "r"(taddr)
      : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

// tanh(x) = 1 - 2/(1 + 2^(2x*log2 e)); absolute error ~2e-7 (ex2.approx / rcp.approx are ~1-2 ulp), exact limits at +-inf
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
__device__ __forceinline__ float act_fast(int act, float z) { return act == GB_ACT_TANH ? tanh_fast(z) : gb::apply_act(act, z); }

// split 8 activations into the three A operands and store them at column offset `col` of this thread's TMEM lane
__device__ __forceinline__ void store_a_operands(uint32_t lane_base, int col, const float (&a)[8], bool swap_bf16) {
  uint32_t hi[8], lo[8], bf[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t h = __float_as_uint(a[i]) & 0xffffe000u;
    hi[i] = h;
    lo[i] = __float_as_uint(a[i] - __uint_as_float(h));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 p = swap_bf16 ? __floats2bfloat162_rn(a[2 * i + 1], a[2 * i]) : __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    bf[i] = *reinterpret_cast<const uint32_t*>(&p);
  }
  tmem_st8(lane_base + COL_AHI + col, hi);
  tmem_st8(lane_base + COL_ALO + col, lo);
  tmem_st4(lane_base + COL_ABF + (col >> 1), bf);
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NTHREADS, 1)
ffae_tc_kernel(const __grid_constant__ TcArgs a, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
               const __grid_constant__ CUtensorMap map_model, const __grid_constant__ CUtensorMap map_ts,
               const __grid_constant__ CUtensorMap map_tu, const __grid_constant__ CUtensorMap map_conf) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_ctrl = warp == 8;
  const int q = warp & 3, h = (warp >> 2) & 1;  // epilogue: lane quadrant / column half
  const int row = q * 32 + lane;                // tile row owned by this epilogue thread
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_x_full = sbase + a.bar_ofs, bar_y_full = bar_x_full + 8, bar_a_ready = bar_x_full + 16, bar_d_ready = bar_x_full + 24,
                 bar_y_free = bar_x_full + 32;
  const bool has_y = a.y != nullptr;
  const int L = a.last_layer + 1;
  const bool swap_bf16 = (a.flags & FLAG_SWAP_BF16) != 0;

  if (tid == 0) {
    mbar_init(bar_x_full, 1);
    mbar_init(bar_y_full, 1);
    mbar_init(bar_a_ready, EPI_THREADS);
    mbar_init(bar_d_ready, 1);
    mbar_init(bar_y_free, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (is_ctrl) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem_base;
  const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);

  uint32_t ph_x = 0, ph_y = 0, ph_a = 0, ph_d = 0, ph_yf = 0;
  int cur_slot = -1;
  const int n_items = a.n_jobs * a.chunks_per_job;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int job_id = item / a.chunks_per_job, chunk = item - job_id * a.chunks_per_job;
    const gb_job job = a.jobs[job_id];
    const int row_begin = chunk * a.rows_per_chunk;
    if (row_begin >= job.n_rows) continue;  // uniform across the CTA
    const int row_end = min(job.n_rows, row_begin + a.rows_per_chunk);
    const int n_tiles = (row_end - row_begin + TILE - 1) / TILE;

    // ---- stage this slot's weights: split to TF32-hi / BF16-lo and lay out as UMMA K-major operands ------------
    if (job.slot != cur_slot) {
      cur_slot = job.slot;
      const float* P = a.params + (long)job.slot * a.pstride;
      for (int i = tid; i < a.w_bytes / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      for (int l = 0; l < L; ++l) {
        const int K = a.K[l], N = a.N[l], Np = a.Np[l];
        const float* Wg = P + a.pofs[l];
        float* whi = reinterpret_cast<float*>(smem + a.whi_ofs[l]);
        __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>(smem + a.wlo_ofs[l]);
        for (int idx = tid; idx < K * N; idx += NTHREADS) {
          const int k = idx / N, n = idx - k * N;
          const float w = __ldg(Wg + idx);
          const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xffffe000u);  // round to nearest TF32
          whi[((k >> 2) * Np + n) * 4 + (k & 3)] = hi;
          wlo[((k >> 3) * Np + n) * 8 + (k & 7)] = __float2bfloat16_rn(w - hi);
        }
        float* bl = reinterpret_cast<float*>(smem + a.bias_ofs[l]);
        for (int n = tid; n < N; n += NTHREADS) bl[n] = __ldg(Wg + K * N + n);
      }
      float* vec = reinterpret_cast<float*>(smem + a.vec_ofs);  // [0,64): scale, [64,128): 1/feat_thr
      for (int j = tid; j < W; j += NTHREADS) {
        vec[j] = a.scale ? __ldg(a.scale + (long)job.slot * W + j) : 0.f;
        vec[W + j] = a.feat_thr ? 1.0f / __ldg(a.feat_thr + (long)job.slot * W + j) : 0.f;
      }
      fence_proxy_async();  // generic-proxy writes above are read by the tensor core (async proxy)
    }
    __syncthreads();

    if (is_ctrl) {
      // =========================================== control warp: TMA producer + MMA issuer (one elected lane)
      if (lane == 0) {
        const int xrow0 = (int)(job.x_row + row_begin);
        mbar_expect_tx(bar_x_full, 2 * BOX_BYTES);
        tma_load_2d(sbase + a.xbox_ofs, &map_x, 0, xrow0, bar_x_full);
        tma_load_2d(sbase + a.xbox_ofs + BOX_BYTES, &map_x, 32, xrow0, bar_x_full);
        if (has_y) {
          mbar_expect_tx(bar_y_full, 2 * BOX_BYTES);
          tma_load_2d(sbase + a.ybox_ofs, &map_y, 0, xrow0, bar_y_full);
          tma_load_2d(sbase + a.ybox_ofs + BOX_BYTES, &map_y, 32, xrow0, bar_y_full);
        }
        for (int t = 0; t < n_tiles; ++t) {
          const bool more = t + 1 < n_tiles;
          const int next_row = (int)(job.x_row + row_begin + (t + 1) * TILE);
          for (int l = 0; l < L; ++l) {
            mbar_wait(bar_a_ready, ph_a);
            ph_a ^= 1;
            tc_fence_after();
            if (l == 0 && more) {  // A0 is in TMEM => the x boxes are free again
              mbar_expect_tx(bar_x_full, 2 * BOX_BYTES);
              tma_load_2d(sbase + a.xbox_ofs, &map_x, 0, next_row, bar_x_full);
              tma_load_2d(sbase + a.xbox_ofs + BOX_BYTES, &map_x, 32, next_row, bar_x_full);
            }
            const int Np = a.Np[l];
            const uint32_t step = 2u * (uint32_t)Np * 16u;  // bytes between consecutive K-steps (two 16-byte chunks)
            const uint32_t id32 = make_idesc(2, Np), id16 = make_idesc(1, Np);
            const uint32_t whi = sbase + a.whi_ofs[l], wlo = sbase + a.wlo_ofs[l];
            uint32_t acc = 0;
            for (int ks = 0; ks < a.k8[l]; ++ks) {  // A_lo * W_hi
              mma_tf32_ts(tmem + COL_D, tmem + COL_ALO + ks * 8, make_bdesc(whi + ks * step, Np * 16, 128), id32, acc);
              acc = 1;
            }
            for (int ks = 0; ks < a.k8[l]; ++ks)  // A_hi * W_hi
              mma_tf32_ts(tmem + COL_D, tmem + COL_AHI + ks * 8, make_bdesc(whi + ks * step, Np * 16, 128), id32, 1);
            for (int ks = 0; ks < a.k16[l]; ++ks)  // bf16(A) * bf16(W_lo)
              mma_bf16_ts(tmem + COL_D, tmem + COL_ABF + ks * 8, make_bdesc(wlo + ks * step, Np * 16, 128), id16, 1);
            mma_commit(bar_d_ready);
          }
          if (has_y) {
            if (more) {
              mbar_wait(bar_y_free, ph_yf);
              mbar_expect_tx(bar_y_full, 2 * BOX_BYTES);
              tma_load_2d(sbase + a.ybox_ofs, &map_y, 0, next_row, bar_y_full);
              tma_load_2d(sbase + a.ybox_ofs + BOX_BYTES, &map_y, 32, next_row, bar_y_full);
            }
            ph_yf ^= 1;
          }
        }
      }
      __syncwarp();
    } else {
      // =========================================== epilogue warps
      const float* vec = reinterpret_cast<const float*>(smem + a.vec_ofs);
      const uint32_t xbox = sbase + a.xbox_ofs + h * BOX_BYTES, ybox = sbase + a.ybox_ofs + h * BOX_BYTES;
      const uint32_t stage = sbase + a.stage_ofs + h * BOX_BYTES;
      float* pair = reinterpret_cast<float*>(smem + a.pair_ofs);  // [2][TILE] row-sum exchange between the column halves
      const uint32_t swz_row = (uint32_t)row * 128u;
      const bool issuer = (q == 0 && lane == 0);
      const float inv_w = 1.0f / (float)W;

      for (int t = 0; t < n_tiles; ++t) {
        const int trow = row_begin + t * TILE;
        const int nrows = min(TILE, row_end - trow);
        const long grow0 = job.out_row + trow;
        const bool full = nrows == TILE;

        // ---- x -> A operand of layer 0 ----------------------------------------------------------------------
        mbar_wait(bar_x_full, ph_x);
        ph_x ^= 1;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          float v[8];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const uint32_t addr = xbox + swz_row + ((uint32_t)((c + cc) ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4 * cc]), "=f"(v[4 * cc + 1]), "=f"(v[4 * cc + 2]), "=f"(v[4 * cc + 3]) : "r"(addr));
          }
          store_a_operands(lane_base, h * 32 + c * 4, v, swap_bf16);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_a_ready);

        // ---- hidden layers: D -> bias, activation -> next A operand ---------------------------------------------
        for (int l = 0; l + 1 < L; ++l) {
          const int half = a.Np[l] >> 1;  // columns this warp owns: [h*half, (h+1)*half)
          const int act = a.act[l];
          const float* bl = reinterpret_cast<const float*>(smem + a.bias_ofs[l]) + h * half;
          mbar_wait(bar_d_ready, ph_d);
          ph_d ^= 1;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 8 < half) {
              float v[8];
              tmem_ld8(lane_base + COL_D + h * half + c * 8, v);
              const float4 b0 = *reinterpret_cast<const float4*>(bl + c * 8), b1 = *reinterpret_cast<const float4*>(bl + c * 8 + 4);
              v[0] = act_fast(act, v[0] + b0.x); v[1] = act_fast(act, v[1] + b0.y); v[2] = act_fast(act, v[2] + b0.z); v[3] = act_fast(act, v[3] + b0.w);
              v[4] = act_fast(act, v[4] + b1.x); v[5] = act_fast(act, v[5] + b1.y); v[6] = act_fast(act, v[6] + b1.z); v[7] = act_fast(act, v[7] + b1.w);
              if (act == GB_ACT_SIGMOID) {  // padded columns must stay exactly zero (sigmoid(0) != 0)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (h * half + c * 8 + i >= a.N[l]) v[i] = 0.f;
              }
              store_a_operands(lane_base, h * half + c * 8, v, swap_bf16);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(bar_a_ready);
        }

        // ---- last layer: model output + anomaly columns ------------------------------------------------------------
        {
          const int l = L - 1;
          const int act = (l == a.n_layers - 1) ? a.act[l] : GB_ACT_LINEAR;
          const float* bl = reinterpret_cast<const float*>(smem + a.bias_ofs[l]) + h * 32;
          float yh[32];
          mbar_wait(bar_d_ready, ph_d);
          ph_d ^= 1;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v[8];
            tmem_ld8(lane_base + COL_D + h * 32 + c * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) yh[c * 8 + i] = act_fast(act, v[i] + bl[c * 8 + i]);
          }
          tc_fence_before();
          float yt[32];
          if (has_y) {
            mbar_wait(bar_y_full, ph_y);
            ph_y ^= 1;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t addr = ybox + swz_row + ((uint32_t)(c ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(yt[4 * c]), "=f"(yt[4 * c + 1]), "=f"(yt[4 * c + 2]), "=f"(yt[4 * c + 3]) : "r"(addr));
            }
            mbar_arrive(bar_y_free);
          }

          // one output array at a time: registers -> swizzled staging box -> TMA tensor store (full tiles),
          // or straight to global for the ragged last tile of a job (a TMA store would spill into the next job's rows)
          auto emit = [&](const float (&val)[32], float* gptr, const CUtensorMap* map) {
            if (gptr == nullptr) return;
            if (full) {
              if (issuer) tma_wait_read0();
              named_bar_sync(1 + h, 128);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t addr = stage + swz_row + ((uint32_t)(c ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(val[4 * c]), "f"(val[4 * c + 1]), "f"(val[4 * c + 2]), "f"(val[4 * c + 3]) : "memory");
              }
              fence_proxy_async();
              named_bar_sync(1 + h, 128);
              if (issuer) {
                tma_store_2d(map, h * 32, (int)grow0, stage);
                tma_commit();
              }
            } else if (row < nrows) {
              float4* dst = reinterpret_cast<float4*>(gptr + (grow0 + row) * (long)W + h * 32);
#pragma unroll
              for (int c = 0; c < 8; ++c) dst[c] = make_float4(val[4 * c], val[4 * c + 1], val[4 * c + 2], val[4 * c + 3]);
            }
          };

          emit(yh, a.o_model, &map_model);
          if (has_y) {
            float d[32], ss = 0.f, su = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              d[i] = fabsf(yh[i] - yt[i]);
              su = fmaf(d[i], d[i], su);
            }
            emit(d, a.o_tu, &map_tu);
            if (a.scale) {
              float e[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                e[i] = d[i] * vec[h * 32 + i];
                ss = fmaf(e[i], e[i], ss);
              }
              emit(e, a.o_ts, &map_ts);
            }
            if (a.o_conf) {
              float c_[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) c_[i] = d[i] * vec[W + h * 32 + i];
              emit(c_, a.o_conf, &map_conf);
            }
            if (a.o_tots || a.o_totu || a.o_totconf) {
              if (h == 1) { pair[row] = ss; pair[TILE + row] = su; }
              named_bar_sync(3, EPI_THREADS);
              if (h == 0 && row < nrows) {
                const float ts_ = (ss + pair[row]) * inv_w, tu_ = (su + pair[TILE + row]) * inv_w;
                if (a.o_tots) a.o_tots[grow0 + row] = ts_;
                if (a.o_totu) a.o_totu[grow0 + row] = tu_;
                if (a.o_totconf) a.o_totconf[grow0 + row] = ts_ / __ldg(a.agg_thr + job.slot);
              }
              named_bar_sync(3, EPI_THREADS);
            }
          }
        }
      }
      if (issuer) tma_wait_read0();  // staging boxes must outlive the TMA reads before the next item restages smem
    }
    __syncthreads();
  }

  if (!is_ctrl && (warp & 3) == 0 && lane == 0) tma_wait_all0();
  tc_fence_before();
  __syncthreads();
  if (is_ctrl) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows][64] fp32 row-major viewed as a 2-D tensor; box = 32 columns x 128 rows, SWIZZLE_128B
int make_map(CUtensorMap* map, const void* base, int64_t rows) {
  EncodeTiledFn fn = get_encode_fn();
  GB_REQUIRE(fn != nullptr, GB_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)W * sizeof(float)};
  cuuint32_t box[2] = {32, TILE};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_REQUIRE(r == CUDA_SUCCESS, GB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GB_OK;
}

}  // namespace

extern "C" int gb_ffae_tc_supported(const gb_ffnet* net) {
  if (gb::validate_ffnet(net) != GB_OK) return GB_E_SHAPE;
  const int L = net->n_layers;
  if (L > MAXL || net->dims[0] != W || net->dims[L] != W) {
    gb::set_error("tcgen05 variant covers 64-tag autoencoders with at most %d layers", MAXL);
    return GB_E_SHAPE;
  }
  for (int l = 1; l < L; ++l)
    if (net->dims[l] > W) {
      gb::set_error("tcgen05 variant needs hidden widths <= %d", W);
      return GB_E_SHAPE;
    }
  return GB_OK;
}

// rows of x / y and of the output arrays are needed for the TMA tensor maps
extern "C" int gb_ffae_infer_score_tc(const gb_ffnet* net, const float* params, const gb_job* jobs, int32_t n_jobs, int32_t max_rows,
                                      int64_t n_x_rows, int64_t n_out_rows, const float* x, const float* y, const float* scale,
                                      const float* feat_thr, const float* agg_thr, float* out_model, float* out_tag_scaled,
                                      float* out_tag_unscaled, float* out_total_scaled, float* out_total_unscaled, float* out_conf,
                                      float* out_total_conf, int32_t flags, void* stream) {
  int rc = gb_ffae_tc_supported(net);
  if (rc != GB_OK) return rc;
  GB_REQUIRE(n_x_rows > 0 && n_out_rows > 0, GB_E_ARG, "the tcgen05 variant needs the row counts of x and of the outputs");
  TcArgs a{};
  const int L = net->n_layers;
  a.n_layers = L;
  const int dbg_last = (flags >> 8) & 0xff;
  a.last_layer = (dbg_last > 0 && dbg_last <= L) ? dbg_last - 1 : L - 1;
  a.flags = flags & 0xff;
  int ofs = 0, pofs = 0;
  for (int l = 0; l < L; ++l) {
    a.K[l] = net->dims[l];
    a.N[l] = net->dims[l + 1];
    a.Np[l] = gb::round_up(a.N[l], 16);
    a.k8[l] = gb::round_up(a.K[l], 8) / 8;
    a.k16[l] = gb::round_up(a.K[l], 16) / 16;
    a.act[l] = net->act[l];
    a.pofs[l] = pofs;
    pofs += a.K[l] * a.N[l] + a.N[l];
    a.whi_ofs[l] = ofs;
    ofs += a.k8[l] * 8 * a.Np[l] * 4;
    a.wlo_ofs[l] = ofs;
    ofs += a.k16[l] * 16 * a.Np[l] * 2;
  }
  for (int l = 0; l < L; ++l) {
    a.bias_ofs[l] = ofs;
    ofs += 64 * 4;  // padded to the widest layer so float4 reads never leave the zero-filled region
  }
  a.w_bytes = gb::round_up(ofs, 16);
  ofs = a.w_bytes;
  a.vec_ofs = ofs; ofs += 2 * W * 4;
  a.pair_ofs = ofs; ofs += 2 * TILE * 4;
  a.bar_ofs = ofs; ofs += 64;
  ofs = gb::round_up(ofs, 1024);
  a.xbox_ofs = ofs; ofs += 2 * BOX_BYTES;
  a.ybox_ofs = ofs; ofs += 2 * BOX_BYTES;
  a.stage_ofs = ofs; ofs += 2 * BOX_BYTES;
  const size_t smem = (size_t)ofs;
  GB_REQUIRE(smem <= 227 * 1024, GB_E_SMEM, "architecture needs %zu bytes of shared memory in the tcgen05 variant", smem);

  int dev = 0, sms = 148;
  GB_CUDA_CHECK(cudaGetDevice(&dev));
  GB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles_per_job = (max_rows + TILE - 1) / TILE;
  int tiles_per_chunk = tiles_per_job;
  // enough work items for every SM, long enough chunks to amortise the weight staging (~16 tiles)
  while (tiles_per_chunk > 16 && (long)n_jobs * ((tiles_per_job + tiles_per_chunk - 1) / tiles_per_chunk) < 4L * sms) tiles_per_chunk = (tiles_per_chunk + 1) / 2;
  if (tiles_per_chunk > 32) tiles_per_chunk = 32;
  a.rows_per_chunk = tiles_per_chunk * TILE;
  a.chunks_per_job = (tiles_per_job + tiles_per_chunk - 1) / tiles_per_chunk;
  a.n_jobs = n_jobs;
  a.pstride = (long)gb_ffnet_param_stride(net);
  a.params = params; a.jobs = jobs; a.y = y; a.scale = scale; a.feat_thr = feat_thr; a.agg_thr = agg_thr;
  a.o_model = out_model; a.o_ts = out_tag_scaled; a.o_tu = out_tag_unscaled; a.o_conf = out_conf;
  a.o_tots = out_total_scaled; a.o_totu = out_total_unscaled; a.o_totconf = out_total_conf;

  CUtensorMap mx, my, mm, mts, mtu, mc;
  if ((rc = make_map(&mx, x, n_x_rows)) != GB_OK) return rc;
  if ((rc = make_map(&my, y ? y : x, n_x_rows)) != GB_OK) return rc;
  if ((rc = make_map(&mm, out_model, n_out_rows)) != GB_OK) return rc;
  if ((rc = make_map(&mts, out_tag_scaled ? out_tag_scaled : out_model, n_out_rows)) != GB_OK) return rc;
  if ((rc = make_map(&mtu, out_tag_unscaled ? out_tag_unscaled : out_model, n_out_rows)) != GB_OK) return rc;
  if ((rc = make_map(&mc, out_conf ? out_conf : out_model, n_out_rows)) != GB_OK) return rc;

  const long items = (long)n_jobs * a.chunks_per_job;
  const int grid = (int)(items < sms ? items : sms);
  GB_CUDA_CHECK(cudaFuncSetAttribute(ffae_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ffae_tc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a, mx, my, mm, mts, mtu, mc);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}
