// K5 + K7: the two small column reductions of the anomaly path.
//   gb_minmax_fit : sklearn MinMaxScaler.fit on the targets (reference diff.py:173)
//   gb_thresholds : rolling(window).min().max() per tag and for the aggregate series (diff.py:222-233)
// Both are HBM-bound single passes over [rows][n_out] arrays: lanes run along the tag axis so every warp
// request is one contiguous segment, partial results meet in shared memory and one atomic per (CTA, column)
// publishes them.
#include <math_constants.h>
#include "gb_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int ROWS_PER_CTA = 1024;

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// ---------------------------------------------------------------- min / max per column
__global__ void minmax_init_kernel(const gb_job* jobs, int n_out, float* ws) {
  const gb_job job = jobs[blockIdx.x];
  float* w = ws + (long)job.slot * 2 * n_out;
  for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
    w[j] = CUDART_INF_F;            // running min
    w[n_out + j] = -CUDART_INF_F;   // running max
  }
}

__global__ void __launch_bounds__(THREADS) minmax_reduce_kernel(const gb_job* jobs, const float* y, int n_out, float* ws) {
  __shared__ float s_min[NWARPS][GB_MAX_WIDTH];
  __shared__ float s_max[NWARPS][GB_MAX_WIDTH];
  const gb_job job = jobs[blockIdx.y];
  const int r0 = blockIdx.x * ROWS_PER_CTA;
  if (r0 >= job.n_rows) return;
  const int r1 = min(job.n_rows, r0 + ROWS_PER_CTA);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = y + job.x_row * (long)n_out;
  for (int j0 = 0; j0 < n_out; j0 += 32) {
    const int j = j0 + lane;
    float lo = CUDART_INF_F, hi = -CUDART_INF_F;
    if (j < n_out) {
      for (int r = r0 + warp; r < r1; r += NWARPS) {
        const float v = __ldg(base + (long)r * n_out + j);
        if (v == v) { lo = fminf(lo, v); hi = fmaxf(hi, v); }
      }
      s_min[warp][j] = lo;
      s_max[warp][j] = hi;
    }
  }
  __syncthreads();
  float* w = ws + (long)job.slot * 2 * n_out;
  for (int j = threadIdx.x; j < n_out; j += THREADS) {
    float lo = s_min[0][j], hi = s_max[0][j];
#pragma unroll
    for (int q = 1; q < NWARPS; ++q) { lo = fminf(lo, s_min[q][j]); hi = fmaxf(hi, s_max[q][j]); }
    if (lo <= hi) {  // at least one finite sample
      atomic_min_float(&w[j], lo);
      atomic_max_float(&w[n_out + j], hi);
    }
  }
}

__global__ void minmax_finalize_kernel(const gb_job* jobs, int n_out, const float* ws, float* scale, float* offset) {
  const gb_job job = jobs[blockIdx.x];
  const float* w = ws + (long)job.slot * 2 * n_out;
  for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
    const float lo = w[j], hi = w[n_out + j];
    float range = hi - lo;
    // sklearn _handle_zeros_in_scale: ranges below 10*eps are treated as 1 (constant feature)
    if (!(range >= 10.f * 1.1920929e-7f)) range = 1.f;
    const float s = 1.f / range;
    scale[(long)job.slot * n_out + j] = s;
    if (offset) offset[(long)job.slot * n_out + j] = -lo * s;
  }
}

// ---------------------------------------------------------------- rolling(window).min().max() per column
__global__ void rollmax_init_kernel(const gb_job* jobs, int n_cols, float* out) {
  const gb_job job = jobs[blockIdx.x];
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) out[(long)job.slot * n_cols + j] = -1.f;
}

// arr: [rows][n_cols] (non-negative values); positions t in [window-1, n_rows) of each job.
__global__ void __launch_bounds__(THREADS) rollmin_max_kernel(const gb_job* jobs, const float* arr, int n_cols, int window,
                                                               float* out) {
  __shared__ float s_max[NWARPS][GB_MAX_WIDTH];
  const gb_job job = jobs[blockIdx.y];
  const int t0 = window - 1 + blockIdx.x * ROWS_PER_CTA;
  if (t0 >= job.n_rows) return;
  const int t1 = min(job.n_rows, t0 + ROWS_PER_CTA);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = arr + job.out_row * (long)n_cols;
  for (int j0 = 0; j0 < n_cols; j0 += 32) {
    const int j = j0 + lane;
    if (j < n_cols) {
      float best = -1.f;
      for (int t = t0 + warp; t < t1; t += NWARPS) {
        float m = CUDART_INF_F;
        bool nan = false;
        for (int i = 0; i < window; ++i) {
          const float v = __ldg(base + (long)(t - i) * n_cols + j);
          nan |= !(v == v);
          m = fminf(m, v);
        }
        if (!nan) best = fmaxf(best, m);  // pandas: a window holding NaN yields NaN, which max() skips
      }
      s_max[warp][j] = best;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_cols; j += THREADS) {
    float best = s_max[0][j];
#pragma unroll
    for (int q = 1; q < NWARPS; ++q) best = fmaxf(best, s_max[q][j]);
    if (best >= 0.f) atomicMax(reinterpret_cast<int*>(&out[(long)job.slot * n_cols + j]), __float_as_int(best));
  }
}

__global__ void rollmax_finalize_kernel(const gb_job* jobs, int n_cols, float* out) {
  const gb_job job = jobs[blockIdx.x];
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) {
    float* p = &out[(long)job.slot * n_cols + j];
    if (*p < 0.f) *p = CUDART_NAN_F;  // fewer rows than the window: pandas gives NaN
  }
}

// ---------------------------------------------------------------- score of existing predictions (diff.py:350-385, 420-444)
__global__ void __launch_bounds__(THREADS) anomaly_score_kernel(const gb_job* jobs, const float* yhat, const float* y, int n_out,
                                                                 const float* scale, const float* feat_thr,
                                                                 const float* agg_thr, float* o_ts, float* o_tu,
                                                                 float* o_tots, float* o_totu, float* o_conf,
                                                                 float* o_totconf) {
  const gb_job job = jobs[blockIdx.y];
  const int r0 = blockIdx.x * ROWS_PER_CTA;
  if (r0 >= job.n_rows) return;
  const int r1 = min(job.n_rows, r0 + ROWS_PER_CTA);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* sc = scale ? scale + (long)job.slot * n_out : nullptr;
  const float* ft = feat_thr ? feat_thr + (long)job.slot * n_out : nullptr;
  const float inv = 1.f / (float)n_out;
  for (int r = r0 + warp; r < r1; r += NWARPS) {
    const long go = (job.out_row + r) * (long)n_out, gy = (job.x_row + r) * (long)n_out;
    float ss = 0.f, su = 0.f;
    for (int j = lane; j < n_out; j += 32) {
      const float d = fabsf(__ldg(yhat + go + j) - __ldg(y + gy + j));
      if (o_tu) o_tu[go + j] = d;
      su += d * d;
      if (sc) {
        const float e = d * __ldg(sc + j);
        if (o_ts) o_ts[go + j] = e;
        ss += e * e;
      }
      if (o_conf) o_conf[go + j] = d / __ldg(ft + j);
    }
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      su += __shfl_xor_sync(0xffffffffu, su, o);
    }
    if (lane == 0) {
      if (o_tots) o_tots[job.out_row + r] = ss * inv;
      if (o_totu) o_totu[job.out_row + r] = su * inv;
      if (o_totconf) o_totconf[job.out_row + r] = ss * inv / __ldg(agg_thr + job.slot);
    }
  }
}

// ---------------------------------------------------------------- column moments of one CV fold (build_model.py:378-446)
// The builder's cross-validation metrics (explained variance, r2, MSE, MAE: per tag and averaged) are all functions of five
// per-column sums over the fold's test rows.  out[job][q][j] (double), with e = yhat - y and y0 = the job's first target row
// (a shift that keeps the second moment of y well conditioned):  q=0 sum e, 1 sum e^2, 2 sum |e|, 3 sum (y-y0), 4 sum (y-y0)^2.
// One CTA per job and a fixed summation order: the result does not depend on the launch.
__global__ void __launch_bounds__(THREADS) cv_moments_kernel(const gb_job* jobs, const float* yhat, const float* y, int n_out,
                                                              double* out) {
  __shared__ double s_acc[NWARPS][5][32];
  const gb_job job = jobs[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* yb = y + job.x_row * (long)n_out;
  const float* pb = yhat + job.out_row * (long)n_out;
  double* o = out + (long)blockIdx.x * 5 * n_out;
  for (int j0 = 0; j0 < n_out; j0 += 32) {
    const int j = j0 + lane;
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0., a4 = 0.;
    if (j < n_out && job.n_rows > 0) {
      const double y0 = (double)__ldg(yb + j);
      for (int r = warp; r < job.n_rows; r += NWARPS) {
        const double yv = (double)__ldg(yb + (long)r * n_out + j);
        const double e = (double)__ldg(pb + (long)r * n_out + j) - yv;
        const double c = yv - y0;
        a0 += e; a1 += e * e; a2 += fabs(e); a3 += c; a4 += c * c;
      }
    }
    s_acc[warp][0][lane] = a0; s_acc[warp][1][lane] = a1; s_acc[warp][2][lane] = a2; s_acc[warp][3][lane] = a3; s_acc[warp][4][lane] = a4;
    __syncthreads();
    if (threadIdx.x < 5 * 32) {
      const int q = threadIdx.x >> 5;
      if (j < n_out) {
        double t = s_acc[0][q][lane];
#pragma unroll
        for (int w = 1; w < NWARPS; ++w) t += s_acc[w][q][lane];
        o[q * n_out + j] = t;
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int gb_anomaly_score(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* yhat, const float* y,
                     int32_t n_out, const float* scale, const float* feat_thr, const float* agg_thr,
                     float* out_tag_scaled, float* out_tag_unscaled, float* out_total_scaled,
                     float* out_total_unscaled, float* out_conf, float* out_total_conf, void* stream) {
  GB_REQUIRE(jobs && yhat && y, GB_E_ARG, "jobs/yhat/y must be non-NULL");
  GB_REQUIRE(n_out >= 1, GB_E_SHAPE, "n_out=%d must be >= 1", n_out);
  GB_REQUIRE(scale || (!out_tag_scaled && !out_total_scaled && !out_total_conf), GB_E_ARG, "scaled outputs requested without scale");
  GB_REQUIRE(!out_conf || feat_thr, GB_E_ARG, "out_conf requested without feat_thr");
  GB_REQUIRE(!out_total_conf || agg_thr, GB_E_ARG, "out_total_conf requested without agg_thr");
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535, GB_E_ARG, "bad n_jobs");
  if (n_jobs == 0 || max_rows <= 0) return GB_OK;
  const int chunks = (max_rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  anomaly_score_kernel<<<dim3(chunks, n_jobs), THREADS, 0, (cudaStream_t)stream>>>(
      jobs, yhat, y, n_out, scale, feat_thr, agg_thr, out_tag_scaled, out_tag_unscaled, out_total_scaled,
      out_total_unscaled, out_conf, out_total_conf);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

int gb_minmax_fit(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* y, int32_t n_out, float* scale,
                  float* offset, float* minmax_ws, int32_t n_slots, void* stream) {
  GB_REQUIRE(jobs && y && scale && minmax_ws, GB_E_ARG, "jobs/y/scale/minmax_ws must be non-NULL");
  GB_REQUIRE(n_out >= 1 && n_out <= GB_MAX_WIDTH, GB_E_SHAPE, "n_out=%d outside [1,%d]", n_out, GB_MAX_WIDTH);
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535 && n_slots >= 0, GB_E_ARG, "bad n_jobs/n_slots");
  if (n_jobs == 0) return GB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<n_jobs, 128, 0, st>>>(jobs, n_out, minmax_ws);
  const int chunks = (max_rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  if (chunks > 0) minmax_reduce_kernel<<<dim3(chunks, n_jobs), THREADS, 0, st>>>(jobs, y, n_out, minmax_ws);
  minmax_finalize_kernel<<<n_jobs, 128, 0, st>>>(jobs, n_out, minmax_ws, scale, offset);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

int gb_thresholds(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* tag_unscaled,
                  const float* total_scaled, int32_t n_out, int32_t window, float* feat_thr, float* agg_thr,
                  int32_t n_slots, void* stream) {
  GB_REQUIRE(jobs, GB_E_ARG, "jobs must be non-NULL");
  GB_REQUIRE((tag_unscaled != nullptr) == (feat_thr != nullptr), GB_E_ARG, "tag_unscaled and feat_thr go together");
  GB_REQUIRE((total_scaled != nullptr) == (agg_thr != nullptr), GB_E_ARG, "total_scaled and agg_thr go together");
  GB_REQUIRE(n_out >= 1 && n_out <= GB_MAX_WIDTH, GB_E_SHAPE, "n_out=%d outside [1,%d]", n_out, GB_MAX_WIDTH);
  GB_REQUIRE(window >= 1, GB_E_ARG, "window=%d must be >= 1", window);
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535 && n_slots >= 0, GB_E_ARG, "bad n_jobs/n_slots");
  if (n_jobs == 0) return GB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int span = max_rows - (window - 1);
  const int chunks = span > 0 ? (span + ROWS_PER_CTA - 1) / ROWS_PER_CTA : 0;
  if (feat_thr) {
    rollmax_init_kernel<<<n_jobs, 128, 0, st>>>(jobs, n_out, feat_thr);
    if (chunks) rollmin_max_kernel<<<dim3(chunks, n_jobs), THREADS, 0, st>>>(jobs, tag_unscaled, n_out, window, feat_thr);
    rollmax_finalize_kernel<<<n_jobs, 128, 0, st>>>(jobs, n_out, feat_thr);
  }
  if (agg_thr) {
    rollmax_init_kernel<<<n_jobs, 32, 0, st>>>(jobs, 1, agg_thr);
    if (chunks) rollmin_max_kernel<<<dim3(chunks, n_jobs), THREADS, 0, st>>>(jobs, total_scaled, 1, window, agg_thr);
    rollmax_finalize_kernel<<<n_jobs, 32, 0, st>>>(jobs, 1, agg_thr);
  }
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

int gb_cv_moments(const gb_job* jobs, int32_t n_jobs, const float* yhat, const float* y, int32_t n_out, double* out,
                  void* stream) {
  GB_REQUIRE(jobs && yhat && y && out, GB_E_ARG, "jobs/yhat/y/out must be non-NULL");
  GB_REQUIRE(n_out >= 1, GB_E_SHAPE, "n_out=%d must be >= 1", n_out);
  GB_REQUIRE(n_jobs >= 0, GB_E_ARG, "bad n_jobs");
  if (n_jobs == 0) return GB_OK;
  cv_moments_kernel<<<n_jobs, THREADS, 0, (cudaStream_t)stream>>>(jobs, yhat, y, n_out, out);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

}  // extern "C"
