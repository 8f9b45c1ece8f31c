// K6: optional smoothing of the anomaly columns (reference diff.py:302-308, 387-415):
//   smm  = rolling(window).median()   sma = rolling(window).mean()   (first window-1 rows NaN, pandas min_periods=window)
//   ewma = ewm(span=window).mean()    (pandas default adjust=True: y_t = sum_i (1-a)^i x_{t-i} / sum_i (1-a)^i, a = 2/(window+1))
// One thread per (job, column) walks the rows in order; lanes run along columns so every step of a warp reads one
// contiguous row segment.  sma/ewma keep their running sums in double (pandas does them in float64); smm keeps the
// window sorted in shared memory and replaces one element per step.  Correct-first: the rolling median in particular is
// a simple O(window) update per row and is the known slow spot of this optional path.
#include <math_constants.h>
#include "gb_common.cuh"

namespace {

constexpr int SM_THREADS = 64;

__global__ void __launch_bounds__(SM_THREADS) smooth_mean_kernel(const gb_job* jobs, const float* arr, int n_cols, int window, int method,
                                                                  float* out) {
  const gb_job job = jobs[blockIdx.y];
  const int j = blockIdx.x * SM_THREADS + threadIdx.x;
  if (j >= n_cols) return;
  const float* src = arr + job.out_row * (long)n_cols + j;
  float* dst = out + job.out_row * (long)n_cols + j;
  const int n = job.n_rows;
  if (method == 1) {  // simple moving average
    double sum = 0.0;
    int bad = 0;  // NaNs currently inside the window
    for (int t = 0; t < n; ++t) {
      const float v = src[(long)t * n_cols];
      if (v == v) sum += (double)v; else ++bad;
      if (t >= window) {
        const float old = src[(long)(t - window) * n_cols];
        if (old == old) sum -= (double)old; else --bad;
      }
      dst[(long)t * n_cols] = (t >= window - 1 && bad == 0) ? (float)(sum / (double)window) : CUDART_NAN_F;
    }
  } else {  // exponentially weighted, adjust=True
    const double decay = 1.0 - 2.0 / ((double)window + 1.0);
    double num = 0.0, den = 0.0;
    for (int t = 0; t < n; ++t) {
      const float v = src[(long)t * n_cols];
      num = num * decay + (double)v;
      den = den * decay + 1.0;
      dst[(long)t * n_cols] = (float)(num / den);
    }
  }
}

// rolling median: sorted window per thread in shared memory, [window][SM_THREADS] so lanes hit different banks
__global__ void __launch_bounds__(SM_THREADS) smooth_median_kernel(const gb_job* jobs, const float* arr, int n_cols, int window, float* out) {
  extern __shared__ float s_win[];
  const gb_job job = jobs[blockIdx.y];
  const int j = blockIdx.x * SM_THREADS + threadIdx.x;
  if (j >= n_cols) return;
  float* win = s_win + threadIdx.x;  // element i at win[i * SM_THREADS]
  const float* src = arr + job.out_row * (long)n_cols + j;
  float* dst = out + job.out_row * (long)n_cols + j;
  const int n = job.n_rows;
  int count = 0;
  for (int t = 0; t < n; ++t) {
    const float v = src[(long)t * n_cols];
    if (count == window) {  // drop the value leaving the window (first element equal to it)
      const float old = src[(long)(t - window) * n_cols];
      int lo = 0, hi = count;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (win[mid * SM_THREADS] < old) lo = mid + 1; else hi = mid;
      }
      for (int i = lo; i + 1 < count; ++i) win[i * SM_THREADS] = win[(i + 1) * SM_THREADS];
      --count;
    }
    int pos = count;  // insert keeping the window sorted
    while (pos > 0 && win[(pos - 1) * SM_THREADS] > v) {
      win[pos * SM_THREADS] = win[(pos - 1) * SM_THREADS];
      --pos;
    }
    win[pos * SM_THREADS] = v;
    ++count;
    float m = CUDART_NAN_F;
    if (count == window) {
      const int h = window >> 1;
      m = (window & 1) ? win[h * SM_THREADS] : 0.5f * (win[(h - 1) * SM_THREADS] + win[h * SM_THREADS]);
    }
    dst[(long)t * n_cols] = m;
  }
}

// x'[r][c] = x[r][c] * a[slot][c] + b[slot][c] in double, rounded once to float: what sklearn's per-feature scalers compute
// in float64 before Keras casts the batch to floatx.
__global__ void affine_f64_kernel(const gb_job* __restrict__ jobs, const double* __restrict__ x, int n_cols, const double* __restrict__ a,
                                  const double* __restrict__ b, float* __restrict__ out) {
  const gb_job job = jobs[blockIdx.y];
  const long total = (long)job.n_rows * n_cols;
  const double* src = x + (long)job.x_row * n_cols;
  float* dst = out + (long)job.out_row * n_cols;
  const double* ja = a + (long)job.slot * n_cols;
  const double* jb = b + (long)job.slot * n_cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % n_cols);
    dst[i] = (float)(__dmul_rn(src[i], ja[c]) + jb[c]);  // two roundings like numpy's X *= scale; X += min (no fma contraction)
  }
}

// q-quantile of the non-NaN values of one column of one job (pandas Series.quantile, interpolation="linear"):
// compacted into shared memory, padded with +inf to a power of two, bitonic sort, linear interpolation at (n-1)*q.
constexpr int Q_THREADS = 1024;
__global__ void __launch_bounds__(Q_THREADS) quantile_kernel(const gb_job* __restrict__ jobs, const float* __restrict__ arr, int n_cols, float q,
                                                             float* __restrict__ out, int cap) {
  extern __shared__ float sv[];
  __shared__ int s_n;
  const gb_job job = jobs[blockIdx.y];
  const int col = blockIdx.x, tid = threadIdx.x;
  const float* src = arr + (long)job.out_row * n_cols + col;
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int r = tid; r < job.n_rows; r += Q_THREADS) {
    const float v = src[(long)r * n_cols];
    if (v == v) sv[atomicAdd(&s_n, 1)] = v;  // order is irrelevant before a sort
  }
  __syncthreads();
  const int n = s_n;
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  for (int i = n + tid; i < p2; i += Q_THREADS) sv[i] = __int_as_float(0x7f800000);
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < p2; i += Q_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sv[i], b = sv[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sv[i] = b; sv[ixj] = a; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) {
    float res = __int_as_float(0x7fc00000);  // all-NaN / empty column -> NaN, as pandas
    if (n > 0) {
      const double pos = (double)(n - 1) * (double)q;
      const int lo = (int)floor(pos), hi = min(lo + 1, n - 1);
      const double frac = pos - (double)lo;
      res = (float)((double)sv[lo] + ((double)sv[hi] - (double)sv[lo]) * frac);
    }
    out[(long)blockIdx.y * n_cols + col] = res;
  }
  (void)cap;
}

}  // namespace

extern "C" int gb_quantile(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* arr, int32_t n_cols, float q, float* out,
                           void* stream) {
  GB_REQUIRE(jobs && arr && out, GB_E_ARG, "jobs/arr/out must be non-NULL");
  GB_REQUIRE(n_cols >= 1 && n_cols <= 65535 && max_rows >= 0, GB_E_ARG, "n_cols=%d max_rows=%d", n_cols, max_rows);
  GB_REQUIRE(q >= 0.f && q <= 1.f, GB_E_ARG, "percentiles should all be in the interval [0, 1], got %g", (double)q);
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535, GB_E_ARG, "bad n_jobs");
  if (n_jobs == 0) return GB_OK;
  int cap = 1;
  while (cap < max_rows) cap <<= 1;
  const size_t smem = (size_t)cap * sizeof(float);
  GB_REQUIRE(smem <= 200 * 1024, GB_E_SMEM, "quantile over %d rows per job does not fit in shared memory (limit 32768 rows)", max_rows);
  GB_CUDA_CHECK(cudaFuncSetAttribute(quantile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  quantile_kernel<<<dim3(n_cols, n_jobs), Q_THREADS, smem, (cudaStream_t)stream>>>(jobs, arr, n_cols, q, out, cap);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

extern "C" int gb_affine_f64(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const double* x, int32_t n_cols, const double* a,
                             const double* b, float* out, void* stream) {
  GB_REQUIRE(jobs && x && a && b && out, GB_E_ARG, "jobs/x/a/b/out must be non-NULL");
  GB_REQUIRE(n_cols >= 1 && max_rows >= 0, GB_E_ARG, "n_cols=%d max_rows=%d", n_cols, max_rows);
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535, GB_E_ARG, "bad n_jobs");
  if (n_jobs == 0 || max_rows == 0) return GB_OK;
  const long per_job = (long)max_rows * n_cols;
  const int bx = (int)((per_job + 256L * 8 - 1) / (256L * 8));
  const dim3 grid(bx < 1 ? 1 : (bx > 1184 ? 1184 : bx), n_jobs);
  affine_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs, x, n_cols, a, b, out);
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

extern "C" int gb_smooth(const gb_job* jobs, int32_t n_jobs, const float* arr, int32_t n_cols, int32_t window, int32_t method, float* out,
                         void* stream) {
  GB_REQUIRE(jobs && arr && out, GB_E_ARG, "jobs/arr/out must be non-NULL");
  GB_REQUIRE(n_cols >= 1 && window >= 1, GB_E_ARG, "n_cols=%d window=%d must be >= 1", n_cols, window);
  GB_REQUIRE(method >= 0 && method <= 2, GB_E_ARG, "method=%d unknown (0 smm, 1 sma, 2 ewma)", method);
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535, GB_E_ARG, "bad n_jobs");
  if (n_jobs == 0) return GB_OK;
  const dim3 grid((n_cols + SM_THREADS - 1) / SM_THREADS, n_jobs);
  if (method == 0) {
    const size_t smem = (size_t)window * SM_THREADS * sizeof(float);
    GB_REQUIRE(smem <= 200 * 1024, GB_E_SMEM, "rolling-median window %d does not fit in shared memory", window);
    GB_CUDA_CHECK(cudaFuncSetAttribute(smooth_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smooth_median_kernel<<<grid, SM_THREADS, smem, (cudaStream_t)stream>>>(jobs, arr, n_cols, window, out);
  } else {
    smooth_mean_kernel<<<grid, SM_THREADS, 0, (cudaStream_t)stream>>>(jobs, arr, n_cols, window, method, out);
  }
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}
