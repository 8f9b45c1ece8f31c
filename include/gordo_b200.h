/*
 * gordo_b200.h -- C ABI of the B200 (sm_100a) implementation of gordo's per-machine
 * autoencoder train-and-score hot path.
 *
 * The reference (equinor/gordo-components) has no native FFI: its plug-in boundary is a
 * Python class path + the sklearn/GordoBase protocol (gordo/machine/model/base.py:10-35,
 * gordo/machine/model/anomaly/base.py:11-23, gordo/serializer/from_definition.py:176-191).
 * This header is the seam *below* that protocol: every arithmetic library call the
 * reference makes on the hot path (Keras Model.predict / Model.fit, sklearn MinMaxScaler,
 * pandas rolling/abs/mean in DiffBasedAnomalyDetector) maps to one entry point here.
 * Each entry point cites the reference call it replaces.  INTEGRATION.md shows the
 * ctypes binding (gordo_components_b200/_cabi.py is the live copy).
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch types.  Unless stated otherwise every
 *    pointer is a DEVICE pointer owned by the caller; nothing here allocates or frees.
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns
 *    without synchronising.  Return value: 0 or a negative gb_status; the message of
 *    the last failure on the calling thread is available from gb_last_error().
 *  - float32 everywhere on device; row-major; rows of all machines are concatenated:
 *    x[total_rows][n_features], y / outputs [total_rows][n_features_out].
 *  - a *slot* is one trained network (a machine, or one CV fold of a machine); a *job*
 *    binds a slot to a contiguous range of rows.
 */
#ifndef GORDO_B200_H
#define GORDO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_ABI_VERSION 2
#define GB_MAX_LAYERS 16
#define GB_MAX_WIDTH 256 /* widest layer / feature count (the feedforward_model / feedforward_symmetric defaults are 256-128-64) */

typedef enum gb_status {
  GB_OK = 0,
  GB_E_ARG = -1,     /* null pointer / bad enum / inconsistent sizes          -> ValueError  */
  GB_E_SHAPE = -2,   /* architecture outside what the kernels support          -> ValueError  */
  GB_E_ALIGN = -3,   /* pointer not 16-byte aligned                            -> ValueError  */
  GB_E_SMEM = -4,    /* architecture does not fit in shared memory             -> ValueError  */
  GB_E_CUDA = -5,    /* CUDA runtime error (message has cudaGetErrorString)    -> RuntimeError*/
  GB_E_DEVICE = -6   /* no sm_100 device                                       -> RuntimeError*/
} gb_status;

typedef enum gb_act { GB_ACT_LINEAR = 0, GB_ACT_TANH = 1, GB_ACT_RELU = 2, GB_ACT_SIGMOID = 3 } gb_act;

/* Dense stack built by feedforward_model / feedforward_symmetric / feedforward_hourglass
 * (gordo/machine/model/factories/feedforward_autoencoder.py:65-104).  dims[0] = n_features,
 * dims[l+1] = units of Dense layer l; l1[l] = activity_regularizer l1 coefficient of layer l
 * (10e-5 on encoder layers i>=1, :80-81), used by training only. */
typedef struct gb_ffnet {
  int32_t n_layers;
  int32_t dims[GB_MAX_LAYERS + 1];
  int32_t act[GB_MAX_LAYERS];
  float l1[GB_MAX_LAYERS];
} gb_ffnet;

/* One unit of work: network `slot` applied to rows [x_row, x_row + n_rows) of x / y,
 * results written to rows [out_row, out_row + n_rows) of the output arrays. */
typedef struct gb_job {
  int32_t slot;
  int32_t n_rows;
  int64_t x_row;
  int64_t out_row;
} gb_job;

/* ---- library ------------------------------------------------------------------------ */
int gb_abi_version(void);
const char* gb_last_error(void);
/* 0 if `device` is an sm_100 part; GB_E_DEVICE otherwise.  Fills sm count if non-null. */
int gb_device_check(int device, int* sm_count);

/* ---- parameter layout ---------------------------------------------------------------
 * Canonical per-slot parameter vector ("Keras order"): for each layer l: kernel
 * W_l[dims[l]][dims[l+1]] row-major (Keras Dense kernel layout [in,out]) then bias
 * b_l[dims[l+1]].  Slots are `gb_ffnet_param_stride()` floats apart (count rounded up to 4). */
size_t gb_ffnet_param_count(const gb_ffnet* net);
size_t gb_ffnet_param_stride(const gb_ffnet* net);

/* ---- K1+K4: predict + anomaly score, fused --------------------------------------------
 * Replaces, per job: KerasBaseEstimator.predict -> keras Model.predict
 * (gordo/machine/model/models.py:289-300) and the arithmetic of
 * DiffBasedAnomalyDetector.anomaly (gordo/machine/model/anomaly/diff.py:350-385, 420-444):
 *   out_model            = net(x)                                   [rows][n_out]
 *   out_tag_unscaled     = |out_model - y|                          [rows][n_out]
 *   out_tag_scaled       = |out_model - y| * scale[slot]            [rows][n_out]   (MinMax offset cancels)
 *   out_total_unscaled   = mean_j(out_tag_unscaled^2)               [rows]
 *   out_total_scaled     = mean_j(out_tag_scaled^2)                 [rows]
 *   out_conf             = out_tag_unscaled / feat_thr[slot]        [rows][n_out]   (diff.py:421 uses the unscaled diff)
 *   out_total_conf       = out_total_scaled / agg_thr[slot]         [rows]
 * y == NULL -> prediction only (all score outputs must be NULL).  Any score output may be
 * NULL and is then skipped; feat_thr / agg_thr NULL -> the confidences must be NULL.
 * params: [n_slots][param_stride]; scale, feat_thr: [n_slots][n_out]; agg_thr: [n_slots].
 * jobs: DEVICE array of n_jobs gb_job; max_rows = max n_rows over jobs (host value, sizes the grid).
 * n_x_rows / n_out_rows: number of rows of the x (and y) array and of the output arrays (TMA tensor extents).
 * variant (low byte): 0 = auto (tcgen05 kernel for the stacks it covers, the row-per-thread kernel for stacks whose widths are
 * all <= 16, else the generic one), 1 = generic fp32 CUDA-core kernel, 2 = tcgen05 split-precision kernel, 3 = row-per-thread
 * fp32 kernel (2 / 3: GB_E_SHAPE if the architecture is outside their range); higher bytes are debug knobs and must be 0. */
int gb_ffae_infer_score(const gb_ffnet* net, const float* params, const gb_job* jobs, int32_t n_jobs,
                        int32_t max_rows, int64_t n_x_rows, int64_t n_out_rows, const float* x, const float* y, const float* scale,
                        const float* feat_thr, const float* agg_thr, float* out_model,
                        float* out_tag_scaled, float* out_tag_unscaled, float* out_total_scaled,
                        float* out_total_unscaled, float* out_conf, float* out_total_conf,
                        int32_t variant, void* stream);

/* GB_OK if the tcgen05 (variant 2) kernel covers this architecture, else GB_E_SHAPE. */
int gb_ffae_tc_supported(const gb_ffnet* net);

/* ---- K4 alone: anomaly score of predictions that already exist ----------------------------
 * Same outputs as gb_ffae_infer_score, for a `yhat` produced elsewhere (a base estimator that is not
 * one of ours, e.g. the sklearn regressors the reference's detector tests use; an LSTM prediction from
 * gb_lstm_infer).  yhat and all outputs are indexed by out_row, y by x_row. */
int gb_anomaly_score(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* yhat, const float* y,
                     int32_t n_out, const float* scale, const float* feat_thr, const float* agg_thr,
                     float* out_tag_scaled, float* out_tag_unscaled, float* out_total_scaled,
                     float* out_total_unscaled, float* out_conf, float* out_total_conf, void* stream);

/* The same arithmetic in float64, as the reference does it (pandas on float64 y: diff.py:268-300 `_scaled_mse_per_timestep`,
 * `_absolute_error`; :350-385, 420-444 in `anomaly`) -- used whenever the predictions did not come out of one of this
 * package's fp32 networks fused with the scoring (foreign base estimators, LSTM outputs): at data magnitude ~100 a float32
 * |yhat - y| is uncertain by 7.6e-6, i.e. percents of a small residual.  All arrays double. */
int gb_anomaly_score_f64(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const double* yhat, const double* y,
                         int32_t n_out, const double* scale, const double* feat_thr, const double* agg_thr,
                         double* out_tag_scaled, double* out_tag_unscaled, double* out_total_scaled,
                         double* out_total_unscaled, double* out_conf, double* out_total_conf, void* stream);

/* ---- K7: MinMaxScaler.fit on the targets (diff.py:173; sklearn MinMaxScaler [3P]) -------
 * per job: scale[slot][j] = 1/(max_j - min_j) (zero range -> 1), offset[slot][j] = -min_j*scale.
 * minmax_ws: workspace [n_slots][2][n_out] floats (overwritten). */
int gb_minmax_fit(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* y, int32_t n_out,
                  float* scale, float* offset, float* minmax_ws, int32_t n_slots, void* stream);

/* Column extrema of float64 targets, for the scaler of a single detector (`scaler.fit(y)` sees float64 y in the reference):
 * minmax[slot][0][j] = min, minmax[slot][1][j] = max over the job's rows (NaNs skipped; +inf / -inf when there is no finite
 * sample).  sklearn's float64 scale_ / min_ arithmetic on them stays on the host. */
int gb_minmax_f64(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const double* y, int32_t n_out, double* minmax,
                  int32_t n_slots, void* stream);

/* ---- K5: thresholds of one CV fold (diff.py:222-233) ------------------------------------
 *   feat_thr[slot][j] = max_t min(tag_unscaled[t-window+1 .. t][j])   (rolling(window).min().max())
 *   agg_thr[slot]     = max_t min(total_scaled[t-window+1 .. t])
 * rows are taken at [out_row, out_row+n_rows) of the score arrays produced by
 * gb_ffae_infer_score for the fold's test rows.  n_rows < window -> NaN (pandas semantics). */
int gb_thresholds(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* tag_unscaled,
                  const float* total_scaled, int32_t n_out, int32_t window, float* feat_thr,
                  float* agg_thr, int32_t n_slots, void* stream);

/* float64 form for the score arrays of gb_anomaly_score_f64 (min / max select, so the thresholds are exact). */
int gb_thresholds_f64(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const double* tag_unscaled,
                      const double* total_scaled, int32_t n_out, int32_t window, double* feat_thr,
                      double* agg_thr, int32_t n_slots, void* stream);

/* ---- K8: column moments behind the builder's cross-validation metrics (build_model.py:250-289, 378-446) ----
 * For job i, over rows yhat[out_row .. out_row+n_rows) and y[x_row .. x_row+n_rows), with e = yhat - y and
 * y0 = the job's first target row:  out[i][q][j] (double) = q0: sum e, q1: sum e^2, q2: sum |e|,
 * q3: sum (y - y0), q4: sum (y - y0)^2.  explained_variance / r2 / mean_squared_error / mean_absolute_error
 * (per tag and uniform-averaged, under any per-tag affine scoring scaler) follow from these on the host. */
int gb_cv_moments(const gb_job* jobs, int32_t n_jobs, const float* yhat, const float* y, int32_t n_out,
                  double* out, void* stream);

/* ---- K6: optional smoothing of anomaly columns (diff.py:302-308, 387-415) ------------------
 * method 0 = smm rolling(window).median(), 1 = sma rolling(window).mean() (first window-1 rows NaN; a window holding a NaN
 * gives NaN), 2 = ewma ewm(span=window).mean() (adjust=True, ignore_na=False: a NaN adds no observation, ages the weights and the
 * previous average is carried forward).  arr / out: [rows][n_cols], rows taken at [out_row, out_row+n_rows); max_rows = max n_rows
 * over jobs (sizes the row-chunk grid of the rolling kernels).  The rolling median keeps one sorted window per thread in shared
 * memory: windows up to 51 200 rows. */
int gb_smooth(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* arr, int32_t n_cols, int32_t window, int32_t method,
              float* out, void* stream);

/* ---- K9: percentile thresholds of DiffBasedKFCVAnomalyDetector (diff.py:623-635: smoothed validation metric
 * .quantile(threshold_percentile)).  out[job][c] = q-quantile (linear interpolation, NaNs skipped -- pandas
 * semantics) of column c of rows [out_row, out_row+n_rows) of arr [rows][n_cols].  Jobs of up to 32768 rows are sorted in shared
 * memory; longer ones (a year of 10-minute data is ~52k rows) find the two order statistics by radix selection over L2. */
int gb_quantile(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const float* arr, int32_t n_cols, float q,
                float* out, void* stream);

/* ---- K8: per-feature affine pre-transform (sklearn MinMaxScaler/StandardScaler/... .transform in front of the
 * network inside a Pipeline: gordo serializer pipelines, e.g. examples/config_crd.yaml "sklearn.preprocessing.MinMaxScaler")
 *   out[out_row+r][c] = (float)(x[x_row+r][c] * a[slot][c] + b[slot][c]), computed in double like sklearn and rounded once,
 * i.e. exactly the float32 batch Keras sees.  (Folding a/b into the first Dense layer instead would cancel catastrophically
 * in fp32 for offset-dominated tags, so the transform stays a separate f64 pass.) */
int gb_affine_f64(const gb_job* jobs, int32_t n_jobs, int32_t max_rows, const double* x, int32_t n_cols, const double* a,
                  const double* b, float* out, void* stream);

/* ---- K2: fit ---------------------------------------------------------------------------
 * Replaces scikeras KerasRegressor.fit -> keras Model.fit (models.py:284) for the Dense
 * stacks above: per job, `epochs` passes over rows [x_row, x_row+n_rows) in batches of
 * `batch_size`, loss = mean((net(x)-y)^2) + sum_l l1[l]*sum|a_l|, Adam (Keras defaults).
 * One CTA per job trains the whole fit with weights resident in shared memory. */
typedef struct gb_fit_hparams {
  int32_t epochs;
  int32_t batch_size;
  int32_t shuffle;        /* 0: sequential order; 1: on-device keyed permutation per (seed, slot, epoch);
                             2: explicit `perm` (parity testing / reproducing a given order) */
  int32_t l1_div_batch;   /* 0: keras 3.3.3 behaviour (activity loss not divided by batch size) */
  float lr, beta1, beta2, eps;
  uint64_t seed;
  int32_t step0;          /* Adam step count already taken (warm start); 0 for a fresh fit */
  int32_t reserved;
} gb_fit_hparams;

/* Adam moments are opaque optimizer state in the kernel's padded layout: gb_ffae_fit_state_stride() floats per slot. */
size_t gb_ffae_fit_state_stride(const gb_ffnet* net);

/* params: [n_slots][param_stride], updated in place.  adam_m / adam_v: [n_slots][state_stride], updated in place
 * (all zero for a fresh fit).  Mini-batches above 32 rows are processed as 32-row chunks whose gradients are summed
 * before the optimizer step (the state arrays carry the scratch for that).
 * perm: [n_jobs][epochs][max_rows] int32 row indices relative to the job (shuffle == 2), else NULL.
 * out_loss / out_acc: [n_jobs][epochs] per-epoch sample-weighted mean loss / categorical accuracy
 * (keras History.history["loss"], ["accuracy"], models.py:339-357). */
int gb_ffae_fit(const gb_ffnet* net, float* params, float* adam_m, float* adam_v, const gb_job* jobs,
                int32_t n_jobs, int32_t max_rows, const float* x, const float* y, const int32_t* perm,
                const gb_fit_hparams* hp, float* out_loss, float* out_acc, void* stream);

/* ---- K3: LSTM autoencoder predict --------------------------------------------------------
 * lstm_model / lstm_symmetric / lstm_hourglass (factories/lstm_autoencoder.py:72-103):
 * LSTM layers (gate order i,f,c,o; sigmoid recurrent activation; zero initial state per
 * window) then Dense.  Replaces KerasLSTMBaseEstimator.predict (models.py:618-660) without
 * materialising windows (create_keras_timeseriesgenerator, models.py:713-793): output row j
 * of a job is the network applied to x rows [x_row + j, x_row + j + lookback). */
typedef struct gb_lstmnet {
  int32_t n_layers;                 /* LSTM layers */
  int32_t n_features, n_features_out;
  int32_t units[GB_MAX_LAYERS];
  int32_t act[GB_MAX_LAYERS];       /* cell/output activation of each LSTM layer (tanh default) */
  int32_t out_act;                  /* Dense activation */
  int32_t lookback;
} gb_lstmnet;

/* per-slot parameter vector: for each layer: kernel [in][4u], recurrent_kernel [u][4u], bias [4u];
 * then Dense kernel [u_last][n_out], bias [n_out]. */
size_t gb_lstm_param_count(const gb_lstmnet* net);
size_t gb_lstm_param_stride(const gb_lstmnet* net);
/* jobs: n_rows = number of *windows* (output rows); x rows read = n_rows + lookback - 1.
 * workspace: gb_lstm_workspace_bytes() bytes of device scratch. */
size_t gb_lstm_workspace_bytes(const gb_lstmnet* net, int32_t n_jobs, int32_t max_rows);
int gb_lstm_infer(const gb_lstmnet* net, const float* params, const gb_job* jobs, int32_t n_jobs,
                  int32_t max_rows, const float* x, float* out_model, void* workspace, void* stream);

/* ---- K3 on tcgen05 (layer widths 1..512, padded to multiples of 64 internally).  One launch per
 * (layer, timestep) advances every window of every job: [h_below,t | h_own,t-1] . [K; U]^T on the tensor cores
 * (FP16-pair split operands, fp32 accumulation in TMEM), LSTM cell in the epilogue, recurrent state in `workspace`
 * (gb_lstm_tc_workspace_bytes; x_rows = rows of the x array, n_slots = rows of params). */
int gb_lstm_tc_supported(const gb_lstmnet* net);
size_t gb_lstm_tc_workspace_bytes(const gb_lstmnet* net, int32_t n_slots, int32_t n_jobs, int32_t max_windows, int64_t x_rows);
int gb_lstm_infer_tc(const gb_lstmnet* net, const float* params, int32_t n_slots, const gb_job* jobs, int32_t n_jobs,
                     int32_t max_windows, const float* x, int64_t x_rows, float* out_model, void* workspace, void* stream);

/* ---- K3-fit: LSTM training (back-propagation through time) ---------------------------------
 * Replaces KerasLSTMBaseEstimator.fit (models.py:557-616): if `primer`, one Adam step on the single
 * window 0 (the reference's `super().fit` on a batch of one, :585-597); then `epochs` passes over the
 * windows in order (shuffle=False, :612-615) in batches of `batch_size` (last partial batch kept),
 * loss = mean((net(window) - target)^2), Adam as in gb_ffae_fit.  Window j of a job = x rows
 * [x_row + j, x_row + j + lookback), target = y row x_row + j + lookback - 1 + lookahead; jobs[].n_rows
 * counts windows.  params / adam_m / adam_v: [n_slots][gb_lstm_param_stride], updated in place;
 * adam_t: [n_slots] optimizer step counters (in/out).  out_loss / out_acc: [n_jobs][epochs].
 * workspace: gb_lstm_fit_workspace_bytes(net, n_jobs) bytes of device scratch (saved gates/states of
 * one batch per job).  One optimizer step is a sequence of launches over (tile, job) grids. */
typedef struct gb_lstm_fit_hparams {
  int32_t epochs, batch_size;   /* batch_size <= 32 */
  int32_t lookahead;            /* 0 = KerasLSTMAutoEncoder, 1 = KerasLSTMForecast */
  int32_t primer;               /* 1 = run the reference's primer step first */
  float lr, beta1, beta2, eps;
} gb_lstm_fit_hparams;
size_t gb_lstm_fit_workspace_bytes(const gb_lstmnet* net, int32_t n_jobs);
int gb_lstm_fit(const gb_lstmnet* net, float* params, float* adam_m, float* adam_v, int32_t* adam_t,
                const gb_job* jobs, int32_t n_jobs, int32_t max_windows, const float* x, const float* y,
                const gb_lstm_fit_hparams* hp, void* workspace, float* out_loss, float* out_acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GORDO_B200_H */
