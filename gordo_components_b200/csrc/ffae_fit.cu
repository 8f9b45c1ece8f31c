// K2: fit of the Dense autoencoder stacks -- one persistent CTA per job trains the whole fit.
//
// Replaces scikeras KerasRegressor.fit -> keras Model.fit (gordo/machine/model/models.py:284) for
// the networks of factories/feedforward_autoencoder.py:65-104:
//   loss = mean((net(x)-y)^2) + sum_l l1[l]*sum|a_l|,  Adam(lr, b1, b2, eps) [keras defaults 1e-3/.9/.999/1e-7],
//   every epoch visits a permutation of the job's rows in batches of batch_size (last partial batch kept).
//
// A fit is a chain of epochs*ceil(n/batch) dependent optimizer steps of ~3 MFLOP each, so it is latency bound,
// not roofline bound: the design keeps everything a step needs on chip.  The slot's weights live in shared
// memory (padded [Kp][Np] image) for the entire fit, every layer's activations of the current mini-batch stay in
// shared memory for the backward pass, the next mini-batch is gathered with cp.async while the current one is
// processed, and the Adam moments (opaque state, same padded layout) stream through L2.  Machines (and CV folds,
// which are just more jobs) are independent, so the grid is simply one CTA per job.
// A mini-batch larger than 32 rows is processed as chunks of 32: the chunks' weight gradients are summed in
// an L2-resident scratch image (second half of the opaque Adam-m state) and the optimizer runs with the last chunk.
// The two products with the batch rows as the outer dimension (a.W and dz.W^T) give a lane four rows and a quarter of the reduction
// index (4x4 register tile, reduce-scatter over the four quarters); the weight gradient gives a thread a 4x2 block of W.
#include <cuda_pipeline.h>
#include "gb_common.cuh"

namespace {

constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int BR = 32;  // rows of a mini-batch chunk: one row per lane

struct FitArgs {
  gb_ffnet net;
  gb::FFImage im;
  gb_fit_hparams hp;
  int apitch[GB_MAX_LAYERS + 1];  // pitch of activation buffer l (l = 0: x staging)
  int aofs[GB_MAX_LAYERS + 1];    // offset of activation buffer l (l >= 1) in smem floats
  int xofs[2], yofs[2], dofs[3];
  int gather_layer, gather_layer2;  // the two forward layers with the fewest tiles (the same layer twice in a one-layer stack): their idle warps issue the cp.async gather of the next chunk
  int d_global;  // how many of the three dz buffers (from the last one) live in the slot's L2-resident state area instead of shared memory
  int ypitch, dpitch;
  int wfloats, smem_floats;
  int n_in, n_out, max_rows;
  long pstride, sstride;
  float* params;
  float* adam_m;
  float* adam_v;
  const gb_job* jobs;
  const float *x, *y;
  const int32_t* perm;
  float *out_loss, *out_acc;
  long long* trace;  // debug (gb_debug_set_fit_trace): cycles of CTA 0 per phase, summed over the fit; NULL in production
};

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}

// keyed bijection on [0, n): 4-round Feistel network on the enclosing power of four, cycle-walked into range
__device__ __forceinline__ uint32_t permute_index(uint32_t i, uint32_t n, uint32_t key) {
  if (n <= 2) return (n == 2) ? (i ^ (key & 1u)) : 0u;
  int bits = 32 - __clz(n - 1);
  if (bits & 1) ++bits;
  const int half = bits >> 1;
  const uint32_t mask = (1u << half) - 1u;
  do {
    uint32_t l = i >> half, r = i & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t t = l ^ (mix32(r * 0x9e3779b9U + key + round * 0x85ebca6bU) & mask);
      l = r;
      r = t;
    }
    i = (l << half) | r;
  } while (i >= n);
  return i;
}

// float -> unsigned with the same ordering (negative values below positive ones)
__device__ __forceinline__ unsigned order_key(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void adam_update(float& w, float g, float& m, float& v, float alpha, float omb1, float omb2,
                                            float eps) {
  m += (g - m) * omb1;
  v += (g * g - v) * omb2;
  float sq;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));  // ~1 ulp, exact 0 at v = 0
  w -= __fdividef(alpha * m, sq + eps);
}

// acc[j][c]: partial sums of rows p + 8 j (j = 0..3) x 4 columns held by lane (p = lane & 7, kq = lane >> 3), to be summed over the four kq.
// Reduce-scatter in two rounds: the lanes 16 apart split rows {0,1} / {2,3}, then the lanes 8 apart split the remaining pair, so lane
// (p, kq) ends with the complete sums of row p + 8 kq (12 shuffles instead of 32 for an all-reduce).
__device__ __forceinline__ void quarter_reduce(const float (&acc)[4][4], int lane, float (&out)[4]) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  float h[2][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const float keep = hi16 ? acc[2 + jj][c] : acc[jj][c], send = hi16 ? acc[jj][c] : acc[2 + jj][c];
      h[jj][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float keep = hi8 ? h[1][c] : h[0][c], send = hi8 ? h[0][c] : h[1][c];
    out[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
}

// WG = false: the slot's padded weight image lives in shared memory for the whole fit (every 64-tag stack).  WG = true: the image does
// not fit beside the activations (e.g. the 128-tag hourglass, 245 KB) and lives in the slot's L2-resident state area instead; the
// code is the same, the loads become global.
template <bool WG, bool DG>  // DG: some dz buffers live in global memory too (kept apart so that the usual case addresses them as shared memory)
__global__ void __launch_bounds__(THREADS, 1) ffae_fit_kernel(const FitArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_red[3][NWARPS];
  __shared__ float s_alpha[2];  // Adam step size of optimizer step t at [t & 1]: written one step ahead, off the critical path
  __shared__ int s_idx[2][BR];
  __shared__ long long s_phase[2 * GB_MAX_LAYERS + 4];

  const int job_id = blockIdx.x;
  const gb_job job = a.jobs[job_id];
  const int n = job.n_rows;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.net.n_layers, n_in = a.n_in, n_out = a.n_out;
  const bool tracing = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
  long long t_mark = 0;
  if (tracing) {
    for (int i = 0; i < 2 * GB_MAX_LAYERS + 4; ++i) s_phase[i] = 0;
    t_mark = clock64();
  }
  auto stamp = [&](int phase) {  // called by everyone right after a barrier; one thread books the cycles since the previous stamp
    if (tracing) { const long long now = clock64(); s_phase[phase] += now - t_mark; t_mark = now; }
  };
  const int B = a.hp.batch_size;
  float* P = a.params + (long)job.slot * a.pstride;
  float* Mg = a.adam_m + (long)job.slot * a.sstride;
  float* sW = WG ? Mg + 2 * a.wfloats : smem;
  float* Vg = a.adam_v + (long)job.slot * a.sstride;

  // ---- weights -> padded smem image; zero every staging buffer ------------------------------------
  for (int i = tid; i < a.smem_floats; i += THREADS) smem[i] = 0.f;
  if (WG)
    for (int i = tid; i < a.wfloats; i += THREADS) sW[i] = 0.f;  // the padding of the image must read as zero
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int K = a.net.dims[l], N = a.net.dims[l + 1], Np = a.im.np[l];
    const float* Wg = P + a.im.pofs[l];
    float* dst = sW + a.im.wofs[l];
    for (int idx = tid; idx < K * N; idx += THREADS) {
      const int k = idx / N, nn = idx - k * N;
      dst[k * Np + nn] = Wg[idx];
    }
    for (int nn = tid; nn < N; nn += THREADS) sW[a.im.bofs[l] + nn] = Wg[K * N + nn];
  }

  const float* xbase = a.x + job.x_row * (long)n_in;
  const float* ybase = a.y + job.x_row * (long)n_out;
  const int steps = (n + B - 1) / B;
  const uint32_t key_base = mix32((uint32_t)a.hp.seed ^ mix32((uint32_t)(a.hp.seed >> 32) + 0x632be5abU * (uint32_t)(job.slot + 1)));

  auto row_index = [&](int e, int i) -> int {
    if (a.hp.shuffle == 0) return i;
    if (a.hp.shuffle == 2) return a.perm[((long)job_id * a.hp.epochs + e) * a.max_rows + i];
    return (int)permute_index((uint32_t)i, (uint32_t)n, mix32(key_base + (uint32_t)e * 0x9e3779b9U));
  };
  // rows [r_lo, r_hi) of chunk c (32 rows) of mini-batch s of epoch e, by n_warps warps
  auto gather = [&](int buf, int e, int s, int c, int first_warp, int n_warps, int r_lo = 0, int r_hi = BR) {
    const int nb = min(BR, min(B, n - s * B) - c * BR);
    float* xs = smem + a.xofs[buf];
    float* ys = smem + a.yofs[buf];
    for (int r = r_lo + warp - first_warp; r < min(nb, r_hi); r += n_warps) {
      const int src = s_idx[buf][r];
      const float* xr = xbase + (long)src * n_in;
      const float* yr = ybase + (long)src * n_out;
      if ((n_in & 3) == 0) {
        for (int c = lane * 4; c < n_in; c += 128) __pipeline_memcpy_async(xs + r * a.apitch[0] + c, xr + c, 16);
      } else {
        for (int c = lane; c < n_in; c += 32) __pipeline_memcpy_async(xs + r * a.apitch[0] + c, xr + c, 4);
      }
      if ((n_out & 3) == 0) {
        for (int c = lane * 4; c < n_out; c += 128) __pipeline_memcpy_async(ys + r * a.ypitch + c, yr + c, 16);
      } else {
        for (int c = lane; c < n_out; c += 32) __pipeline_memcpy_async(ys + r * a.ypitch + c, yr + c, 4);
      }
    }
    __pipeline_commit();
  };

  const float omb1 = 1.f - a.hp.beta1, omb2 = 1.f - a.hp.beta2, eps = a.hp.eps;
  int t_step = a.hp.step0;
  int cur = 0;
  // the visiting order is resolved one chunk ahead of its gather by the last warp (a row per lane): the keyed permutation costs a
  // few hundred instructions per row, which every warp would otherwise repeat in front of its cp.async
  auto advance = [&](int& e, int& s, int& c) -> bool {  // next chunk in visiting order; false past the last epoch
    const int nch = (min(B, n - s * B) + BR - 1) / BR;
    if (++c == nch) { c = 0; if (++s == steps) { s = 0; ++e; } }
    return e < a.hp.epochs;
  };
  auto stage_indices = [&](int buf, int e, int s, int c) {
    const int nb = min(BR, min(B, n - s * B) - c * BR);
    if (lane < nb) s_idx[buf][lane] = row_index(e, s * B + c * BR + lane);
  };
  auto adam_alpha = [&](int t_int) -> float {  // lr * sqrt(1 - b2^t) / (1 - b1^t)
    const double t = (double)t_int;
    return (float)((double)a.hp.lr * sqrt(1.0 - pow((double)a.hp.beta2, t)) / (1.0 - pow((double)a.hp.beta1, t)));
  };
  if (tid == 0) s_alpha[(t_step + 1) & 1] = adam_alpha(t_step + 1);
  if (warp == NWARPS - 1) {
    int e1 = 0, s1 = 0, c1 = 0;
    stage_indices(0, 0, 0, 0);
    if (advance(e1, s1, c1)) stage_indices(1, e1, s1, c1);
  }
  __syncthreads();
  stamp(2 * L + 2);  // set-up
  gather(0, 0, 0, 0, 0, NWARPS);
  float* Gacc = Mg + a.wfloats;  // gradient sums of a multi-chunk mini-batch (same padded layout as the weights)
  // dz buffer b: shared memory, or -- for stacks whose activations leave no room (256-wide encoders) -- the unused part of the slot's
  // Adam-v state area (the second and third third of it), which stays in L2
  auto dz_buf = [&](int b) -> float* {
    if (!DG) return smem + a.dofs[b];
    return b < 3 - a.d_global ? smem + a.dofs[b] : Vg + a.wfloats + (long)(b - (3 - a.d_global)) * BR * a.dpitch;
  };

  for (int e = 0; e < a.hp.epochs; ++e) {
    float acc_sq = 0.f, acc_reg = 0.f, acc_hit = 0.f;
    for (int s = 0; s < steps; ++s) {
      const int nbt = min(B, n - s * B);          // rows of this mini-batch
      const int nchunks = (nbt + BR - 1) / BR;
      ++t_step;
     for (int c = 0; c < nchunks; ++c) {
      const int nb = min(BR, nbt - c * BR);        // rows of this chunk
      const bool first_chunk = c == 0, last_chunk = c + 1 == nchunks;
      // ---- prefetch the next chunk, then wait for the current one ------------------------
      int ne = e, ns = s, nc = c;
      const bool more = advance(ne, ns, nc);
      const int ne1 = ne, ns1 = ns, nc1 = nc;  // the next chunk: gathered below, by the warps without a tile in the narrowest layer
      __pipeline_wait_prior(0);              // this chunk's rows (requested during the previous chunk) have landed
      __syncthreads();
      stamp(0);
      const bool more2 = more && advance(ne, ns, nc);  // (ne, ns, nc): the chunk after next

      // ---- forward ---------------------------------------------------------------------------
      for (int l = 0; l < L; ++l) {
        const int Kp = a.im.kp[l], Np = a.im.np[l], N = a.net.dims[l + 1], act = a.net.act[l];
        const float* in = (l == 0) ? smem + a.xofs[cur] : smem + a.aofs[l];
        float* out = smem + a.aofs[l + 1];
        const int ip = a.apitch[l], op = a.apitch[l + 1];
        const float* Wl = sW + a.im.wofs[l];
        const float* bl = sW + a.im.bofs[l];
        const float l1c = a.net.l1[l] * (a.hp.l1_div_batch ? 1.f : (float)nbt);
        // cp.async of the next chunk: off the step's critical path (at the head of a chunk it cost 2 k cycles), half of the rows in each of
        // the two narrowest layers, by the warps without a tile there (a row costs its warp ~700 cycles of dependent address work)
        if (more && (l == a.gather_layer || l == a.gather_layer2)) {
          const int busy = min(Np >> 2, NWARPS), first = busy < NWARPS ? busy : 0;
          const bool both = a.gather_layer == a.gather_layer2, second = l == a.gather_layer;
          if (warp >= first) gather(cur ^ 1, ne1, ns1, nc1, first, NWARPS - first, (both || !second) ? 0 : BR / 2, (both || second) ? BR : BR / 2);
        }
        if (l == 0) {  // the last two warps have no tile in the first layer of a 64-tag hourglass (14 tiles): they prepare the next step
          if (warp == NWARPS - 1 && more2) stage_indices(cur, ne, ns, nc);  // read by the gather at the top of the next chunk
          if (warp == NWARPS - 2 && first_chunk && lane == 0) s_alpha[(t_step + 1) & 1] = adam_alpha(t_step + 1);  // read after the loss barrier of step t+1
        }
        // A warp owns 32 rows x 4 output columns; lane = (row group p, K quarter kq): rows p, p+8, p+16, p+24 against every fourth
        // block of four k.  Per block a lane loads 4 + 4 float4 for 64 FMA (a row per lane with the whole K needs 1 + 4 for 16: the
        // shared-memory return path, 128 B/clk, bounded these loops); the four K quarters are summed by a two-round reduce-scatter
        // over the lanes that leaves lane (p, kq) with row p + 8 kq.
        const int p8 = lane & 7, kq = lane >> 3, Kb = Kp >> 2;
        for (int task = warp; task < (Np >> 2); task += NWARPS) {
          const int n0 = task << 2;
          float acc[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
          const float* arow = in + p8 * ip;
          const float* wcol = Wl + n0;
          for (int kb = kq; kb < Kb; kb += 4) {
            const int k = kb << 2;
            float4 av[4], wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const float4*>(arow + 8 * j * ip + k);
#pragma unroll
            for (int t = 0; t < 4; ++t) wv[t] = *reinterpret_cast<const float4*>(wcol + (k + t) * Np);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j][0] = fmaf(av[j].x, wv[0].x, acc[j][0]); acc[j][1] = fmaf(av[j].x, wv[0].y, acc[j][1]); acc[j][2] = fmaf(av[j].x, wv[0].z, acc[j][2]); acc[j][3] = fmaf(av[j].x, wv[0].w, acc[j][3]);
              acc[j][0] = fmaf(av[j].y, wv[1].x, acc[j][0]); acc[j][1] = fmaf(av[j].y, wv[1].y, acc[j][1]); acc[j][2] = fmaf(av[j].y, wv[1].z, acc[j][2]); acc[j][3] = fmaf(av[j].y, wv[1].w, acc[j][3]);
              acc[j][0] = fmaf(av[j].z, wv[2].x, acc[j][0]); acc[j][1] = fmaf(av[j].z, wv[2].y, acc[j][1]); acc[j][2] = fmaf(av[j].z, wv[2].z, acc[j][2]); acc[j][3] = fmaf(av[j].z, wv[2].w, acc[j][3]);
              acc[j][0] = fmaf(av[j].w, wv[3].x, acc[j][0]); acc[j][1] = fmaf(av[j].w, wv[3].y, acc[j][1]); acc[j][2] = fmaf(av[j].w, wv[3].z, acc[j][2]); acc[j][3] = fmaf(av[j].w, wv[3].w, acc[j][3]);
            }
          }
          float s4[4];
          quarter_reduce(acc, lane, s4);
          const int row = p8 + 8 * kq;
          const float4 bv = *reinterpret_cast<const float4*>(bl + n0);
          float4 o;
          o.x = (n0 + 0 < N) ? gb::apply_act(act, s4[0] + bv.x) : 0.f;
          o.y = (n0 + 1 < N) ? gb::apply_act(act, s4[1] + bv.y) : 0.f;
          o.z = (n0 + 2 < N) ? gb::apply_act(act, s4[2] + bv.z) : 0.f;
          o.w = (n0 + 3 < N) ? gb::apply_act(act, s4[3] + bv.w) : 0.f;
          *reinterpret_cast<float4*>(out + row * op + n0) = o;
          if (l1c != 0.f && row < nb) acc_reg += l1c * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(o.w));
        }
        __syncthreads();
        stamp(1 + l);
      }

      // ---- loss, accuracy, dz of the output layer: dz = (dL/dyhat + l1*sign(a)) * act'(a) ---------------
      {
        const float* yh = smem + a.aofs[L];
        const int yp = a.apitch[L];
        const float* yt = smem + a.yofs[cur];
        float* G = dz_buf(0);
        const int NpL = a.im.np[L - 1], actL = a.net.act[L - 1];
        const float cL = a.net.l1[L - 1] / (a.hp.l1_div_batch ? (float)nbt : 1.f);
        const float gscale = 2.f / ((float)nbt * (float)n_out);
        for (int r = warp; r < BR; r += NWARPS) {
          // keras "accuracy" on 2-D float targets: argmax match (binary if width 1); first maximum wins, as np.argmax.  The
          // values are compared as order-preserving integer keys so that the warp-wide maximum is one REDUX.
          unsigned kp = 0u, kt = 0u;
          int bp = 0x7fffffff, bt = 0x7fffffff;
          for (int j = lane; j < NpL; j += 32) {
            float g = 0.f;
            if (r < nb && j < n_out) {
              const float ao = yh[r * yp + j], t = yt[r * a.ypitch + j];
              const float d = ao - t;
              acc_sq += d * d;
              g = gscale * d;
              if (cL != 0.f) g += cL * ((ao > 0.f) ? 1.f : ((ao < 0.f) ? -1.f : 0.f));
              g *= gb::act_grad_from_output(actL, ao);
              const unsigned ka = order_key(ao), kb = order_key(t);
              if (ka > kp) { kp = ka; bp = j; }
              if (kb > kt) { kt = kb; bt = j; }
              if (n_out == 1) acc_hit += ((ao > 0.5f ? 1.f : 0.f) == t) ? 1.f : 0.f;
            }
            G[r * a.dpitch + j] = g;
          }
          if (n_out > 1 && r < nb) {
            const unsigned mp = __reduce_max_sync(0xffffffffu, kp), mt = __reduce_max_sync(0xffffffffu, kt);
            const int ip = __reduce_min_sync(0xffffffffu, kp == mp ? bp : 0x7fffffff);
            const int it = __reduce_min_sync(0xffffffffu, kt == mt ? bt : 0x7fffffff);
            if (lane == 0) acc_hit += (ip == it) ? 1.f : 0.f;
          }
        }
      }
      __syncthreads();
      stamp(L + 1);
      const float alpha = s_alpha[t_step & 1];

      // ---- backward + Adam, pipelined over the layers ------------------------------------------------------
      // dz of layer l lives in D buffer (L-1-l) % 3.  Phase p (one barrier each) runs, on disjoint data,
      //   B(p):   dz_{p-1} = (dz_p . W_p^T + l1*sign(a)) * act'(a)      warps from the top, one 4-column task each
      //   C(p+1): dW = a_in^T . dz, db, Adam in place                    all threads, one 4x2 block of W each
      // B(p) reads W_p while C(p+1) writes W_{p+1}; the third buffer keeps dz_{p+1} alive while B(p) writes dz_{p-1}.
      auto input_grad = [&](int l) {  // B(l), l >= 1
        const int Kp = a.im.kp[l], Np = a.im.np[l], K = a.net.dims[l], actp = a.net.act[l - 1];
        const float* D = dz_buf((L - 1 - l) % 3);
        float* Dn = dz_buf((L - l) % 3);
        const float* Wl = sW + a.im.wofs[l];
        const float* aprev = smem + a.aofs[l];  // output of layer l-1
        const float cp = a.net.l1[l - 1] / (a.hp.l1_div_batch ? (float)nbt : 1.f);
        const int p8 = lane & 7, kq = lane >> 3, Nb = Np >> 2;  // lane = (row group, quarter of the n blocks): as in the forward pass
        for (int task = NWARPS - 1 - warp; task < (Kp >> 2); task += NWARPS) {
          const int k0 = task << 2;
          float acc[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
          const float* drow = D + p8 * a.dpitch;
          const float* wrow = Wl + k0 * Np;
          for (int nb4 = kq; nb4 < Nb; nb4 += 4) {
            const int nn = nb4 << 2;
            float4 dv[4], wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dv[j] = *reinterpret_cast<const float4*>(drow + 8 * j * a.dpitch + nn);
#pragma unroll
            for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(wrow + c * Np + nn);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                acc[j][c] = fmaf(dv[j].x, wv[c].x, acc[j][c]);
                acc[j][c] = fmaf(dv[j].y, wv[c].y, acc[j][c]);
                acc[j][c] = fmaf(dv[j].z, wv[c].z, acc[j][c]);
                acc[j][c] = fmaf(dv[j].w, wv[c].w, acc[j][c]);
              }
          }
          float s4[4];
          quarter_reduce(acc, lane, s4);
          const int row = p8 + 8 * kq;
          const bool live = row < nb;
          const float4 ao = *reinterpret_cast<const float4*>(aprev + row * a.apitch[l] + k0);
          auto dz = [&](float g, float o, int j) -> float {
            if (!live || j >= K) return 0.f;
            if (cp != 0.f) g += cp * ((o > 0.f) ? 1.f : ((o < 0.f) ? -1.f : 0.f));
            return g * gb::act_grad_from_output(actp, o);
          };
          float4 o;
          o.x = dz(s4[0], ao.x, k0 + 0); o.y = dz(s4[1], ao.y, k0 + 1); o.z = dz(s4[2], ao.z, k0 + 2); o.w = dz(s4[3], ao.w, k0 + 3);
          *reinterpret_cast<float4*>(Dn + row * a.dpitch + k0) = o;
        }
      };
      auto weight_step = [&](int l) {  // C(l)
        const int Kp = a.im.kp[l], Np = a.im.np[l];
        const float* D = dz_buf((L - 1 - l) % 3);
        const float* ain = (l == 0) ? smem + a.xofs[cur] : smem + a.aofs[l];
        const int ip = a.apitch[l];
        float* Wl = sW + a.im.wofs[l];
        float* Ml = Mg + a.im.wofs[l];
        float* Vl = Vg + a.im.wofs[l];
        float* Gl = Gacc + a.im.wofs[l];
        const int nhalf = Np >> 1, nblocks = (Kp >> 2) * nhalf;
        for (int bid = tid; bid < nblocks; bid += THREADS) {  // consecutive threads along n: the moments stream coalesced
          const int kb = bid / nhalf, n0 = (bid - kb * nhalf) << 1, k0 = kb << 2;
          const bool adam = nchunks == 1 || last_chunk;
          float2 mq[4], vq[4];  // Adam moments of this block: requested now, consumed after the reduction over the batch rows
          if (adam) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              mq[i] = *reinterpret_cast<const float2*>(Ml + (k0 + i) * Np + n0);
              vq[i] = *reinterpret_cast<const float2*>(Vl + (k0 + i) * Np + n0);
            }
          }
          float2 gs[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) gs[i] = make_float2(0.f, 0.f);
#pragma unroll 4
          for (int r = 0; r < BR; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(ain + r * ip + k0);
            const float2 d = *reinterpret_cast<const float2*>(D + r * a.dpitch + n0);
            gs[0].x = fmaf(av.x, d.x, gs[0].x); gs[0].y = fmaf(av.x, d.y, gs[0].y);
            gs[1].x = fmaf(av.y, d.x, gs[1].x); gs[1].y = fmaf(av.y, d.y, gs[1].y);
            gs[2].x = fmaf(av.z, d.x, gs[2].x); gs[2].y = fmaf(av.z, d.y, gs[2].y);
            gs[3].x = fmaf(av.w, d.x, gs[3].x); gs[3].y = fmaf(av.w, d.y, gs[3].y);
          }
          if (nchunks > 1) {  // multi-chunk mini-batch: sum the chunks' gradients in the L2 scratch image; Adam with the last chunk
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2* gp = reinterpret_cast<float2*>(Gl + (k0 + i) * Np + n0);
              if (!first_chunk) {
                const float2 o = *gp;
                gs[i].x += o.x; gs[i].y += o.y;
              }
              if (!last_chunk) *gp = gs[i];
            }
          }
          if (adam) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int off = (k0 + i) * Np + n0;
              float2 w = *reinterpret_cast<float2*>(Wl + off);
              float2 m = mq[i];
              float2 v = vq[i];
              adam_update(w.x, gs[i].x, m.x, v.x, alpha, omb1, omb2, eps);
              adam_update(w.y, gs[i].y, m.y, v.y, alpha, omb1, omb2, eps);
              *reinterpret_cast<float2*>(Wl + off) = w;
              *reinterpret_cast<float2*>(Ml + off) = m;
              *reinterpret_cast<float2*>(Vl + off) = v;
            }
          }
        }
        for (int j = THREADS - 1 - tid; j < Np; j += THREADS) {
          float g = 0.f;
          for (int r = 0; r < BR; ++r) g += D[r * a.dpitch + j];
          const int off = a.im.bofs[l] + j;
          if (nchunks > 1) {
            if (!first_chunk) g += Gacc[off];
            if (!last_chunk) { Gacc[off] = g; continue; }
          }
          float w = sW[off], m = Mg[off], v = Vg[off];
          adam_update(w, g, m, v, alpha, omb1, omb2, eps);
          sW[off] = w; Mg[off] = m; Vg[off] = v;
        }
      };
      for (int p = L - 1; p >= 0; --p) {
        if (p > 0) input_grad(p);
        if (p + 1 < L) weight_step(p + 1);
        if (p == 0) weight_step(0);
        __syncthreads();
        stamp(L + 2 + (L - 1 - p));
      }
      cur ^= 1;
     }  // chunks
    }
    // ---- epoch statistics (keras History: sample-weighted mean of the per-batch total loss) -------------
    float v0 = acc_sq, v1 = acc_reg, v2 = acc_hit;
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if (lane == 0) { s_red[0][warp] = v0; s_red[1][warp] = v1; s_red[2][warp] = v2; }
    __syncthreads();
    if (tid == 0) {
      float q0 = 0.f, q1 = 0.f, q2 = 0.f;
      for (int w = 0; w < NWARPS; ++w) { q0 += s_red[0][w]; q1 += s_red[1][w]; q2 += s_red[2][w]; }
      a.out_loss[(long)job_id * a.hp.epochs + e] = (q0 / (float)n_out + q1) / (float)n;
      if (a.out_acc) a.out_acc[(long)job_id * a.hp.epochs + e] = q2 / (float)n;
    }
    __syncthreads();
  }

  // ---- trained weights back to the canonical layout ------------------------------------------------------
  for (int l = 0; l < L; ++l) {
    const int K = a.net.dims[l], N = a.net.dims[l + 1], Np = a.im.np[l];
    float* Wg = P + a.im.pofs[l];
    const float* src = sW + a.im.wofs[l];
    for (int idx = tid; idx < K * N; idx += THREADS) {
      const int k = idx / N, nn = idx - k * N;
      Wg[idx] = src[k * Np + nn];
    }
    for (int nn = tid; nn < N; nn += THREADS) Wg[K * N + nn] = sW[a.im.bofs[l] + nn];
  }
  if (tracing) {
    stamp(2 * L + 3);  // epoch statistics + write-back
    for (int i = 0; i < 2 * GB_MAX_LAYERS + 4; ++i) a.trace[i] = s_phase[i];
  }
}

long long* g_fit_trace = nullptr;

int odd_pitch(int width) {
  int p4 = (width + 3) / 4;
  if ((p4 & 1) == 0) ++p4;
  return p4 * 4;
}

}  // namespace

extern "C" {

size_t gb_ffae_fit_state_stride(const gb_ffnet* net) {
  if (gb::validate_ffnet(net) != GB_OK) return 0;
  return 3 * (size_t)gb::round_up(gb::make_ff_image(net, 4).total, 4);  // moments + gradient scratch of multi-chunk mini-batches + weight image of wide stacks
}

// debug aid (not part of the public header): per-phase cycle sums of CTA 0 into a device buffer of 2*GB_MAX_LAYERS+4 int64
// (0 gather wait, 1..L forward layers, L+1 loss, L+2.. backward phases, 2L+2 set-up, 2L+3 tail); NULL switches it off
int gb_debug_set_fit_trace(void* dev_buf) {
  g_fit_trace = static_cast<long long*>(dev_buf);
  return GB_OK;
}

int gb_ffae_fit(const gb_ffnet* net, float* params, float* adam_m, float* adam_v, const gb_job* jobs, int32_t n_jobs,
                int32_t max_rows, const float* x, const float* y, const int32_t* perm, const gb_fit_hparams* hp,
                float* out_loss, float* out_acc, void* stream) {
  int rc = gb::validate_ffnet(net);
  if (rc != GB_OK) return rc;
  GB_REQUIRE(params && adam_m && adam_v && jobs && x && y && hp && out_loss, GB_E_ARG,
             "params/adam_m/adam_v/jobs/x/y/hp/out_loss must be non-NULL");
  GB_REQUIRE(hp->epochs >= 1, GB_E_ARG, "epochs=%d must be >= 1", hp->epochs);
  GB_REQUIRE(hp->batch_size >= 1, GB_E_ARG, "batch_size=%d must be >= 1", hp->batch_size);
  GB_REQUIRE(hp->shuffle >= 0 && hp->shuffle <= 2, GB_E_ARG, "shuffle=%d unknown", hp->shuffle);
  GB_REQUIRE(hp->shuffle != 2 || perm, GB_E_ARG, "shuffle=2 needs perm");
  GB_REQUIRE(gb::aligned16(params) && gb::aligned16(adam_m) && gb::aligned16(adam_v) && gb::aligned16(x) &&
                 gb::aligned16(y),
             GB_E_ALIGN, "params/adam/x/y must be 16-byte aligned");
  if (n_jobs == 0 || max_rows == 0) return GB_OK;

  FitArgs a{};
  a.net = *net;
  a.im = gb::make_ff_image(net, 4);
  a.hp = *hp;
  const int L = net->n_layers;
  a.n_in = net->dims[0];
  a.n_out = net->dims[L];
  a.max_rows = max_rows;
  a.pstride = (long)gb_ffnet_param_stride(net);
  a.sstride = (long)gb_ffae_fit_state_stride(net);
  a.wfloats = gb::round_up(a.im.total, 4);
  a.gather_layer = 0;
  for (int l = 1; l < L; ++l)
    if (a.im.np[l] <= a.im.np[a.gather_layer]) a.gather_layer = l;  // the last of the narrowest layers
  a.gather_layer2 = a.gather_layer;
  for (int l = 0, best = 1 << 30; l < L; ++l)
    if (l != a.gather_layer && a.im.np[l] < best) { best = a.im.np[l]; a.gather_layer2 = l; }  // the narrowest of the others
  size_t smem = 0;
  bool w_global = false;
  // first everything in shared memory; if that does not fit, the weight image in L2; then, one by one, the dz buffers in L2 as well
  for (int pass = 0; pass < 5; ++pass) {
    w_global = pass >= 1;
    a.d_global = pass >= 2 ? pass - 1 : 0;
    int ofs = w_global ? 0 : a.wfloats;
    a.apitch[0] = odd_pitch(a.im.kp[0]);
    for (int l = 1; l <= L; ++l) {
      a.apitch[l] = odd_pitch(a.im.np[l - 1]);
      a.aofs[l] = ofs;
      ofs += BR * a.apitch[l];
    }
    for (int b = 0; b < 2; ++b) { a.xofs[b] = ofs; ofs += BR * a.apitch[0]; }
    a.ypitch = odd_pitch(gb::round_up(a.n_out, 4));
    for (int b = 0; b < 2; ++b) { a.yofs[b] = ofs; ofs += BR * a.ypitch; }
    a.dpitch = odd_pitch(a.im.max_np);
    for (int b = 0; b < 3 - a.d_global; ++b) { a.dofs[b] = ofs; ofs += BR * a.dpitch; }
    a.smem_floats = ofs;
    smem = (size_t)ofs * sizeof(float);
    if (smem <= 227 * 1024 && (long)a.d_global * BR * a.dpitch <= 2L * a.wfloats) break;
  }
  GB_REQUIRE(smem <= 227 * 1024 && (long)a.d_global * BR * a.dpitch <= 2L * a.wfloats, GB_E_SMEM,
             "architecture needs %zu bytes of shared memory for the activations of one mini-batch chunk", smem);
  a.params = params; a.adam_m = adam_m; a.adam_v = adam_v; a.jobs = jobs; a.x = x; a.y = y; a.perm = perm;
  a.out_loss = out_loss; a.out_acc = out_acc;
  a.trace = g_fit_trace;
  auto launch = [&](auto kernel) -> int {
    GB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<n_jobs, THREADS, smem, (cudaStream_t)stream>>>(a);
    return GB_OK;
  };
  if (a.d_global > 0) rc = launch(ffae_fit_kernel<true, true>);
  else if (w_global) rc = launch(ffae_fit_kernel<true, false>);
  else rc = launch(ffae_fit_kernel<false, false>);
  if (rc != GB_OK) return rc;
  GB_CUDA_CHECK(cudaGetLastError());
  return GB_OK;
}

}  // extern "C"
