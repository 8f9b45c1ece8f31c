"""
Fleet API: score (and shard) many machines at once.

``anomaly_many`` is the call a fleet user makes: host arrays in, host arrays out, with the host<->device copies
pipelined against the fused kernel on a few CUDA streams.  ``partition``/``assign_machines``/``gather_summaries`` are
the whole multi-GPU story: machines are independent (the reference runs one Kubernetes pod per machine,
gordo/workflow/workflow_generator/resources/argo-workflow.yml.template:1544-1557), so ranks own disjoint contiguous
blocks of machines and the only communication is the broadcast of the assignment and the gather of per-machine
summaries -- there is no exchange step inside fit, predict or anomaly.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import engine

PER_TAG = ("model-output", "tag-anomaly-scaled", "tag-anomaly-unscaled", "anomaly-confidence")
PER_ROW = ("total-anomaly-scaled", "total-anomaly-unscaled", "total-anomaly-confidence")


# ------------------------------------------------------------------------------------------------ sharding
def partition(n_machines: int, world: int) -> List[range]:
    """Contiguous, balanced blocks: the first (n % world) ranks get one extra machine."""
    base, extra = divmod(n_machines, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append(range(start, start + size))
        start += size
    return out


def assign_machines(n_machines: int, world: int, rank: int, dist=None) -> np.ndarray:
    """Rank 0 computes the partition and broadcasts it (NCCL/gloo object broadcast); returns this rank's machine ids."""
    if dist is None or world == 1:
        return np.arange(n_machines)
    payload = [[list(r) for r in partition(n_machines, world)] if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return np.asarray(payload[0][rank], dtype=np.int64)


def gather_summaries(local, world: int, dist=None):
    """all_gather of one fixed-size summary tensor per rank (e.g. max total-anomaly-confidence per machine)."""
    if dist is None or world == 1:
        return local
    torch = engine._torch()
    buf = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, local.contiguous())
    return buf


def random_glorot_params(eng: "engine.FFEngine", n_slots: int, generator):
    """Synthetic fleet: glorot-uniform kernels, small uniform biases, generated on the device (bench/test plumbing)."""
    torch = engine._torch()
    params = torch.zeros((n_slots, eng.param_stride), dtype=torch.float32, device=eng.device)
    ofs = 0
    for i, o in zip(eng.dims[:-1], eng.dims[1:]):
        lim = float(np.sqrt(6.0 / (i + o)))
        params[:, ofs:ofs + i * o] = (torch.rand((n_slots, i * o), generator=generator, device=eng.device) * 2 - 1) * lim
        ofs += i * o
        params[:, ofs:ofs + o] = (torch.rand((n_slots, o), generator=generator, device=eng.device) * 2 - 1) * 0.1
        ofs += o
    return params


# ------------------------------------------------------------------------------------------------ host-buffer scoring
class HostPipeline:
    """Pinned host buffers + per-stream device staging for ``anomaly_many``; reusable across calls of the same shape."""

    def __init__(self, eng: "engine.FFEngine", n_machines: int, rows: int, chunk_machines: int = 50, n_streams: int = 3,
                 want: Sequence[str] = PER_TAG + PER_ROW):
        torch = engine._torch()
        self.eng, self.M, self.R = eng, n_machines, rows
        self.chunk = max(1, min(chunk_machines, n_machines))
        self.want = tuple(want)
        dev = eng.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        cr = self.chunk * rows
        self.stage = []
        for _ in self.streams:
            st = {"x": torch.empty((cr, eng.n_in), dtype=torch.float32, device=dev), "y": torch.empty((cr, eng.n_out), dtype=torch.float32, device=dev), "out": {}}
            for k in self.want:
                st["out"][k] = torch.empty((cr, eng.n_out) if k in PER_TAG else (cr,), dtype=torch.float32, device=dev)
            self.stage.append(st)
        total = n_machines * rows
        self.host_out = {k: torch.empty((total, eng.n_out) if k in PER_TAG else (total,), dtype=torch.float32).pin_memory() for k in self.want}
        self.chunks = []
        for c0 in range(0, n_machines, self.chunk):
            m = min(self.chunk, n_machines - c0)
            jobs = engine.make_jobs(np.arange(c0, c0 + m), rows, np.arange(m, dtype=np.int64) * rows)
            self.chunks.append((c0, m, engine.jobs_to_device(jobs, dev)))
        self.h2d_bytes = total * (eng.n_in + eng.n_out) * 4
        self.d2h_bytes = sum(v.numel() * 4 for v in self.host_out.values())

    def run(self, params, x_host, y_host, scale, feat_thr, agg_thr, variant: int = 0) -> Dict[str, "object"]:
        """x_host / y_host: pinned float32 host tensors [M*R, T].  Returns pinned host tensors (valid after the sync below)."""
        torch = engine._torch()
        R = self.R
        for i, (c0, m, jobs) in enumerate(self.chunks):
            s = self.streams[i % len(self.streams)]
            st = self.stage[i % len(self.streams)]
            rows = slice(c0 * R, (c0 + m) * R)
            with torch.cuda.stream(s):
                st["x"][: m * R].copy_(x_host[rows], non_blocking=True)
                st["y"][: m * R].copy_(y_host[rows], non_blocking=True)
                self.eng.infer_score(params, jobs, m, R, st["x"], st["y"], scale, feat_thr, agg_thr, out_rows=self.chunk * R,
                                     want=self.want, variant=variant, out=st["out"])
                for k in self.want:
                    self.host_out[k][rows].copy_(st["out"][k][: m * R], non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return self.host_out


def anomaly_many(eng: "engine.FFEngine", params, x_host, y_host, scale, feat_thr=None, agg_thr=None, rows: Optional[int] = None,
                 pipeline: Optional[HostPipeline] = None, variant: int = 0):
    """
    Fleet form of ``DiffBasedAnomalyDetector.anomaly``: machine m owns rows [m*rows, (m+1)*rows) of the host arrays and
    slot m of ``params`` / ``scale`` / thresholds.  Returns a dict of host arrays named like the anomaly frame's blocks.
    """
    torch = engine._torch()
    xh = x_host if hasattr(x_host, "is_pinned") else torch.from_numpy(np.ascontiguousarray(x_host, dtype=np.float32))
    yh = y_host if hasattr(y_host, "is_pinned") else torch.from_numpy(np.ascontiguousarray(y_host, dtype=np.float32))
    n_machines = params.shape[0]
    rows = rows or xh.shape[0] // n_machines
    want = [k for k in PER_TAG + PER_ROW if not ((feat_thr is None and k == "anomaly-confidence") or (agg_thr is None and k == "total-anomaly-confidence"))]
    pipe = pipeline or HostPipeline(eng, n_machines, rows, want=want)
    if not xh.is_pinned():
        xh = xh.pin_memory()
    if not yh.is_pinned():
        yh = yh.pin_memory()
    return pipe.run(params, xh, yh, scale, feat_thr, agg_thr, variant)


def time_e2e(eng, params, jobs_h, x_dev, y_dev, scale, feat_thr, agg_thr, steps: int = 3, variant: int = 0):
    """End-to-end windows/s of ``anomaly_many``: pinned host inputs, H2D + kernel + D2H of every output inside the timed region."""
    torch = engine._torch()
    M = len(jobs_h)
    R = int(jobs_h["n_rows"][0])
    xh = torch.empty(x_dev.shape, dtype=torch.float32).pin_memory()
    yh = torch.empty(y_dev.shape, dtype=torch.float32).pin_memory()
    xh.copy_(x_dev)
    yh.copy_(y_dev)
    pipe = HostPipeline(eng, M, R)
    anomaly_many(eng, params, xh, yh, scale, feat_thr, agg_thr, rows=R, pipeline=pipe, variant=variant)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        anomaly_many(eng, params, xh, yh, scale, feat_thr, agg_thr, rows=R, pipeline=pipe, variant=variant)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"ms_per_step": dt * 1e3, "h2d_bytes": pipe.h2d_bytes, "d2h_bytes": pipe.d2h_bytes}


# ------------------------------------------------------------------------------------------------ fleet build (train + thresholds)
class FleetBuild:
    """
    Result of ``build_fleet``: everything ``ModelBuilder._build`` (gordo/builder/build_model.py:192-339) produces for one
    machine -- final weights, target scaler, CV thresholds (per fold and final), loss histories -- for all machines at once.
    """

    def __init__(self, eng, n_machines, n_splits, params, scale, offset, feat_thr, agg_thr, loss, acc, fold_loss, fold_feat_thr, fold_agg_thr,
                 fold_params=None, cv_moments=None, in_scale=None, in_offset=None, fold_in_scale=None, fold_in_offset=None, steps_per_epoch=None):
        self.steps_per_epoch = steps_per_epoch                                 # optimizer steps per epoch of the final fit (keras History.params["steps"])
        # float64 scale_ / min_ of the MinMaxScaler in front of the network ([M, T]; per CV fold [M, K, T]); None without one
        self.in_scale, self.in_offset, self.fold_in_scale, self.fold_in_offset = in_scale, in_offset, fold_in_scale, fold_in_offset
        self.eng, self.n_machines, self.n_splits = eng, n_machines, n_splits
        self.fold_params = fold_params                                         # [M, K, stride]: the CV models (cv["estimator"] of the reference)
        self.cv_moments = cv_moments                                           # [M, K, 5, T] float64: gb_cv_moments of every fold's test block
        self.params, self.scale, self.offset = params, scale, offset          # [M, stride], [M, T], [M, T]
        self.feat_thr, self.agg_thr = feat_thr, agg_thr                        # [M, T], [M]  (last fold, diff.py:257-264)
        self.loss, self.acc = loss, acc                                        # [M, epochs]
        self.fold_loss, self.fold_feat_thr, self.fold_agg_thr = fold_loss, fold_feat_thr, fold_agg_thr  # [M, K, ...]

    @staticmethod
    def _fill_minmax(sc, scale, offset, names):
        """Give a MinMaxScaler the fitted attributes sklearn's ``fit`` would have left (feature_range (0, 1))."""
        sc.scale_, sc.min_ = scale, offset
        sc.data_min_ = -offset / scale
        sc.data_range_ = 1.0 / scale
        sc.data_max_ = sc.data_min_ + sc.data_range_
        sc.n_features_in_, sc.n_samples_seen_ = len(scale), 0
        if names is not None and all(isinstance(n, str) for n in names):
            sc.feature_names_in_ = np.asarray(names, dtype=object)
        return sc

    def detector(self, m: int, tags=None, template=None, input_tags=None):
        """
        Materialise machine ``m`` as a ``DiffBasedAnomalyDetector`` (picklable, servable by gordo.server).  ``template``: an
        unfitted detector built from the machine's own definition (same architecture) to fill in, so that ``kind`` and the
        other constructor arguments survive into ``get_params`` / ``into_definition``.
        """
        import pandas as pd
        from sklearn.preprocessing import MinMaxScaler

        from .machine.model.anomaly.diff import DiffBasedAnomalyDetector
        from .machine.model.factories.specs import FFNetSpec
        from .machine.model.models import FittedNet, History, KerasAutoEncoder

        eng = self.eng
        T = eng.n_out
        tags = list(tags) if tags is not None else list(range(T))
        from sklearn.pipeline import Pipeline

        if template is not None:
            est = template.base_estimator
            ae = est.steps[-1][1] if isinstance(est, Pipeline) else est
            if isinstance(est, Pipeline) != (self.in_scale is not None):
                raise ValueError("the template's input scaler and the fleet's do not match")
            if isinstance(est, Pipeline):
                self._fill_minmax(est.steps[0][1], self.in_scale[m].cpu().numpy(), self.in_offset[m].cpu().numpy(), input_tags)
            ae.kwargs.update({"n_features": eng.n_in, "n_features_out": T})
            ae._prepare_model()
            if list(ae.model.spec.dims) != list(eng.dims) or list(ae.model.spec.acts) != list(eng.acts):
                raise ValueError("template architecture differs from the fleet's")
            ae.model.weights = eng.unpack_params(self.params[m : m + 1])[0]
        else:
            ae = KerasAutoEncoder(kind="feedforward_model", n_features=eng.n_in, n_features_out=T)
            spec = FFNetSpec(list(eng.dims), list(eng.acts), list(eng.l1))
            ae.model = FittedNet(spec, eng.unpack_params(self.params[m : m + 1])[0])
        hist = {"loss": [float(v) for v in self.loss[m].cpu().numpy()], "accuracy": [float(v) for v in self.acc[m].cpu().numpy()]}
        ae._history = History(hist, {"verbose": 0, "epochs": len(hist["loss"]), "steps": self.steps_per_epoch}, list(range(len(hist["loss"]))))
        sc = self._fill_minmax(MinMaxScaler(), self.scale[m].cpu().numpy().astype(np.float64), self.offset[m].cpu().numpy().astype(np.float64), None)
        if template is not None:
            det = template
            det.scaler = sc
        else:
            det = DiffBasedAnomalyDetector(base_estimator=ae, scaler=sc)
        det.feature_thresholds_ = pd.Series(self.feat_thr[m].cpu().numpy().astype(np.float64), index=tags, name=f"fold-{self.n_splits - 1}")
        det.aggregate_threshold_ = float(self.agg_thr[m])
        ff = self.fold_feat_thr[m].cpu().numpy().astype(np.float64)
        det.feature_thresholds_per_fold_ = pd.DataFrame(ff, columns=tags, index=[f"fold-{k}" for k in range(self.n_splits)])
        det.aggregate_thresholds_per_fold_ = {f"fold-{k}": float(self.fold_agg_thr[m, k]) for k in range(self.n_splits)}
        det.smooth_feature_thresholds_per_fold_ = pd.DataFrame()
        det.smooth_aggregate_thresholds_per_fold_ = {}
        det.smooth_aggregate_threshold_ = None
        det.smooth_feature_thresholds_ = None
        return det


def dump_fleet(fb: "FleetBuild", root: str, names: Sequence[str], tags: Optional[Sequence[Sequence[str]]] = None,
               metadata: Optional[Dict[str, dict]] = None, info: Optional[dict] = None) -> List[str]:
    """
    Writes every machine of a ``FleetBuild`` in the layout ``gordo.serializer.dump`` produces and ``gordo.serializer.load`` /
    ``load_metadata`` / gordo.server read (gordo/serializer/serializer.py:149-196): ``<root>/<name>/model.pkl`` (the pickled
    detector), ``metadata.json`` (user metadata + the model's own ``get_metadata()`` under ``metadata.build_metadata.model.
    model_meta`` -- where ModelBuilder puts it, gordo/builder/build_model.py:291-321) and, when given, ``info.json``.
    Returns the directories written.
    """
    import json
    import os
    import pickle

    out = []
    for m, name in enumerate(names):
        det = fb.detector(m, tags=None if tags is None else tags[m])
        dest = os.path.join(root, name)
        os.makedirs(dest, exist_ok=True)
        with open(os.path.join(dest, "model.pkl"), "wb") as f:
            pickle.dump(det, f)
        meta = dict((metadata or {}).get(name, {}))
        meta.setdefault("name", name)
        meta.setdefault("metadata", {}).setdefault("build_metadata", {}).setdefault("model", {})["model_meta"] = det.get_metadata()
        with open(os.path.join(dest, "metadata.json"), "w") as f:
            json.dump(meta, f, default=str)
        if info is not None:
            with open(os.path.join(dest, "info.json"), "w") as f:
                json.dump(info, f, default=str)
        out.append(dest)
    return out


def build_fleet(eng: "engine.FFEngine", x, y, rows: int, epochs: int = 1, batch_size: int = 32, n_splits: int = 3, seed: int = 0,
                adam: Optional[Dict[str, float]] = None, shuffle: bool = True, generator=None, input_scaler: bool = False) -> FleetBuild:
    """
    The batched form of ``gordo build`` for one architecture bucket: for every machine the 3-fold TimeSeriesSplit
    cross-validation (fit on each prefix, thresholds from the following test block: diff.py:176-266) and the final fit on
    all rows (build_model.py:257-321) -- ``(n_splits + 1) * n_machines`` fits in ONE gb_ffae_fit launch (one CTA per fit),
    then fold scoring, threshold reduction and scaler statistics, each a single launch.

    x, y: device tensors [n_machines * rows, T]; machine m owns rows [m*rows, (m+1)*rows).

    ``input_scaler``: the network sits behind a MinMaxScaler (``Pipeline([MinMaxScaler(), KerasAutoEncoder])``, the shape of
    gordo's example configs, examples/config.yaml:74-81).  Inside cross validation every fold clone fits that scaler on its
    own training prefix, so every fit slot gets its own scaled copy of its machine's rows -- sklearn's float64 arithmetic on
    the exact column extrema (``gb_minmax_fit`` finds them, ``gb_affine_f64`` applies them), rounded to float32 once.
    """
    torch = engine._torch()
    dev = eng.device
    M, K, N = x.shape[0] // rows, n_splits, rows
    test = N // (K + 1)
    if test == 0:
        raise ValueError("Too many splits for number of samples")
    starts = [N - (K - k) * test for k in range(K)]  # sklearn TimeSeriesSplit: fold k trains on [0, starts[k]), tests the next `test` rows
    g = generator or torch.Generator(device=dev).manual_seed(seed)
    # slots: [0, M) final models, then fold k of machine m at M + k*M + m
    params = random_glorot_params(eng, M * (K + 1), g)
    # Keras initialises biases to zero
    ofs = 0
    for i, o in zip(eng.dims[:-1], eng.dims[1:]):
        ofs += i * o
        params[:, ofs:ofs + o] = 0
        ofs += o
    base = np.arange(M, dtype=np.int64) * N
    fit_slots = np.concatenate([np.arange(M)] + [M + k * M + np.arange(M) for k in range(K)])
    fit_rows = np.concatenate([np.full(M, N)] + [np.full(M, starts[k]) for k in range(K)])
    fit_x = np.concatenate([base] * (K + 1))
    in_scale = in_offset = None
    if input_scaler:
        S = M * (K + 1)
        prefix_jobs = engine.jobs_to_device(engine.make_jobs(fit_slots, fit_rows, fit_x), dev)
        _, _, lo, hi = engine.minmax_fit(prefix_jobs, S, N, x, eng.n_in, S, dev, return_minmax=True)
        lo, span = lo.double(), hi.double() - lo.double()
        span[~(span >= 10 * np.finfo(np.float64).eps)] = 1.0  # sklearn _handle_zeros_in_scale (also catches all-NaN columns)
        in_scale = 1.0 / span
        in_offset = -lo * in_scale
        # slot s works on its own copy of machine (s mod M)'s rows, at rows [s*N, (s+1)*N) of the replicated arrays
        copy_jobs = engine.jobs_to_device(engine.make_jobs(np.arange(S), N, fit_x, np.arange(S, dtype=np.int64) * N), dev)
        x = engine.affine_f64(copy_jobs, S, N, x.double(), in_scale.contiguous(), in_offset.contiguous(), out_rows=S * N)
        y = y.view(M, N, -1).repeat(K + 1, 1, 1).view(S * N, -1)
        base_of = lambda k: (M + k * M + np.arange(M, dtype=np.int64)) * N  # noqa: E731
        fit_x = np.arange(S, dtype=np.int64) * N
    else:
        base_of = lambda k: base  # noqa: E731
    fit_jobs = engine.jobs_to_device(engine.make_jobs(fit_slots, fit_rows, fit_x), dev)
    loss, acc, _ = eng.fit(params, fit_jobs, len(fit_slots), N, x, y, epochs=epochs, batch_size=batch_size, shuffle=shuffle, adam=adam, seed=seed)
    # scalers: final on all rows, fold k on its training prefix (diff.py:173 inside each CV clone)
    scale, offset = eng.minmax_fit(fit_jobs, len(fit_slots), N, y, M * (K + 1))
    # fold scoring on the test blocks: compact output rows [(k*M + m)*test, ...)
    sc_slots = np.concatenate([M + k * M + np.arange(M) for k in range(K)])
    sc_x = np.concatenate([base_of(k) + starts[k] for k in range(K)])
    sc_out = np.arange(K * M, dtype=np.int64) * test
    sc_jobs = engine.jobs_to_device(engine.make_jobs(sc_slots, test, sc_x, sc_out), dev)
    res = eng.infer_score(params, sc_jobs, K * M, test, x, y, scale, out_rows=K * M * test, want=("tag-anomaly-unscaled", "total-anomaly-scaled"))
    feat, agg = eng.thresholds(sc_jobs, K * M, test, res["tag-anomaly-unscaled"], res["total-anomaly-scaled"], M * (K + 1), window=6)
    T = eng.n_out
    # the evaluation metrics of ModelBuilder's cross validation (build_model.py:250-289) reduce to five sums per (fold, tag)
    moments = engine.cv_moments(sc_jobs, K * M, res["model-output"], y, T).view(K, M, 5, T).permute(1, 0, 2, 3).contiguous()
    fold_params = params[M:].view(K, M, -1).permute(1, 0, 2).contiguous()
    fold_feat = feat[M:].view(K, M, T).permute(1, 0, 2).contiguous()
    fold_agg = agg[M:].view(K, M).t().contiguous()
    E = loss.shape[1]
    return FleetBuild(eng, M, K, params[:M].contiguous(), scale[:M].contiguous(), offset[:M].contiguous(), fold_feat[:, K - 1].contiguous(),
                      fold_agg[:, K - 1].contiguous(), loss[:M], acc[:M], loss[M:].view(K, M, E).permute(1, 0, 2), fold_feat, fold_agg,
                      fold_params=fold_params, cv_moments=moments,
                      in_scale=None if in_scale is None else in_scale[:M].contiguous(), in_offset=None if in_offset is None else in_offset[:M].contiguous(),
                      fold_in_scale=None if in_scale is None else in_scale[M:].view(K, M, -1).permute(1, 0, 2).contiguous(),
                      fold_in_offset=None if in_offset is None else in_offset[M:].view(K, M, -1).permute(1, 0, 2).contiguous(),
                      steps_per_epoch=(N + int(batch_size) - 1) // int(batch_size))
