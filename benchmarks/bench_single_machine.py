"""
BASELINE configs[0]: ONE machine through the sklearn-style API -- DiffBasedAnomalyDetector(KerasAutoEncoder
(feedforward_hourglass)), 8 tags, 10 000 rows: cross_validate (3 folds) + fit + anomaly, defaults (epochs=1) and epochs=100.
The CPU column is the reference's control flow on the NumPy oracle (3 fold fits + final fit + predict + diff.py arithmetic).

    python benchmarks/bench_single_machine.py [--rows 10000] [--tags 8] [--cpu-epochs 1]
"""
import argparse, json, os, sys, time
import numpy as np
import pandas as pd
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10000)
    ap.add_argument("--tags", type=int, default=8)
    ap.add_argument("--cpu-epochs", type=int, default=1, help="epochs of the CPU oracle run (its time scales linearly)")
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from gordo_components_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from gordo_components_b200.machine.model.models import KerasAutoEncoder
    from oracle import anomaly_math as am
    from oracle import keras_math as km

    N, T = a.rows, a.tags
    Xv = np.random.default_rng(0).random((N, T))
    idx = pd.date_range("2019-01-01", periods=N, freq="10min", tz="UTC")
    X = pd.DataFrame(Xv, columns=[f"tag-{i}" for i in range(T)], index=idx)

    def gpu_run(epochs):
        det = DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass", epochs=epochs))
        t0 = time.perf_counter()
        det.cross_validate(X=X, y=X)
        t1 = time.perf_counter()
        det.fit(X, X)
        t2 = time.perf_counter()
        frame = det.anomaly(X, X, frequency=pd.Timedelta("10min"))
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return {"cross_validate_s": t1 - t0, "fit_s": t2 - t1, "anomaly_s": t3 - t2, "total_s": t3 - t0, "frame_shape": list(frame.shape)}

    gpu_run(1)  # warm-up: library load, engine caches
    out = {"workload": f"1 machine x {T} tags x {N} rows: cross_validate(3 folds) + fit + anomaly through the sklearn-style API",
           "gpu_epochs_1": gpu_run(1), "gpu_epochs_100": gpu_run(100)}
    # CPU oracle with the reference's control flow
    spec = km.ff_hourglass_spec(T)
    E = a.cpu_epochs
    t0 = time.perf_counter()
    for tr, te in am.time_series_split(N, 3):
        w, _, _ = km.ff_fit(spec, km.init_ff_weights(spec, np.random.default_rng(1)), Xv[tr], Xv[tr], epochs=E, batch_size=32)
        pred = km.ff_predict(spec, w, Xv[te])
        sc, mn = am.minmax_fit(Xv[tr])
        am.fold_thresholds(Xv[te], pred, sc, mn, 6)
    w, _, _ = km.ff_fit(spec, km.init_ff_weights(spec, np.random.default_rng(2)), Xv, Xv, epochs=E, batch_size=32)
    pred = km.ff_predict(spec, w, Xv)
    sc, mn = am.minmax_fit(Xv)
    am.anomaly_arrays(pred, Xv, sc, mn, np.ones(T), 1.0)
    out[f"cpu_oracle_epochs_{E}_s_1core"] = time.perf_counter() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
