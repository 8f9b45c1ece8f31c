"""
LSTM training throughput (gb_lstm_fit): machines x windows trained per second for the BASELINE configs[3] architecture
(128 tags, lstm_symmetric dims (256,128,64), lookback 144) and a smaller one, with the CPU oracle's BPTT timed beside it.

    python benchmarks/bench_lstm_fit.py [--machines 8] [--rows 400] [--tags 128] [--lookback 144] [--batch 32] [--cpu 1]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--machines", type=int, default=8)
    ap.add_argument("--rows", type=int, default=400)
    ap.add_argument("--tags", type=int, default=128)
    ap.add_argument("--lookback", type=int, default=144)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cpu", type=int, default=1)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from gordo_components_b200 import engine
    from oracle import keras_math as km

    spec = km.lstm_symmetric_spec(a.tags, lookback_window=a.lookback)
    eng = engine.LSTMEngine(spec.n_features, spec.units, spec.acts, spec.n_features_out, spec.out_func, spec.lookback_window)
    dev = eng.device
    M, N = a.machines, a.rows
    nwin = N - a.lookback + 1
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand((M * N, a.tags), generator=g, device=dev)
    w0 = km.init_lstm_weights(spec, np.random.default_rng(0))
    params = eng.pack_params([w0] * M)
    jobs = engine.jobs_to_device(engine.make_jobs(np.arange(M), nwin, np.arange(M, dtype=np.int64) * N), dev)
    eng.fit(params.clone(), jobs, M, min(nwin, a.batch), x, x, epochs=1, batch_size=a.batch, primer=False)  # warm-up: one step
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss, acc, _ = eng.fit(params, jobs, M, nwin, x, x, epochs=1, batch_size=a.batch, primer=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = 1 + (nwin + a.batch - 1) // a.batch
    out = {"workload": f"{M} machines x {a.tags}-tag lstm_symmetric(256,128,64), lookback {a.lookback}, {nwin} windows, batch {a.batch}, 1 epoch",
           "ms": ms, "steps": steps, "ms_per_step": ms / steps, "window_epochs_per_s": M * nwin / (ms * 1e-3),
           "algorithmic_tflops": 3 * spec.flop_per_window * M * nwin / (ms * 1e-3) / 1e12, "loss": float(loss.mean())}
    if a.cpu:
        X = np.random.default_rng(1).random((a.lookback + 3, a.tags)).astype(np.float32)
        t0 = time.perf_counter()
        km.lstm_fit(spec, w0, X, X, epochs=1, batch_size=4)
        dt = time.perf_counter() - t0
        out["cpu_oracle_windows_per_s_1core"] = 5 / dt  # primer (1 window) + one batch of 4
    print(json.dumps(out))


if __name__ == "__main__":
    main()
