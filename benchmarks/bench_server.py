"""
BASELINE configs[4] shape: many concurrent small anomaly requests (benchmarks/test_ml_server.py: 100 rows per POST) against
device-resident models.  Compares one launch per request with the request coalescer (serving.AnomalyCoalescer).

    python benchmarks/bench_server.py [--machines 1000] [--requests 2000] [--rows 100] [--threads 8]
"""
import argparse, json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--machines", type=int, default=1000)
    ap.add_argument("--requests", type=int, default=2000)
    ap.add_argument("--rows", type=int, default=100)
    ap.add_argument("--threads", type=int, default=8)  # gunicorn threads per worker in the reference (gordo/cli/cli.py:288-296)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from gordo_components_b200 import engine, fleet, serving
    from gordo_components_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass

    eng = engine.ff_engine_for(feedforward_hourglass(64))
    dev = eng.device
    M = a.machines
    g = torch.Generator(device=dev).manual_seed(0)
    params = fleet.random_glorot_params(eng, M, g)
    scale = torch.rand((M, 64), generator=g, device=dev) + 0.5
    feat = torch.rand((M, 64), generator=g, device=dev) + 0.5
    agg = torch.rand((M,), generator=g, device=dev) + 0.5
    rng = np.random.default_rng(0)
    reqs = [(int(rng.integers(0, M)), rng.random((a.rows, 64)).astype(np.float32)) for _ in range(a.requests)]

    def per_request(slot, X):
        n = len(X)
        jobs = engine.jobs_to_device(engine.make_jobs([slot], [n], [0]), dev)
        xd = torch.from_numpy(X).to(dev)
        res = eng.infer_score(params, jobs, 1, n, xd, xd, scale, feat, agg)
        return {k: v.cpu().numpy() for k, v in res.items()}

    def drive(fn):
        idx = iter(range(len(reqs)))
        lock = threading.Lock()
        lat = []

        def worker():
            while True:
                with lock:
                    i = next(idx, None)
                if i is None:
                    return
                t0 = time.perf_counter()
                fn(*reqs[i])
                lat.append(time.perf_counter() - t0)

        ts = [threading.Thread(target=worker) for _ in range(a.threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        dt = time.perf_counter() - t0
        return {"windows_per_s": len(reqs) * a.rows / dt, "requests_per_s": len(reqs) / dt, "p50_ms": 1e3 * float(np.median(lat)), "p99_ms": 1e3 * float(np.quantile(lat, 0.99))}

    for _ in range(20):
        per_request(*reqs[0])
    out = {"workload": f"{a.requests} requests x {a.rows} rows x 64 tags over {M} resident machines, {a.threads} client threads",
           "per_request_launch": drive(per_request)}
    co = serving.AnomalyCoalescer(eng, params, scale, feat, agg, max_wait_ms=0.2)
    co.anomaly(*reqs[0], reqs[0][1])
    out["coalesced"] = drive(lambda s, X: co.anomaly(s, X, X))
    out["coalesced"]["launches"] = co.batches
    # the load-test shape: every request in flight at once (benchmarks/test_ml_server.py fires them concurrently)
    t0 = time.perf_counter()
    futs = [co.submit(s, X, X) for s, X in reqs]
    [f.result() for f in futs]
    dt = time.perf_counter() - t0
    out["coalesced_all_in_flight"] = {"windows_per_s": len(reqs) * a.rows / dt, "launches_total": co.batches}
    co.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
