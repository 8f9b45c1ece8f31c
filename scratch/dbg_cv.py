"""Debug: which (machine, fold) does each fleet cv_moments entry correspond to?"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from gordo_components_b200 import engine, fleet
from oracle import keras_math as km
from test_gpu_builder import _series, numpy_moments

M, N, T, K = 3, 400, 6, 3
spec = km.ff_hourglass_spec(T)
eng = engine.FFEngine(spec.dims, spec.acts, spec.l1)
frames = [_series(N, T, s) for s in range(M)]
x = torch.from_numpy(np.concatenate([f.values for f in frames])).to(eng.device)
fb = fleet.build_fleet(eng, x, x, rows=N, epochs=2, n_splits=K, seed=3)
torch.cuda.synchronize()
test = N // (K + 1)
mom = fb.cv_moments.cpu().numpy()
cand = {}
for m in range(M):
    for k in range(K):
        for kp in range(K):  # params of fold kp on test block of fold k
            start = N - (K - k) * test
            jobs = engine.jobs_to_device(engine.make_jobs([0], [test], [m * N + start], [0]), eng.device)
            pred = eng.infer_score(fb.fold_params[m, kp:kp + 1].contiguous(), jobs, 1, test, x, out_rows=test)["model-output"].cpu().numpy()
            cand[(m, k, kp)] = numpy_moments(pred, frames[m].values[start:start + test])
        jobs = engine.jobs_to_device(engine.make_jobs([0], [test], [m * N + start], [0]), eng.device)
        pred = eng.infer_score(fb.params[m:m + 1].contiguous(), jobs, 1, test, x, out_rows=test)["model-output"].cpu().numpy()
        cand[(m, k, "final")] = numpy_moments(pred, frames[m].values[start:start + test])
for m in range(M):
    for k in range(K):
        errs = {key: float(np.max(np.abs(v - mom[m, k]) / (np.abs(mom[m, k]) + 1e-9))) for key, v in cand.items()}
        best = min(errs, key=errs.get)
        own = errs[(m, k, k)]
        per_q = np.max(np.abs(cand[(m, k, k)] - mom[m, k]) / (np.abs(mom[m, k]) + 1e-9), axis=1)
        print((m, k), "best", best, f"{errs[best]:.2e}", "own", f"{own:.2e}", "per-q", np.array2string(per_q, precision=2))
