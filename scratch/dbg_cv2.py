"""Debug: gb_cv_moments on the fleet's own scoring jobs, raw numbers."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from gordo_components_b200 import engine
from test_gpu_builder import _series, numpy_moments
np.set_printoptions(precision=4, linewidth=200)
M, N, T, K = 3, 400, 6, 3
frames = [_series(N, T, s) for s in range(M)]
xh = np.concatenate([f.values for f in frames])
dev = engine.cuda_device()
x = torch.from_numpy(xh).to(dev)
test = N // (K + 1)
starts = [N - (K - k) * test for k in range(K)]
base = np.arange(M, dtype=np.int64) * N
sc_slots = np.concatenate([M + k * M + np.arange(M) for k in range(K)])
sc_x = np.concatenate([base + starts[k] for k in range(K)])
sc_out = np.arange(K * M, dtype=np.int64) * test
jh = engine.make_jobs(sc_slots, test, sc_x, sc_out)
print("jobs", jh)
jobs = engine.jobs_to_device(jh, dev)
yhat = torch.zeros((K * M * test, T), dtype=torch.float32, device=dev)
got = engine.cv_moments(jobs, K * M, yhat, x, T).cpu().numpy()
for j in range(K * M):
    k, m = divmod(j, M)
    blk = xh[m * N + starts[k]: m * N + starts[k] + test]
    want = numpy_moments(np.zeros_like(blk), blk)
    print(j, (m, k), "maxrel", np.max(np.abs(got[j] - want) / (np.abs(want) + 1e-9)))
    if j == 0:
        print("got\n", got[j], "\nwant\n", want)
# same thing with slots = arange (as in the passing unit test)
jh2 = engine.make_jobs(np.arange(K * M), test, sc_x, sc_out)
got2 = engine.cv_moments(engine.jobs_to_device(jh2, dev), K * M, yhat, x, T).cpu().numpy()
print("slots=arange equal:", np.array_equal(got, got2))
