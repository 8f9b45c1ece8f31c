"""Layer-by-layer check of the tcgen05 kernel against the float64 oracle (debug knobs in the variant word)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine
from oracle import keras_math as km

ap = argparse.ArgumentParser()
ap.add_argument('--last', type=int, default=0)   # 0 = full net, l>0: stop after layer l (1-based), output = pre-activation + bias
ap.add_argument('--swap', type=int, default=0)
ap.add_argument('--rows', type=int, default=300)
ap.add_argument('--machines', type=int, default=3)
ap.add_argument('--score', type=int, default=1)
a = ap.parse_args()
T = 64
spec = km.ff_hourglass_spec(T)
rng = np.random.default_rng(0)
M, R = a.machines, a.rows
ws = []
for m in range(M):
    w = km.init_ff_weights(spec, np.random.default_rng(10 + m))
    ws.append([(W, rng.uniform(-0.2, 0.2, b.shape).astype(np.float32)) for W, b in w])
X = (rng.random((M * R, T)) * 2 - 0.5).astype(np.float32)
y = (X + rng.normal(0, 0.05, X.shape)).astype(np.float32)
eng = engine.FFEngine(spec.dims, spec.acts, spec.l1)
dev = eng.device
jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
scale = torch.rand((M, T), device=dev) + 0.5
feat = torch.rand((M, T), device=dev) * 0.2 + 0.05
agg = torch.rand((M,), device=dev) * 0.1 + 0.01
variant = 2 | (a.swap << 8) | (a.last << 16)
params = eng.pack_params(ws)
if a.score:
    res = eng.infer_score(params, jobs, M, R, xd, yd, scale, feat, agg, variant=variant)
else:
    res = eng.infer_score(params, jobs, M, R, xd, variant=variant)
torch.cuda.synchronize()
out = res['model-output'].cpu().numpy()
worst = 0
for m in range(M):
    acts = km.ff_forward(spec, ws[m], X[m*R:(m+1)*R], dtype=np.float64, return_all=True)
    if a.last == 0 or a.last == spec.n_layers:
        want = acts[-1]; n = T
    else:
        l = a.last - 1
        want = acts[l] @ ws[m][l][0].astype(np.float64) + ws[m][l][1]
        n = want.shape[1]
    got = out[m*R:(m+1)*R, :n]
    err = np.abs(got - want)
    worst = max(worst, err.max())
    if m == 0:
        print('machine0 rows 0..2 got', got[:2, :4], 'want', want[:2, :4])
        print('err by row block', [float(err[i:i+128].max()) for i in range(0, R, 128)])
        print('err by col block', [float(err[:, j:j+8].max()) for j in range(0, n, 8)])
print(f'last={a.last} swap={a.swap} max abs err {worst:.3e}')
if a.score and (a.last == 0):
    ref = eng.infer_score(params, jobs, M, R, xd, yd, scale, feat, agg, variant=1)
    torch.cuda.synchronize()
    for k in res:
        d = (res[k] - ref[k]).abs().max().item()
        print(f'  vs fma kernel {k}: max abs diff {d:.3e}')
