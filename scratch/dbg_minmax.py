import numpy as np, torch, sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine
dev = engine.cuda_device()
g = np.load('tests/golden/anomaly_plain.npz')
y = g['y'].astype(np.float32); n, T = y.shape
yd = torch.from_numpy(y).to(dev)
starts = [int(g[f'fold{i}_test_start']) for i in range(3)]
for jobs_h in (engine.make_jobs([0,1,2,3], starts+[n], [0,0,0,0]), engine.make_jobs([0], [75], [0]), engine.make_jobs([0,1], [75,150], [0,0])):
    print(jobs_h)
    sc, off = engine.minmax_fit(engine.jobs_to_device(jobs_h, dev), len(jobs_h), n, yd, T, len(jobs_h), dev)
    torch.cuda.synchronize()
    print(sc.cpu().numpy()); print(off.cpu().numpy())
for i in range(3): print('want', g[f'fold{i}_scale'], g[f'fold{i}_min'])
