"""Time the fused kernel on the BASELINE workload for a list of variant words (debug knobs)."""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine, fleet
from gordo_components_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
spec = feedforward_hourglass(64)
eng = engine.ff_engine_for(spec)
dev = eng.device
M, R = 1000, 10000
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand((M * R, 64), generator=g, device=dev)
params = fleet.random_glorot_params(eng, M, g)
jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
scale = torch.rand((M, 64), generator=g, device=dev) + 0.5
feat = torch.rand((M, 64), generator=g, device=dev) + 0.5
agg = torch.rand((M,), generator=g, device=dev) + 0.5
out = {}
variants = [int(v, 0) for v in sys.argv[1:]] or [2]
for var in variants:
    for _ in range(3):
        eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=var)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=var)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"variant {var:#x}: {ms:.3f} ms  {M*R/ms/1e6:.3f} G windows/s", flush=True)
