"""HBM roofline fraction of the generic fp32 fused kernel for several tag counts (1000 machines, rows scaled to ~2.5 GB of x)."""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine, fleet
from gordo_components_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
VAR = int(os.environ.get('VAR', 1))
for T in ((4, 8, 16) if VAR == 3 else (4, 8, 16, 32, 64, 128)):
    spec = feedforward_hourglass(T)
    eng = engine.ff_engine_for(spec)
    dev = eng.device
    M = 1000
    R = max(1000, int(10000 * 64 / T) // 128 * 128) if T <= 64 else 5000
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand((M * R, T), generator=g, device=dev)
    params = fleet.random_glorot_params(eng, M, g)
    jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
    scale = torch.rand((M, T), generator=g, device=dev) + 0.5
    feat = torch.rand((M, T), generator=g, device=dev) + 0.5
    agg = torch.rand((M,), generator=g, device=dev) + 0.5
    out = {}
    for _ in range(2):
        eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=VAR)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=VAR)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bytes_per_window = 4 * T * 6 + 12
    flop = sum(2 * a * b for a, b in zip(spec.dims[:-1], spec.dims[1:]))
    print(f"T={T:4d} rows/machine={R:6d}: {ms:8.3f} ms  {M*R/ms/1e6:7.3f} G windows/s  {M*R*bytes_per_window/ms/1e6:7.0f} GB/s ({M*R*bytes_per_window/ms/1e6/6575*100:4.1f}% of HBM peak)  {M*R*flop/ms/1e9:6.1f} TFLOP/s", flush=True)
    del x, out
    torch.cuda.empty_cache()
