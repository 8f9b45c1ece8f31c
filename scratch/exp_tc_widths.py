"""tcgen05 kernel vs generic kernel for tag counts below 64."""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine, fleet
from gordo_components_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
for T in (20, 32, 40, 48, 56, 64):
    spec = feedforward_hourglass(T)
    eng = engine.ff_engine_for(spec)
    dev = eng.device
    M, R = 1000, 10000
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand((M * R, T), generator=g, device=dev)
    params = fleet.random_glorot_params(eng, M, g)
    jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
    scale = torch.rand((M, T), generator=g, device=dev) + 0.5
    feat = torch.rand((M, T), generator=g, device=dev) + 0.5
    agg = torch.rand((M,), generator=g, device=dev) + 0.5
    out = {}
    for var in (1, 2):
        for _ in range(2):
            eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=var)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.infer_score(params, jobs, M, R, x, x, scale, feat, agg, out=out, variant=var)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        bpw = 4 * T * 6 + 12
        print(f"T={T:3d} variant {var}: {ms:7.3f} ms  {M*R/ms/1e6:6.3f} G windows/s  {M*R*bpw/ms/1e6:6.0f} GB/s ({M*R*bpw/ms/1e6/6575*100:4.1f}% of HBM peak)", flush=True)
    del x, out
    torch.cuda.empty_cache()
