"""Timeline of CTA 0 of the tcgen05 kernel (clock64 stamps) -> per-phase latencies."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from gordo_components_b200 import engine, fleet, _cabi
from oracle import keras_math as km
spec = km.ff_hourglass_spec(64)
eng = engine.FFEngine(spec.dims, spec.acts, spec.l1)
dev = eng.device
M, R = int(os.environ.get('TR_M', 148)), 128 * int(os.environ.get('TR_TILES', 16))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand((M * R, 64), generator=g, device=dev); y = x.clone()
params = fleet.random_glorot_params(eng, M, g)
jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
scale = torch.ones((M, 64), device=dev); feat = torch.ones((M, 64), device=dev); agg = torch.ones((M,), device=dev)
out = {}
for _ in range(3): eng.infer_score(params, jobs, M, R, x, y, scale, feat, agg, out=out)
torch.cuda.synchronize()
SLOTS = 320
buf = torch.zeros(8 + 4 * SLOTS, dtype=torch.int64, device=dev)
lib = _cabi.load_library()
lib.gb_debug_set_trace(C.c_void_p(buf.data_ptr()), SLOTS)
eng.infer_score(params, jobs, M, R, x, y, scale, feat, agg, out=out)
torch.cuda.synchronize()
lib.gb_debug_set_trace(None, 0)
b = buf.cpu().numpy()
c0, c1, n0, n1 = [int(v) for v in b[4 + 4 * SLOTS: 8 + 4 * SLOTS]]
print(f'kernel: {c1 - c0} cycles in {n1 - n0} ns -> {(c1 - c0) / max(1, n1 - n0):.3f} GHz')
rec = []
for role in range(4):
    n = int(b[role])
    for v in b[4 + role * SLOTS: 4 + role * SLOTS + n]:
        v = int(v) & ((1 << 64) - 1)
        rec.append((v >> 24, role, (v >> 12) & 0xfff, (v >> 8) & 0xf, (v >> 4) & 0xf, v & 0xf))
rec.sort()
t0 = rec[0][0]
names = {1: 'ctrl wake(a_ready)', 2: 'ctrl committed', 3: 'epi X arrive', 4: 'epi wait d (hidden)', 5: 'epi woke d (hidden)', 6: 'epi arrive a (hidden)', 7: 'epi wait d (final)', 8: 'epi woke d (final)', 9: 'epi final done', 10: 'out: D read, d_free arrived', 12: 'out: stores issued', 13: 'ITEM START', 14: 'ITEM staged'}
print('events', len(rec))
for clk, role, tile, layer, slot, code in rec[:400]:
    print(f"{clk - t0:8d}  role={role} tile={tile:2d} l={layer} s={slot}  {names.get(code, code)}")
