"""Timeline of CTA 0 of the tcgen05 kernel in steady state -> mean latency of every hop of the per-layer chain (slot 0)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import __graft_entry__ as ge
ge.build()
from gordo_components_b200 import engine, fleet, _cabi
from gordo_components_b200.machine.model.factories.feedforward_autoencoder import feedforward_hourglass
spec = feedforward_hourglass(64)
eng = engine.FFEngine(spec.dims, spec.acts, spec.l1)
dev = eng.device
M, R = int(os.environ.get('TR_M', 296)), 128 * int(os.environ.get('TR_TILES', 78))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand((M * R, 64), generator=g, device=dev); y = x.clone()
params = fleet.random_glorot_params(eng, M, g)
jobs = engine.jobs_to_device(engine.uniform_jobs(M, R), dev)
scale = torch.ones((M, 64), device=dev); feat = torch.ones((M, 64), device=dev); agg = torch.ones((M,), device=dev)
out = {}
for _ in range(3): eng.infer_score(params, jobs, M, R, x, y, scale, feat, agg, out=out)
torch.cuda.synchronize()
lib = _cabi.load_library()
SLOTS = lib.gb_debug_trace_slots()
buf = torch.zeros(8 + 4 * SLOTS + 2 * 148, dtype=torch.int64, device=dev)
lib.gb_debug_set_trace(C.c_void_p(buf.data_ptr()), SLOTS)
eng.infer_score(params, jobs, M, R, x, y, scale, feat, agg, out=out)
torch.cuda.synchronize()
lib.gb_debug_set_trace(None, 0)
b = buf.cpu().numpy()
c0, c1, n0, n1 = [int(v) for v in b[4 + 4 * SLOTS: 8 + 4 * SLOTS]]
cta = b[8 + 4 * SLOTS:].reshape(-1, 2).astype(np.float64)
cta = cta[cta[:, 0] > 0]
if len(cta):
    start0, dur = cta[:, 0].min(), cta[:, 1] - cta[:, 0]
    print(f'{len(cta)} CTAs: life min/mean/max {dur.min() / 1e3:.1f} / {dur.mean() / 1e3:.1f} / {dur.max() / 1e3:.1f} us; last end - first start {(cta[:, 1].max() - start0) / 1e3:.1f} us; '
          f'deciles {np.round(np.quantile(dur, [0.1, 0.3, 0.5, 0.7, 0.9]) / 1e3, 1)}')
print(f'kernel: {c1 - c0} cycles in {n1 - n0} ns -> {(c1 - c0) / max(1, n1 - n0):.3f} GHz; tiles per CTA {M * R // 128 / 148:.1f}')
rec = []
for role in range(4):
    n = int(b[role])
    for v in b[4 + role * SLOTS: 4 + role * SLOTS + n]:
        v = int(v) & ((1 << 64) - 1)
        rec.append((v >> 24, role, (v >> 12) & 0xfff, (v >> 8) & 0xf, (v >> 4) & 0xf, v & 0xf))
rec.sort()
ev = {}  # (code, tile, layer, slot) -> clk  (first occurrence)
for clk, role, tile, layer, slot, code in rec:
    ev.setdefault((code, tile, layer, slot), clk)
tiles0 = sorted({t for (c, t, l, s) in ev if c == 1 and s == 0 and l == 0})
print('slot-0 tiles traced:', tiles0[:6], '...', len(tiles0))
L = 7
hop = {k: [] for k in ('issue', 'commit->woke', 'epilogue', 'arrive->ctrl', 'layer')}
tile_time = []
for i, t in enumerate(tiles0):
    ok = all((1, t, l, 0) in ev and (2, t, l, 0) in ev for l in range(L))
    if not ok:
        continue
    for l in range(L):
        hop['issue'].append((l, ev[(2, t, l, 0)] - ev[(1, t, l, 0)]))
        if l + 1 < L and (5, t, l, 0) in ev and (6, t, l, 0) in ev:
            hop['commit->woke'].append((l, ev[(5, t, l, 0)] - ev[(2, t, l, 0)]))
            hop['epilogue'].append((l, ev[(6, t, l, 0)] - ev[(5, t, l, 0)]))
            hop['arrive->ctrl'].append((l, ev[(1, t, l + 1, 0)] - ev[(6, t, l, 0)]))
            hop['layer'].append((l, ev[(1, t, l + 1, 0)] - ev[(1, t, l, 0)]))
    if i + 1 < len(tiles0) and (1, tiles0[i + 1], 0, 0) in ev:
        tile_time.append(ev[(1, tiles0[i + 1], 0, 0)] - ev[(1, t, 0, 0)])
    # output-layer: commit -> out woke (8) -> parked (10) -> stores issued (12) -> done (9)
for k, v in hop.items():
    per = {}
    for l, d in v:
        per.setdefault(l, []).append(d)
    print(f'{k:14s}', ' '.join(f'l{l}:{np.mean(d):6.0f}' for l, d in sorted(per.items())), f' | sum {sum(np.mean(d) for d in per.values()):7.0f}')
print('slot-0 tile period (cycles between layer-0 wakes):', np.mean(tile_time) if tile_time else None, '-> per tile', (np.mean(tile_time) / 2) if tile_time else None)
for code, name in ((8, 'out woke f'), (10, 'parked'), (12, 'stores issued'), (9, 'emit done')):
    ds = [ev[(code, t, L - 1, 0)] - ev[(2, t, L - 1, 0)] for t in tiles0 if (code, t, L - 1, 0) in ev and (2, t, L - 1, 0) in ev]
    if ds:
        print(f'output layer commit -> {name:14s}: {np.mean(ds):7.0f}')
xs = [ev[(1, t, 0, 0)] - ev[(3, t, 0, 0)] for t in tiles0 if (3, t, 0, 0) in ev]
if xs:
    print('x split arrive -> ctrl wake l0:', np.mean(xs))
if os.environ.get('TR_DUMP'):
    t0 = rec[0][0]
    for clk, role, tile, layer, slot, code in rec[:int(os.environ['TR_DUMP'])]:
        print(f"{clk - t0:8d} role={role} tile={tile:3d} l={layer} s={slot} code={code}")
