#!/bin/bash
# fit kernel: parity tests, then the configs[2] share with the per-phase cycle trace of CTA 0
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests -q -m gpu -x -k "fit or build or cross or dropin or fleet or early" 2>&1 | tail -${PYTAIL:-5}
cat > /tmp/fit_run.py <<'PY'
import torch, sys, os, ctypes, json
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from gordo_components_b200 import engine, fleet, _cabi
from benchmarks import secondary as sec
lib = _cabi.load_library()
r = sec.fit_share(torch, engine, fleet, machines=148, rows=10000, epochs=4)
print(json.dumps(r))
buf = torch.zeros(36, dtype=torch.int64, device="cuda")
lib.gb_debug_set_fit_trace(ctypes.c_void_p(buf.data_ptr()))
r2 = sec.fit_share(torch, engine, fleet, machines=148, rows=10000, epochs=4)
lib.gb_debug_set_fit_trace(ctypes.c_void_p(0))
steps = 5 * 313   # warm-up epoch + 4 timed epochs write the same buffer: the last fit (4 epochs) wins
t = buf.cpu().tolist(); L = 7
steps = 4 * 313
names = ["gather wait"] + [f"fwd {l}" for l in range(L)] + ["loss"] + [f"bwd phase p={p}" for p in range(L - 1, -1, -1)] + ["set-up", "tail"]
tot = sum(t)
for nme, c in zip(names, t):
    print(f"{nme:18s} {c / steps:9.0f} cycles/step  {100 * c / tot:5.1f}%")
print("total cycles/step", tot / steps, "traced us/step", r2["us_per_optimizer_step"])
PY
python /tmp/fit_run.py 2>&1 | tail -25
