#!/bin/bash
cd "$(dirname "$0")/.."
GB_DEFINES="${1:-}" python gordo_components_b200/csrc/build.py > /dev/null || exit 1
GB_DEFINES="${1:-}" timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "lstm" 2>&1 | tail -3
GB_DEFINES="${1:-}" timeout -k 10 300 python - <<'PY'
import json, torch, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from gordo_components_b200 import engine
from benchmarks import secondary as sec
import bench
peaks, _ = bench.measured_peaks()
import subprocess, threading
smi = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active', '--format=csv,noheader', '-lms', '100'], stdout=subprocess.PIPE, text=True)
for _ in range(2):
    r = sec.lstm_share(torch, engine, peaks=peaks)
smi.terminate()
lines = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
print('clocks under load (sm MHz, W, reasons):', lines[len(lines) // 2 - 2: len(lines) // 2 + 3])
print(json.dumps({k: r[k] for k in ("ms", "windows_per_s", "algorithmic_tflops", "tensor_pipe_frac")}))
PY
