#!/bin/bash
# training kernel evidence for profiles/: per-phase trace, configs[2] share, ncu full capture
cd "$(dirname "$0")/.."
bash scratch/r2_fit.sh > gpurun_out/r02_fit_trace.txt 2>&1
python - > gpurun_out/r02_fit_share.json 2>gpurun_out/r02_fit_share.err <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from gordo_components_b200 import engine, fleet
from benchmarks import secondary as sec
print(json.dumps(sec.fit_share(torch, engine, fleet)))
PY
bash scratch/r2_fit_ncu.sh
