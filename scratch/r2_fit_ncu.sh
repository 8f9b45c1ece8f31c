#!/bin/bash
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null || exit 1
cat > /tmp/fit_run.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from gordo_components_b200 import engine, fleet
from benchmarks import secondary as sec
print(sec.fit_share(torch, engine, fleet, machines=148, rows=2048, epochs=2))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffae_fit -s 1 -c 1 -o gpurun_out/prof_fit_r02 python /tmp/fit_run.py > gpurun_out/ncu_fit_r02.log 2>&1; tail -2 gpurun_out/ncu_fit_r02.log; ls -la gpurun_out/prof_fit_r02.ncu-rep
