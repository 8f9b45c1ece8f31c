#!/bin/bash
# ncu full capture of the six step kernels of one timestep (32 machines x 1400 rows, lookback 6); $1 = GB_DEFINES, $2 = output tag
cd "$(dirname "$0")/.."
GB_DEFINES="${1:-}" python gordo_components_b200/csrc/build.py > /dev/null || exit 1
TAG="${2:-r02}"
cat > /tmp/lstm_run.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from gordo_components_b200 import engine
from benchmarks import secondary as sec
sec.lstm_share(torch, engine, machines=32, rows=1400, lookback=6)
PY
GB_DEFINES="${1:-}" timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_step -s 12 -c 6 -o gpurun_out/prof_lstm_$TAG python /tmp/lstm_run.py > gpurun_out/ncu_lstm_$TAG.log 2>&1; tail -2 gpurun_out/ncu_lstm_$TAG.log; ls -la gpurun_out/prof_lstm_$TAG.ncu-rep
