#!/bin/bash
# per-launch durations and SM clock of the step kernels at the full configs[3] share (lookback shortened)
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null || exit 1
cat > /tmp/lstm_run.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from gordo_components_b200 import engine
from benchmarks import secondary as sec
print(sec.lstm_share(torch, engine, machines=32, rows=10000, lookback=6))
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__cycles_elapsed.max --clock-control none -k regex:lstm_tc_step -s 36 -c 12 --csv --log-file gpurun_out/lstm_launches.csv python /tmp/lstm_run.py > gpurun_out/lstm_launches.log 2>&1
tail -3 gpurun_out/lstm_launches.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/lstm_launches.csv')) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
cur = {}
for r in rows[1:]:
    cur.setdefault(r[ix['ID']], {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
for k, v in cur.items(): print(k, v)
PY
