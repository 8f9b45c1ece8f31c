#!/bin/bash
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null
for mr in "296 9984" "1036 9984" "1000 9984" "1000 10000" "1184 10000" "888 10000"; do
  set -- $mr
  timeout 300 python bench.py --machines $1 --rows $2 --steps 10 --warmup 3 --secondary 0 --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 x $2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
