#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run_variant() {  # $1 label, $2 defines
  GB_DEFINES="$2" timeout 300 python gordo_components_b200/csrc/build.py > /dev/null || { echo "$1: build failed"; return 1; }
  GB_DEFINES="$2" timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ffae_infer_score or work_split or registered_factory or jobs_slots or ffae_against_reference_generated_fixture or full_size" > gpurun_out/r2_$1_pytest.log 2>&1
  tail -3 gpurun_out/r2_$1_pytest.log
  grep -q " passed" gpurun_out/r2_$1_pytest.log && ! grep -q "failed\|error" gpurun_out/r2_$1_pytest.log || { echo "$1: parity not green, no bench"; return 1; }
  GB_DEFINES="$2" timeout 300 python bench.py --steps 10 --warmup 3 --secondary 0 > gpurun_out/r2_$1_bench.json 2> gpurun_out/r2_$1_bench.err
  python -c "import json; d=json.load(open('gpurun_out/r2_$1_bench.json')); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['strong']['value'])"
}
for v in "$@"; do
  label=$(echo "$v" | tr ' =' '__'); [ -z "$label" ] && label=default
  run_variant "$label" "$v"
  if [ "${TRACE:-0}" = "1" ]; then GB_DEFINES="$v" GB_TC_TRACE_FROM=20 timeout 300 python scratch/trace_digest.py 2>&1 | tail -16; fi
done
