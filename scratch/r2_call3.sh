#!/bin/bash
# Round 2, call 3: elect.sync issue + constant-folded layer tables (STATIC) against the generic instantiation
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run_variant() {  # $1 label, $2 defines
  GB_DEFINES="$2" timeout 300 python gordo_components_b200/csrc/build.py > /dev/null || { echo "$1: build failed"; return 1; }
  GB_DEFINES="$2" timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ffae_infer_score or work_split or registered_factory or jobs_slots or ffae_against_reference_generated_fixture or full_size" > gpurun_out/r2_$1_pytest.log 2>&1
  tail -3 gpurun_out/r2_$1_pytest.log
  grep -q " passed" gpurun_out/r2_$1_pytest.log && ! grep -q "failed\|error" gpurun_out/r2_$1_pytest.log || { echo "$1: parity not green, no bench"; return 1; }
  GB_DEFINES="$2" timeout 300 python bench.py --steps 10 --warmup 3 --secondary ${3:-0} > gpurun_out/r2_$1_bench.json 2> gpurun_out/r2_$1_bench.err
  python -c "import json; d=json.load(open('gpurun_out/r2_$1_bench.json')); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['strong']['value'])"
}
run_variant generic "GB_TC_STATIC=0"
run_variant static "" 1
timeout 600 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2_pytest_all3.log 2>&1; tail -8 gpurun_out/r2_pytest_all3.log
python -c "import json; d=json.load(open('gpurun_out/r2_static_bench.json')); print(json.dumps(d['secondary'], indent=1)[:3000]); print(d['cpu_baseline'])"
