#!/bin/bash
# Round 2, first GPU call: whole suite (no -x), production bench, then the 128-column layer-0 kernel (v17) and the three-slot
# kernel (v18) swapped in inside the box's own copy (parity before bench, short timeouts), then the host-side benches.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CSRC=gordo_components_b200/csrc
run_variant() {
  timeout 120 python -c "import __graft_entry__ as g; g.build()" || { echo "$1: build failed"; return 1; }
  timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ffae_infer_score or work_split or registered_factory or jobs_slots or reference_generated_fixture or full_size" > gpurun_out/r2_$1_pytest.log 2>&1
  tail -4 gpurun_out/r2_$1_pytest.log
  grep -q " passed" gpurun_out/r2_$1_pytest.log && ! grep -q "failed\|error" gpurun_out/r2_$1_pytest.log || { echo "$1: parity not green, no bench"; return 1; }
  timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_$1_bench.json 2> gpurun_out/r2_$1_bench.err
  python -c "import json; d=json.load(open('gpurun_out/r2_$1_bench.json')); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
}

timeout 600 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2_pytest_all.log 2>&1; tail -15 gpurun_out/r2_pytest_all.log
run_variant prod

cp $CSRC/ffae_infer_tc.cu /tmp/ffae_infer_tc_prod.cu
sed 's#"../gordo_components_b200/csrc/gb_common.cuh"#"gb_common.cuh"#' scratch/ffae_infer_tc_v17_layer0_128col.cu > $CSRC/ffae_infer_tc.cu && run_variant v17
if [ $? -eq 0 ]; then
  sed 's#"../gordo_components_b200/csrc/gb_common.cuh"#"gb_common.cuh"#' scratch/ffae_infer_tc_v18_three_slots.cu > $CSRC/ffae_infer_tc.cu && run_variant v18
fi
cp /tmp/ffae_infer_tc_prod.cu $CSRC/ffae_infer_tc.cu
timeout 120 python -c "import __graft_entry__ as g; g.build()"

timeout 200 python benchmarks/bench_fleet_builder.py --machines 125 --epochs 10 > gpurun_out/r2_fleet_builder.json 2> gpurun_out/r2_fleet_builder.err; tail -1 gpurun_out/r2_fleet_builder.json
timeout 200 python benchmarks/bench_requests.py > gpurun_out/r2_requests.json 2> gpurun_out/r2_requests.err; tail -1 gpurun_out/r2_requests.json
timeout 200 python benchmarks/bench_requests.py --bucket > gpurun_out/r2_requests_bucket.json 2>> gpurun_out/r2_requests.err; tail -1 gpurun_out/r2_requests_bucket.json
