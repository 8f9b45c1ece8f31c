#!/bin/bash
# Plan for the first GPU call of the next round (run under gpurun; everything happens in the box's own copy of the repo):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scratch/round2_first_call.sh'
# 0. where round 1 ended (incl. the two tests that have not yet run on a GPU), 1. the 128-column layer-0 scheme (v17),
# 2. three tiles in flight (v18, SHORT timeouts: a barrier mistake hangs), 3. the configuration-driven fleet build.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CSRC=gordo_components_b200/csrc
run_variant() {  # $1 = label: parity of the fused kernel first, then the headline bench
  timeout 60 python -c "import __graft_entry__ as g; g.build()" || { echo "$1: build failed"; return 1; }
  timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ffae_infer_score or work_split or registered_factory or jobs_slots or reference_generated_fixture or full_size" 2>&1 | tail -4 | tee gpurun_out/r2_$1_pytest.log
  grep -q " passed" gpurun_out/r2_$1_pytest.log && ! grep -q "failed\|error" gpurun_out/r2_$1_pytest.log || { echo "$1: parity not green, no bench"; return 1; }
  timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_$1_bench.json 2> gpurun_out/r2_$1_bench.err
  python -c "import json; d=json.load(open('gpurun_out/r2_$1_bench.json')); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
}

timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_pytest_all.log
run_variant prod

cp $CSRC/ffae_infer_tc.cu /tmp/ffae_infer_tc_prod.cu
python scratch/make_v17.py && sed 's#"../gordo_components_b200/csrc/gb_common.cuh"#"gb_common.cuh"#' scratch/ffae_infer_tc_v17_layer0_128col.cu > $CSRC/ffae_infer_tc.cu && run_variant v17
if [ $? -eq 0 ]; then
  python scratch/make_v18.py && sed 's#"../gordo_components_b200/csrc/gb_common.cuh"#"gb_common.cuh"#' scratch/ffae_infer_tc_v18_three_slots.cu > $CSRC/ffae_infer_tc.cu && run_variant v18
fi
cp /tmp/ffae_infer_tc_prod.cu $CSRC/ffae_infer_tc.cu
timeout 60 python -c "import __graft_entry__ as g; g.build()"

timeout 200 python benchmarks/bench_fleet_builder.py --machines 125 --epochs 10 > gpurun_out/r2_fleet_builder.json 2> gpurun_out/r2_fleet_builder.err; tail -1 gpurun_out/r2_fleet_builder.json
timeout 200 python benchmarks/bench_fleet_builder.py --machines 125 --epochs 10 --scaled --single 0 > gpurun_out/r2_fleet_builder_scaled.json 2>> gpurun_out/r2_fleet_builder.err; tail -1 gpurun_out/r2_fleet_builder_scaled.json
timeout 200 python benchmarks/bench_requests.py > gpurun_out/r2_requests.json 2> gpurun_out/r2_requests.err; tail -1 gpurun_out/r2_requests.json
timeout 200 python benchmarks/bench_requests.py --bucket > gpurun_out/r2_requests_bucket.json 2>> gpurun_out/r2_requests.err; tail -1 gpurun_out/r2_requests_bucket.json
