#!/bin/bash
# Round 2 profile artefacts: launch list of the bench command, ncu --set full of the headline kernel, full GPU test log, bench lines
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null
timeout 600 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; python -c "import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_reference.json')); print('reference', d['value'], d['cpu_baseline']['cores'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --secondary 0 --e2e-steps 1 > gpurun_out/r02_launches_bench.log 2>&1; tail -2 gpurun_out/r02_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffae_tc -s 2 -c 1 -o gpurun_out/prof_tc_r02 python bench.py --machines 300 --steps 3 --warmup 1 --secondary 0 --e2e-steps 1 > gpurun_out/ncu_r02.log 2>&1; ls -la gpurun_out/prof_tc_r02.ncu-rep
GB_TC_TRACE_FROM=20 timeout 300 python scratch/trace_digest.py > gpurun_out/r02_trace_digest.txt 2>&1; tail -14 gpurun_out/r02_trace_digest.txt
