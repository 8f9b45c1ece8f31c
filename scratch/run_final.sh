mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_full9.log; tail -3 gpurun_out/pytest_full9.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench9_tc.json 2>gpurun_out/bench9_tc.err; python -c "import json; d=json.load(open('gpurun_out/bench9_tc.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/bench9_tc.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench9_ref.json 2>gpurun_out/bench9_ref.err; python -c "import json; d=json.load(open('gpurun_out/bench9_ref.json')); print('ref', d['value'], d['cpu_baseline']['cores'])"
