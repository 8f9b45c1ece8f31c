#!/bin/bash
cd "$(dirname "$0")/.."
python gordo_components_b200/csrc/build.py > /dev/null
out=gpurun_out/r02_compute_sanitizer.txt
echo "# compute-sanitizer runs over the -m gpu parity tests (round 2, B200)" > $out
K1="infer_score_matches or jobs_slots or work_split or tcgen05_matches or float64_score or quantile or thresholds_edge or smoothing or more_jobs or ffae_fit_matches or wide_symmetric"
echo -e "\n## memcheck: python -m pytest tests -m gpu -k '$K1'" >> $out
timeout -k 10 1200 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -p no:cacheprovider -k "$K1" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 >> $out
K2="generic_architectures or quantile or thresholds_edge or smoothing or float64_score or ffae_fit_matches or wide_symmetric"
echo -e "\n## racecheck (shared-memory hazards; kernels without tcgen05/TMA async proxies): -k '$K2'" >> $out
timeout -k 10 1200 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -p no:cacheprovider -k "$K2" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard" | head -20 >> $out
cat $out
K3="lstm_infer_tcgen05 or ffae_fit_matches or infer_score_matches"
echo -e "\n## synccheck (barrier usage: named barriers of the LSTM epilogue groups, cluster barriers, training kernel): -k '$K3'" >> $out
timeout -k 10 1200 compute-sanitizer --tool synccheck python -m pytest tests -m gpu -q -p no:cacheprovider -k "$K3" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Barrier|error" | head -20 >> $out
tail -6 $out
