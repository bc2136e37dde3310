#!/bin/bash
# GPU call r02b: everything written in the first hours of round 2 — full GPU suite (scale parity, free-running replay, post-attempt
# depth, reference scripts, open-loop actuation), bench (both arms), k_run ncu at the bench's launch shape, compute-sanitizer.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02b.sh'
O=gpurun_out/r02b
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 -s > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 30 $O/pytest_gpu.log
cp gpurun_out/replay_256_report.json gpurun_out/free_run_gpu_*.json $O/ 2>/dev/null
stamp "bench reference arm"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
stamp "bench ours"
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -c 3000 $O/bench.json
stamp "ncu k_run at the bench launch shape (4096 envs x 256 iterations)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 4 -c 1 -f -o $O/k_run python bench.py --steps 3 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_k_run.log 2>&1
stamp "launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 3 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_ll.log 2>&1
stamp "sanitizer"
SAN_TIMEOUT=150 bash tools/sanitize.sh $O/sanitize
stamp "done"
cat $O/timeline.log
