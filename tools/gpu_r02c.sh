#!/bin/bash
# GPU call r02c: dynamic env scheduling in k_run + register-resident group Cholesky + unaligned stage barriers.
O=gpurun_out/r02c
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "parity tests"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_scene_b_gpu.py tests/test_replay_gpu.py tests/test_scale_parity_gpu.py tests/test_facade_gpu.py -m gpu -q -x -s > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -n 12 $O/pytest.log
run() { name=$1; shift; ( env "$@" timeout 120 python bench.py --steps 6 --legs '' --e2e-steps 1 --cpu-seconds 0 2>$O/bench_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms/step %.1f' % d['ms_per_step'], d['env_status_flags'])" ) >> $O/sweep.log 2>&1; tail -1 $O/sweep.log; }
stamp "sweep"
run default GE_X=0
run quota32 GE_QUOTA=32
run quota128 GE_QUOTA=128
run quota256 GE_QUOTA=256
run wpb3 GE_WPB=3
run wpb2 GE_WPB=2
run nosync GE_STAGE_SYNC=0
run wpb3_nosync GE_WPB=3 GE_STAGE_SYNC=0
stamp "synccheck + racecheck"
timeout 150 compute-sanitizer --tool synccheck --print-limit 5 python tools/sanitize_target.py 8 2 > $O/synccheck.log 2>&1; grep -E "ERROR SUMMARY" $O/synccheck.log | tail -1
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_target.py 4 0 > $O/racecheck.log 2>&1; grep -E "RACECHECK SUMMARY" $O/racecheck.log | tail -1
stamp "ncu k_run (steady state: launch 8)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 8 -c 1 -f -o $O/k_run python bench.py --steps 6 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_k_run.log 2>&1
stamp "done"
cat $O/sweep.log
