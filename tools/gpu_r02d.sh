#!/bin/bash
# GPU call r02d: the big-scene build = CTA (4 warps) per environment.  Parity first, then throughput.
O=gpurun_out/r02d
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "scene B parity (stage / trajectory / render / variants / grasp attempt)"
timeout 700 python -m pytest tests/test_scene_b_gpu.py tests/test_scale_parity_gpu.py -m gpu -q -x -s > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
tail -n 25 $O/pytest.log
stamp "scene A suite (unchanged warp-per-env build must still pass)"
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_facade_gpu.py -m gpu -q -x > $O/pytest_a.log 2>&1; echo "exit $?" >> $O/pytest_a.log
tail -n 3 $O/pytest_a.log
stamp "scene B throughput"
GE_VERBOSE=1 timeout 300 python tools/bench_scene_b.py 1024 100 > $O/scene_b_1024.log 2>&1; cat $O/scene_b_1024.log; GE_MAXCON=96 GE_VERBOSE=1 timeout 300 python tools/bench_scene_b.py 1024 100 > $O/scene_b_1024_mc96.log 2>&1; tail -3 $O/scene_b_1024_mc96.log
timeout 300 python tools/bench_scene_b.py 2048 100 > $O/scene_b_2048.log 2>&1; tail -1 $O/scene_b_2048.log
timeout 300 python tools/bench_scene_b.py 4096 50 > $O/scene_b_4096.log 2>&1; tail -1 $O/scene_b_4096.log
stamp "sanitizer on the CTA build (memcheck + synccheck + racecheck has no shared workspace to look at)"
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_target.py 4 3 > $O/memcheck.log 2>&1; grep -E "ERROR SUMMARY" $O/memcheck.log | tail -1
timeout 200 compute-sanitizer --tool synccheck --print-limit 5 python tools/sanitize_target.py 4 3 > $O/synccheck.log 2>&1; grep -E "ERROR SUMMARY" $O/synccheck.log | tail -1
stamp "ncu scene B"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 3 -c 1 -f -o $O/k_run_b python tools/bench_scene_b.py 1024 8 > $O/ncu_b.log 2>&1
stamp "done"
