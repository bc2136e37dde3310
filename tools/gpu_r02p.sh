#!/bin/bash
# GPU call r02p: guard build on the deterministic config 5 repro (bench action order), learner kernel tests after the vectorised
# backward kernels, then the final bench line and the k_run capture for roofline.traffic
O=gpurun_out/r02p
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "config 5 repro, bench action order, guard build"
timeout 300 python tools/repro_config5.py 1024 1 0 bench > $O/repro.log 2>&1; echo "exit $?" >> $O/repro.log; grep -v "CUDAEvent\|^frame" $O/repro.log | tail -n 4
stamp "learner tests + timing"
timeout 300 python -m pytest tests/test_qnet_learn.py -m gpu -q > $O/pytest_learn.log 2>&1; echo "exit $?" >> $O/pytest_learn.log; tail -n 3 $O/pytest_learn.log
timeout 120 python tools/learn_profile.py 12 4 > $O/learn_plain.log 2>&1; tail -n 2 $O/learn_plain.log
stamp "engine parity smoke (scene A + B) on the rebuilt library"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_scene_b_gpu.py tests/test_replay_gpu.py -m gpu -q -x > $O/pytest_engine.log 2>&1; echo "exit $?" >> $O/pytest_engine.log; tail -n 3 $O/pytest_engine.log
stamp "bench"
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err; tail -c 2500 $O/bench.json; echo; grep -v "CUDAEvent\|^frame" $O/bench.err | tail -n 4
stamp "ncu k_run at the bench launch shape"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 8 -c 1 -f -o $O/k_run python bench.py --steps 6 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_k_run.log 2>&1; tail -n 2 $O/ncu_k_run.log
stamp "done"
