#!/bin/bash
# GPU call r02f: big-scene build after the race fix - full parity, profile; first run of the learner and the training-time observation transform
O=gpurun_out/r02f
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "scene B + scale parity + scene A"
timeout 900 python -m pytest tests/test_scene_b_gpu.py tests/test_scale_parity_gpu.py tests/test_parity_gpu.py tests/test_facade_gpu.py -m gpu -q -x -s > $O/pytest_engine.log 2>&1; echo "exit $?" >> $O/pytest_engine.log
tail -n 8 $O/pytest_engine.log
stamp "learner + obs transform + qnet"
timeout 600 python -m pytest tests/test_qnet_learn.py -m gpu -q -s > $O/pytest_qnet.log 2>&1; echo "exit $?" >> $O/pytest_qnet.log
grep -v "^$" $O/pytest_qnet.log | tail -n 75
stamp "ncu scene B (CTA build)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 3 -c 1 -f -o $O/k_run_b python tools/bench_scene_b.py 1024 8 > $O/ncu_b.log 2>&1
stamp "done"
