#!/bin/bash
# GPU call r02i: first run of the tcgen05 weight-gradient kernel and the new first-conv wgrad; scene B with the 32-lane contact
# accumulation; then the whole GPU suite on the new default builds (246-register warp-per-env kernel)
O=gpurun_out/r02i
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "learner tests (tcgen05 wgrad)"
timeout 300 python -m pytest tests/test_qnet_learn.py -m gpu -q -x > $O/pytest_learn.log 2>&1; echo "exit $?" >> $O/pytest_learn.log
grep -v "^  [01]\.\|^$" $O/pytest_learn.log | tail -n 25
stamp "learner timing"
timeout 120 python tools/learn_profile.py 12 4 > $O/learn_plain.log 2>&1; tail -n 4 $O/learn_plain.log
GQ_WGRAD_TC=0 timeout 120 python tools/learn_profile.py 12 3 > $O/learn_plain_cudacore.log 2>&1; tail -n 2 $O/learn_plain_cudacore.log
stamp "scene B bench"
timeout 300 python tools/bench_scene_b.py 1024 100 > $O/scene_b.log 2>&1; tail -n 4 $O/scene_b.log
stamp "full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 14 $O/pytest_gpu.log
stamp "done"
