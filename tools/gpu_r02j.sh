#!/bin/bash
# GPU call r02j: k_render with the screen-box culling (timing + ncu), whole GPU suite on the current builds
O=gpurun_out/r02j
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "render timing"
timeout 200 python tools/render_profile.py 4096 > $O/render_plain.log 2>&1; tail -n 3 $O/render_plain.log
stamp "full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 22 $O/pytest_gpu.log
stamp "ncu k_render"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render$ --launch-skip 1 -c 1 -f -o $O/k_render python tools/render_profile.py 4096 > $O/ncu_render.log 2>&1; tail -n 2 $O/ncu_render.log
stamp "done"
