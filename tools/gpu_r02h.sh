#!/bin/bash
# GPU call r02h (1 GPU): ncu of k_render, launch lists of the Q-net forward (current default) and of one learner update
O=gpurun_out/r02h
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "render plain"
timeout 200 python tools/render_profile.py 4096 > $O/render_plain.log 2>&1; tail -n 3 $O/render_plain.log
stamp "ncu k_render"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render$ --launch-skip 1 -c 1 -f -o $O/k_render python tools/render_profile.py 4096 > $O/ncu_render.log 2>&1; tail -n 2 $O/ncu_render.log
stamp "qnet launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/qnet_launches.csv python tools/qnet_profile.py 64 2 > $O/qnet_ll.log 2>&1; tail -n 2 $O/qnet_ll.log
stamp "learner plain + launch list"
timeout 200 python tools/learn_profile.py 12 3 > $O/learn_plain.log 2>&1; tail -n 3 $O/learn_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/learn_launches.csv python tools/learn_profile.py 12 2 > $O/learn_ll.log 2>&1; tail -n 2 $O/learn_ll.log
stamp "done"
