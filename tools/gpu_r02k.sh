#!/bin/bash
# GPU call r02k = the final run of round 2: whole GPU suite, smoke(), both bench arms, k_run ncu at the bench launch shape (-> profiles/k_run_traffic.json),
# launch list of the bench, Q-net forward with / without CUDA-graph replay, learner launch list, compute-sanitizer on the final build
O=gpurun_out/r02k
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 20 $O/pytest_gpu.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
stamp "bench reference arm"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference.json; echo
stamp "bench"
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json; echo; tail -n 3 $O/bench.err
stamp "qnet forward: default vs GQ_GRAPH=1"
timeout 120 python tools/qnet_profile.py 64 5 > $O/qnet_plain.log 2>&1; tail -n 2 $O/qnet_plain.log
GQ_GRAPH=1 timeout 120 python tools/qnet_profile.py 64 5 > $O/qnet_graph.log 2>&1; tail -n 2 $O/qnet_graph.log
GQ_GRAPH=1 timeout 200 python -m pytest tests/test_qnet.py -m gpu -q -k "full_forward or greedy" > $O/pytest_qnet_graph.log 2>&1; tail -n 2 $O/pytest_qnet_graph.log
stamp "ncu k_run at the bench launch shape"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_run --launch-skip 8 -c 1 -f -o $O/k_run python bench.py --steps 6 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_k_run.log 2>&1; tail -n 2 $O/ncu_k_run.log
stamp "launch list of the bench"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 3 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/ncu_ll.log 2>&1; tail -n 1 $O/ncu_ll.log
stamp "learner launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/learn_launches.csv python tools/learn_profile.py 12 2 > $O/learn_ll.log 2>&1; tail -n 2 $O/learn_ll.log
stamp "ncu k_render"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render$ --launch-skip 1 -c 1 -f -o $O/k_render python tools/render_profile.py 4096 > $O/ncu_render.log 2>&1; tail -n 2 $O/ncu_render.log
stamp "compute-sanitizer"
SAN_TIMEOUT=300 bash tools/sanitize.sh $O/sanitize > $O/sanitize.log 2>&1; cat $O/sanitize/summary.txt
stamp "done"
