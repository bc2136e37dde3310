#!/bin/bash
# Final GPU call of round 1: Q-net glue changes + TMA weight tiles (toggle), full GPU suite under the best configuration, bench, ncu.
O=gpurun_out/r01j
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }

stamp "qnet tests, defaults"
timeout 120 python -m pytest tests/test_qnet.py -m gpu -x -q > $O/qnet_tests_default.log 2>&1; echo "exit $?" >> $O/qnet_tests_default.log
stamp "qnet tests, GQ_TMA=1"
GQ_TMA=1 timeout 90 python -m pytest tests/test_qnet.py -m gpu -x -q > $O/qnet_tests_tma.log 2>&1; echo "exit $?" >> $O/qnet_tests_tma.log
TMA_OK=$(tail -1 $O/qnet_tests_tma.log | grep -c "exit 0")

stamp "qnet sweeps"
sweep() { name=$1; shift; ( env "$@" timeout 60 python tools/qnet_profile.py 64 5 2>&1 | tail -2 | sed "s/^/$name: /" ) >> $O/qnet_sweep.log; }
sweep default GQ_X=0
sweep head_unfused GQ_FUSE_HEAD=0
sweep first_1px GQ_FIRST1=1
if [ "$TMA_OK" = "1" ]; then
  sweep tma GQ_TMA=1
  sweep tma_npw4 GQ_TMA=1 GQ_NPW=4
fi
cat $O/qnet_sweep.log
BEST=$(python - <<'PY'
import re
best, cfg = 0.0, ""
envs = {"default": "", "tma": "GQ_TMA=1", "tma_npw4": "GQ_TMA=1 GQ_NPW=4"}
for line in open("gpurun_out/r01j/qnet_sweep.log"):
    m = re.match(r"(\w+): rep 4: .* ([\d.]+) TFLOP/s", line)
    if m and m.group(1) in envs and float(m.group(2)) > best:
        best, cfg = float(m.group(2)), envs[m.group(1)]
print(cfg)
PY
)
stamp "best configuration: '$BEST'"
echo "$BEST" > $O/best_env.txt

stamp "all GPU tests under the best configuration"
env $BEST timeout 330 python -m pytest tests -m gpu -x -q > $O/all_gpu_tests.log 2>&1; echo "exit $?" >> $O/all_gpu_tests.log
tail -n 3 $O/all_gpu_tests.log

stamp "bench.py (defaults, best configuration)"
env $BEST timeout 200 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -c 1500 $O/bench.json

stamp "ncu: conv capture, qnet launch list, bench launch list"
env $BEST timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 19 -c 4 -f -o $O/conv_tc python tools/qnet_profile.py 64 2 > $O/qnet_full.log 2>&1
env $BEST timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/qnet_launches.csv python tools/qnet_profile.py 64 2 > $O/qnet_ll.log 2>&1
env $BEST timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-seconds 1 --qnet-images 64 --scene-b-envs 0 > $O/bench_ncu.log 2>&1
stamp "done"
for f in qnet_tests_default qnet_tests_tma all_gpu_tests; do echo "== $f"; tail -n 3 $O/$f.log; done
cat $O/timeline.log
