#!/bin/bash
# One GPU call: Q-net conv v2 parity + sweeps, re-synchronised 256-action replay, ncu captures, default bench, remaining GPU tests.
# Every step has its own timeout; logs under gpurun_out/r01i/.
O=gpurun_out/r01i
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }

stamp "qnet tests, defaults (conv v2, fused tail, 4 producer warps)"
timeout 150 python -m pytest tests/test_qnet.py -m gpu -x -q > $O/qnet_tests_default.log 2>&1; echo "exit $?" >> $O/qnet_tests_default.log
stamp "qnet tests, GQ_NPW=8 GQ_CGB=1 GQ_ST=4"
GQ_NPW=8 GQ_CGB=1 GQ_ST=4 timeout 100 python -m pytest tests/test_qnet.py -m gpu -x -q > $O/qnet_tests_npw8.log 2>&1; echo "exit $?" >> $O/qnet_tests_npw8.log
stamp "qnet sweeps"
sweep() { name=$1; shift; ( env "$@" timeout 60 python tools/qnet_profile.py 64 5 2>&1 | tail -2 | sed "s/^/$name: /" ) >> $O/qnet_sweep.log; }
sweep v1_unfused GQ_KERNEL=1 GQ_FUSE_TAIL=0
sweep v2_unfused GQ_FUSE_TAIL=0
sweep v2_default GQ_X=0
sweep v2_npw8 GQ_NPW=8
sweep v2_st4 GQ_ST=4
sweep v2_npw8_cgb_st4 GQ_NPW=8 GQ_CGB=1 GQ_ST=4
sweep v2_npw8_st4 GQ_NPW=8 GQ_ST=4
sweep v2_bn128 GQ_BN=128
cat $O/qnet_sweep.log

stamp "replay test"
timeout 150 python -m pytest tests/test_replay_gpu.py -m gpu -x -q -s > $O/replay.log 2>&1; echo "exit $?" >> $O/replay.log
cp gpurun_out/replay_256_report.json $O/ 2>/dev/null

# best sweep configuration (by the last rep's TFLOP/s) among the v2 ones; used for the captures and the bench below
BEST=$(python - <<'PY'
import re
best, cfg = 0.0, ""
envs = {"v2_default": "", "v2_npw8": "GQ_NPW=8", "v2_cgb": "GQ_CGB=1", "v2_st4": "GQ_ST=4", "v2_npw8_cgb_st4": "GQ_NPW=8 GQ_CGB=1 GQ_ST=4",
        "v2_npw8_st4": "GQ_NPW=8 GQ_ST=4", "v2_bn128": "GQ_BN=128", "v2_bn128_npw8": "GQ_BN=128 GQ_NPW=8"}
for line in open("gpurun_out/r01i/qnet_sweep.log"):
    m = re.match(r"(\w+): rep 4: .* ([\d.]+) TFLOP/s", line)
    if m and m.group(1) in envs and float(m.group(2)) > best:
        best, cfg = float(m.group(2)), envs[m.group(1)]
print(cfg)
PY
)
stamp "best v2 configuration: '$BEST'"
echo "$BEST" > $O/best_env.txt

stamp "ncu: qnet launch list + conv full capture (best configuration)"
env $BEST timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/qnet_launches.csv python tools/qnet_profile.py 64 2 > $O/qnet_ll.log 2>&1
env $BEST timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 19 -c 4 -f -o $O/conv_tc python tools/qnet_profile.py 64 2 > $O/qnet_full.log 2>&1

stamp "bench.py (defaults, best qnet configuration)"
env $BEST timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -c 3000 $O/bench.json

stamp "remaining GPU tests"
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_facade_gpu.py tests/test_scene_b_gpu.py -m gpu -x -q > $O/other_gpu_tests.log 2>&1; echo "exit $?" >> $O/other_gpu_tests.log
stamp "done"
tail -3 $O/qnet_tests_default.log $O/qnet_tests_npw8.log $O/replay.log $O/other_gpu_tests.log
cat $O/timeline.log
