#!/bin/bash
# last (short) GPU call of round 1: Q-net tests with the corrected head_up2 reference, TMA toggle, bench
O=gpurun_out/r01k
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "qnet tests, defaults"
timeout 100 python -m pytest tests/test_qnet.py -m gpu -q > $O/qnet_tests_default.log 2>&1; echo "exit $?" >> $O/qnet_tests_default.log
stamp "qnet tests, GQ_TMA=1"
GQ_TMA=1 timeout 80 python -m pytest tests/test_qnet.py -m gpu -q > $O/qnet_tests_tma.log 2>&1; echo "exit $?" >> $O/qnet_tests_tma.log
TMA_OK=$(tail -1 $O/qnet_tests_tma.log | grep -c "exit 0")
stamp "sweeps (TMA ok: $TMA_OK)"
sweep() { name=$1; shift; ( env "$@" timeout 50 python tools/qnet_profile.py 64 5 2>&1 | tail -2 | sed "s/^/$name: /" ) >> $O/qnet_sweep.log; }
sweep default GQ_X=0
if [ "$TMA_OK" = "1" ]; then sweep tma GQ_TMA=1; sweep tma_npw4 GQ_TMA=1 GQ_NPW=4; fi
cat $O/qnet_sweep.log
BEST=$(python - <<'PY'
import re
best, cfg = 0.0, ""
envs = {"default": "", "tma": "GQ_TMA=1", "tma_npw4": "GQ_TMA=1 GQ_NPW=4"}
for line in open("gpurun_out/r01k/qnet_sweep.log"):
    m = re.match(r"(\w+): rep 4: .* ([\d.]+) TFLOP/s", line)
    if m and m.group(1) in envs and float(m.group(2)) > best:
        best, cfg = float(m.group(2)), envs[m.group(1)]
print(cfg)
PY
)
echo "$BEST" > $O/best_env.txt
stamp "bench.py, best configuration '$BEST'"
env $BEST timeout 150 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -c 1200 $O/bench.json
stamp "done"
tail -n 3 $O/qnet_tests_default.log $O/qnet_tests_tma.log
