#!/bin/bash
# GPU call r02g: LPT island balancing + lanes-per-env experiments (64/128/256) on scene B, register-budget experiment on scene A,
# learner whole-step test + first learner bench leg, ncu of the PILED phase of scene B (profile-from-start off)
O=gpurun_out/r02g
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp "scene B parity (default build)"
timeout 600 python -m pytest tests/test_scene_b_gpu.py -m gpu -q -x > $O/pytest_scene_b.log 2>&1; echo "exit $?" >> $O/pytest_scene_b.log
tail -n 4 $O/pytest_scene_b.log
for L in default l256 l64; do
  stamp "bench_scene_b $L"
  if [ $L = default ]; then unset GE_LIB; else export GE_LIB=$PWD/exp_libs/libgrasp_engine_$L.so; fi
  timeout 300 python tools/bench_scene_b.py 1024 100 > $O/scene_b_$L.log 2>&1; tail -n 4 $O/scene_b_$L.log
done
stamp "scene B parity (l256 build)"
GE_LIB=$PWD/exp_libs/libgrasp_engine_l256.so timeout 600 python -m pytest tests/test_scene_b_gpu.py -m gpu -q -x > $O/pytest_scene_b_l256.log 2>&1; echo "exit $?" >> $O/pytest_scene_b_l256.log
tail -n 4 $O/pytest_scene_b_l256.log
unset GE_LIB
stamp "learner tests"
timeout 600 python -m pytest tests/test_qnet_learn.py -m gpu -q -s > $O/pytest_learn.log 2>&1; echo "exit $?" >> $O/pytest_learn.log
grep -v "^  [01]\.\|^$" $O/pytest_learn.log | tail -n 12
stamp "bench scene A default (+ learner leg)"
timeout 600 python bench.py --steps 4 --warmup 3 --legs learn --e2e-steps 1 --cpu-seconds 2 > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.err
python - <<'PY'
import json
for n in ("default",):
    try:
        d = json.loads(open("gpurun_out/r02g/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f e2e %.0f learner %s" % (d["value"], d["e2e"]["value"], d.get("learner")))
    except Exception as e:
        print(n, "failed", e)
PY
stamp "bench scene A v0regs"
GE_LIB=$PWD/exp_libs/libgrasp_engine_v0regs.so timeout 600 python bench.py --steps 4 --warmup 3 --legs '' --e2e-steps 1 --cpu-seconds 2 > $O/bench_v0regs.json 2> $O/bench_v0regs.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02g/bench_v0regs.json").read().strip().splitlines()[-1])
    print("v0regs value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("v0regs failed", e)
PY
stamp "ncu scene B piled phase (default build)"
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_run -c 1 -f -o $O/k_run_b_pile python tools/bench_scene_b.py 1024 40 profile > $O/ncu_b.log 2>&1
tail -n 3 $O/ncu_b.log
stamp "done"
