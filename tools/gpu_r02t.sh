#!/bin/bash
# GPU call r02t (last of round 2): the headline fell from 4.20 M (r02g) to 3.45 M (r02r) - which change?  Quick benches of four libraries,
# then the full bench line with the default library (warp-per-env build back on the one-lane-per-dof contact accumulation)
O=gpurun_out/r02t
mkdir -p $O
for L in default head noguard v0regs; do
  if [ $L = default ]; then unset GE_LIB; else export GE_LIB=$PWD/exp_libs/libgrasp_engine_$L.so; fi
  timeout 200 python bench.py --steps 4 --warmup 3 --legs '' --e2e-steps 1 --cpu-seconds 0 > $O/quick_$L.json 2> $O/quick_$L.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/quick_$L.json").read().strip().splitlines()[-1])
    print("$L value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$L failed", e)
PY
done
unset GE_LIB
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("FULL value %.0f e2e %.0f cfg3 %s cfg4 %s cfg5 %s qnet %.1f learner %.2f ms" % (d["value"], d["e2e"]["value"], d["configs"]["3"]["value"], d["configs"]["4"]["value"], d["configs"]["5"].get("value"), d["qnet"]["tflops"], d["learner"]["ms_per_update"]))
PY
