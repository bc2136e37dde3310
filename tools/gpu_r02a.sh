#!/bin/bash
# First GPU call of round 2 (prepared at the end of round 1, when the GPU budget was spent): measure what was built but never run.
#   gpurun --timeout 600 -- 'bash tools/gpu_r02a.sh'
O=gpurun_out/r02a
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
sweep() { name=$1; shift; ( env "$@" timeout 60 python tools/qnet_profile.py 64 5 2>&1 | tail -2 | sed "s/^/$name: /" ) >> $O/qnet_sweep.log; }

stamp "qnet tests, default build (TMA weights)"
timeout 120 python -m pytest tests/test_qnet.py -m gpu -q > $O/qnet_default.log 2>&1; echo "exit $?" >> $O/qnet_default.log
for P in 1 2; do
  stamp "qnet tests, GQ_PERSIST=$P (never run before: a hang is cut by the timeout)"
  GQ_PERSIST=$P timeout 90 python -m pytest tests/test_qnet.py -m gpu -q -x > $O/qnet_persist$P.log 2>&1; echo "exit $?" >> $O/qnet_persist$P.log
done
stamp "sweeps"
sweep default GQ_X=0
for P in 1 2; do
  if tail -1 $O/qnet_persist$P.log | grep -q "exit 0"; then sweep persist$P GQ_PERSIST=$P; fi
done
cat $O/qnet_sweep.log
stamp "ncu: default conv (TMA weights) full capture + qnet launch list"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 19 -c 4 -f -o $O/conv_tc python tools/qnet_profile.py 64 2 > $O/qnet_full.log 2>&1
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/qnet_launches.csv python tools/qnet_profile.py 64 2 > $O/qnet_ll.log 2>&1
for P in 1 2; do
  if grep -q "^persist$P" $O/qnet_sweep.log; then
    GQ_PERSIST=$P timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 19 -c 4 -f -o $O/conv_persist$P python tools/qnet_profile.py 64 2 > $O/qnet_full_p$P.log 2>&1
  fi
done
stamp "façade: record_grasps path + batched data generation (host-tested only so far)"
timeout 120 python tools/generate_data_batched.py --envs 16 --episodes 1 --steps 2 --out $O/Data > $O/generate_data.log 2>&1; echo "exit $?" >> $O/generate_data.log
stamp "done"
for f in qnet_default qnet_persist1 qnet_persist2 generate_data; do echo "== $f"; tail -n 3 $O/$f.log; done
cat $O/timeline.log
