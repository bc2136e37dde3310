#!/bin/bash
# compute-sanitizer passes over the warp-per-env kernels (race / failure detection).  Run on a GPU box:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh gpurun_out/sanitize'
# memcheck: out-of-bounds / misaligned accesses (global + shared);  racecheck: shared-memory hazards between the lanes / warps of a CTA
# (the engine separates its lane loops with __syncwarp / __syncthreads only);  synccheck: divergent barriers;  initcheck: reads of
# uninitialised device memory.  Each tool gets its own time box; a tool that hits it is reported as "timeout", not as clean.
O=${1:-gpurun_out/sanitize}
mkdir -p $O
for tool in memcheck synccheck racecheck initcheck; do
  args="12 2"; [ $tool = racecheck ] && args="4 1"   # racecheck slows kernels down ~100x: fewer sub-steps (both scenes keep their workspace in shared memory)
  timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $args > $O/$tool.log 2>&1
  rc=$?
  echo "$tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/$tool.log | tail -1)" | tee -a $O/summary.txt
done
