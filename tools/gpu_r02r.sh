#!/bin/bash
# GPU call r02r: the final bench lines of round 2 (config 5 in a child process per rank)
O=gpurun_out/r02r
mkdir -p $O
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
tail -c 3000 $O/bench.json; echo; grep -v "CUDAEvent\|^frame" $O/bench.err | tail -n 3
