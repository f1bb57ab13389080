#!/bin/bash
# first GPU call of round 2: full GPU test tier, the default bench line, the k_matvec_lin soak and the sanitizers
out=gpurun_out
mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $out/r02a_tests.log 2>&1
tail -5 $out/r02a_tests.log
(time timeout 1200 python bench.py) > $out/r02a_bench.json 2> $out/r02a_bench.err
tail -c 600 $out/r02a_bench.err
timeout 600 python tools/lin_soak.py 1000 512 > $out/r02a_soak.log 2>&1
cat $out/r02a_soak.log | tail -6
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/lin_soak.py 2 64 T8,T4,T8G1,T4G1 > $out/r02a_${tool}_lin.log 2>&1
  grep -c "hazard\|Error" $out/r02a_${tool}_lin.log; tail -3 $out/r02a_${tool}_lin.log
  for k in gemm fused recur; do
    timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py $k > $out/r02a_${tool}_$k.log 2>&1
    tail -2 $out/r02a_${tool}_$k.log
  done
done
ls -la $out | grep r02a_
