#!/bin/bash
# Where does the unmodified reference GPU build stop working on this B200?  (runs on the GPU box)
# Generates k-SAT instances of growing size and runs the stock reference CLI on each.
cd "$(dirname "$0")/.."
mkdir -p build gpurun_out
g++ -O2 -std=c++17 -DCNFGEN_MAIN -o build/cnfgen tools/cnfgen.cpp || exit 1
for spec in "200000 852000 3" "400000 1704000 3" "1000000 4260000 3" "100000 2100000 5"; do
  set -- $spec
  ./build/cnfgen ksat 7 /tmp/probe.cnf $1 $2 $3 > /dev/null
  echo "=== ksat n=$1 m=$2 k=$3"
  timeout 300 ./oracle/_ref/parafrost_gpu /tmp/probe.cnf -no-solve -profilegpu --verbose=2 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -v "^c *$" | grep -i -E "error|fail|terminate|what|memory|arena|Electing|SIGmA|simplif|BVE|free|cap" | head -40
done
