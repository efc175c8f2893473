#!/bin/bash
# Runs every GPU test in its own process (an illegal access in one kernel must not poison the CUDA
# context of the remaining tests) with a per-test timeout; logs under gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" > gpurun_out/tests.txt
pass=0; fail=0
: > gpurun_out/tests.log
while read -r t; do
  out=$(timeout 300 python -m pytest "$t" -x -q 2>&1); rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $t" >> gpurun_out/tests.log
  else fail=$((fail+1)); echo "FAIL($rc) $t" >> gpurun_out/tests.log; echo "$out" | tail -40 >> gpurun_out/tests.log; fi
done < gpurun_out/tests.txt
echo "passed=$pass failed=$fail"
grep -E "^(PASS|FAIL)" gpurun_out/tests.log
