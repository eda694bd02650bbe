#!/bin/bash
# compute-sanitizer over profiles/sanitize_run.py (run on the GPU box): memcheck, racecheck (shared-memory hazards),
# synccheck.  Logs -> gpurun_out/sanitizer/.  The global-memory races the kernels rely on are INTENDED and listed in
# profiles/README.md (stale L1 words of the visited bitmap, atomicExch stamps, cross-GPU flag spins); racecheck looks
# at shared memory only, which is where an unintended race would be a bug.
set -u
out=gpurun_out/sanitizer
mkdir -p $out
for tool in memcheck racecheck synccheck; do
  scale=13
  [ $tool = memcheck ] && scale=14
  timeout 900 compute-sanitizer --tool $tool --print-limit 50 --log-file $out/$tool.log python profiles/sanitize_run.py $scale > $out/$tool.stdout 2>&1
  echo "rc=$?" >> $out/$tool.stdout
  tail -3 $out/$tool.log
  tail -2 $out/$tool.stdout
done
