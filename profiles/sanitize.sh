#!/bin/bash
# compute-sanitizer over profiles/sanitize_run.py (run on the GPU box): memcheck over everything (both level loops, the
# one-rank peer-memory BFS with its persistent kernels); racecheck (shared-memory hazards) and synccheck over the
# host-driven level loop first, then -- in separate runs, so that a tool limitation cannot void the first log -- over the
# graph-driven loops.  Logs -> gpurun_out/sanitizer/.  The global-memory races the kernels rely on are INTENDED and listed
# in profiles/README.md (stale L1 words of the visited bitmap, atomicExch stamps, cross-GPU flag spins); racecheck looks
# at shared memory only, which is where an unintended race would be a bug.
set -u
out=gpurun_out/sanitizer
mkdir -p $out
run() {  # tool tag scale loops p2p
  timeout 900 compute-sanitizer --tool $1 --print-limit 30 --log-file $out/$1_$2.log python profiles/sanitize_run.py $3 $4 $5 > $out/$1_$2.stdout 2>&1
  echo "rc=$?" >> $out/$1_$2.stdout
  echo "== $1 $2"; tail -2 $out/$1_$2.log; tail -2 $out/$1_$2.stdout
}
run memcheck all 14 both 1
run racecheck hostloop 13 host 0
run synccheck hostloop 13 host 0
run racecheck graphloop 13 graph 0
run synccheck graphloop 13 graph 0
run racecheck p2p 12 host 1
