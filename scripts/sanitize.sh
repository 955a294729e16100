#!/bin/bash
# compute-sanitizer over every kernel family: memcheck everywhere, racecheck + synccheck on the chain kernels (shared-memory
# hand-offs, named barriers).  Run on the GPU box: bash scripts/sanitize.sh [outdir]
out=${1:-gpurun_out}
cs=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; which=$2; shift 2
  echo "=== $tool $which $*"
  env "$@" timeout 600 $cs --tool $tool --error-exitcode 7 --print-limit 20 python scripts/sanitize_target.py $which 2>&1 | grep -v "^$" | tail -12
  echo "rc=${PIPESTATUS[0]}"
}
{
for w in chain multi fc glm analytic; do run memcheck $w A=1; done
run memcheck stepwise PMP_PERSISTENT=0
for w in chain multi; do run racecheck $w A=1; run synccheck $w A=1; done
run racecheck stepwise PMP_PERSISTENT=0
run synccheck fc A=1
} > $out/sanitize.log 2>&1
grep -c "rc=0" $out/sanitize.log; grep "rc=" $out/sanitize.log | sort | uniq -c
