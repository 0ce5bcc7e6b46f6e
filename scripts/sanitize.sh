#!/bin/bash
# compute-sanitizer over small instances of every kernel family (scripts/sanitize_cases.py): memcheck on all of them, racecheck
# (shared-memory hazards: K1's DSMEM column broadcast, the mbarrier / named-barrier protocols of K3, K5, K10) on the
# tensor-core / cluster kernels.  usage (on the GPU box, from the repo root): bash scripts/sanitize.sh [out_dir]
O=${1:-gpurun_out/sanitize}
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py > $O/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.txt
tail -4 $O/sanitizer_memcheck.txt
for fam in ${RACE_FAMS:-k1 k3 k3f k5 k10}; do
  timeout 600 $CS --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py $fam > $O/sanitizer_racecheck_$fam.txt 2>&1
  echo "racecheck $fam rc=$?" | tee -a $O/sanitizer_racecheck_$fam.txt
  tail -3 $O/sanitizer_racecheck_$fam.txt
done
