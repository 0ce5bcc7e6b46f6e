#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms, every workload), per-kernel times, step breakdowns, ncu launch list + one
# `--set full` capture per kernel, compute-sanitizer.
# usage (from the repo root, on the box): bash scripts/gpu_round.sh [tag] [stages]   stages = subset of "tbkclns" (default all)
TAG=${1:-r01}
ST=${2:-tbkclns}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [[ $ST == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
fi
if [[ $ST == *b* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c2-dense128.json 2> $O/bench_c2-dense128.err; echo "bench rc=$?"; tail -c 3000 $O/bench_c2-dense128.json
  for w in c2-natural c4-gowalla256 c4-gowalla-real c4-dense256 c5-eval; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; cut -c1-160 $O/bench_$w.json
  done
  timeout 600 python bench.py --workload c3-preprocess --steps 5 --warmup 3 > $O/bench_c3-preprocess.json 2> $O/bench_c3-preprocess.err; echo "bench c3 rc=$?"; cut -c1-160 $O/bench_c3-preprocess.json
  timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
fi
if [[ $ST == *k* ]]; then
  timeout 900 python scripts/kbench.py c2-dense128 --k1 > $O/kbench.log 2>&1; echo "kbench rc=$?"; grep -v Warn $O/kbench.log | head -24
  timeout 300 python scripts/k5bench.py --dbg --timeline > $O/k5bench.log 2>&1; echo "k5bench rc=$?"; grep -v Warn $O/k5bench.log | head -12
  for w in c2-dense128 c2-natural c4-gowalla256; do
    timeout 300 python scripts/step_kernels.py $w > $O/step_kernels_$w.txt 2>&1; echo "step_kernels $w rc=$?"
  done
  timeout 300 python scripts/e2e_breakdown.py > $O/e2e_breakdown.log 2>&1; echo "e2e_breakdown rc=$?"
fi
if [[ $ST == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-report --no-cuda-graph --loader-workers 0 > $O/launches.log 2>&1; echo "ncu launches rc=$?"
  python scripts/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; gzip -f $O/launches.csv
fi
if [[ $ST == *n* ]]; then
  # NCU_SPECS="regex:name ..." restricts the captures (e.g. to the kernels that changed since the last pass)
  for spec in ${NCU_SPECS:-k2_bias_fwd_kernel:k2_fwd k2_bias_bwd_kernel:k2_bwd k3_attn_fwd:k3_fwd k3_attn_bwd:k3_bwd k1_apsp_kernel:k1 \
              k4_:k4 k5_head_kernel:k5 k10_gemm_kernel:k10}; do
    k=${spec%%:*}; n=${spec##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 4 -o $O/$n -f \
        python scripts/kbench.py c2-dense128 --iters=1 > $O/$n.log 2>&1; echo "ncu $n rc=$?"
    # summaries are made here and the reports dropped: gpurun copies back at most 64 MiB
    (echo "== $n (ncu --set full --clock-control none; scripts/ncu_summary.py --stalls)"; python scripts/ncu_summary.py $O/$n.ncu-rep --stalls) >> $O/ncu_kernels.txt 2>&1
    if [[ $n != k3_bwd && $n != k5 ]]; then rm -f $O/$n.ncu-rep; fi
  done
  if [[ -z "$NCU_SPECS" ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k3s_attn" -c 4 -o $O/k3s -f \
      python scripts/step_kernels.py c2-natural --steps=1 > $O/k3s.log 2>&1; echo "ncu k3s rc=$?"
  (echo "== k3s: SIMT attention for graphs of <= 16 tokens, c2-natural batch"; python scripts/ncu_summary.py $O/k3s.ncu-rep --stalls) >> $O/ncu_kernels.txt 2>&1
  rm -f $O/k3s.ncu-rep
  fi
fi
if [[ $ST == *s* ]]; then
  bash scripts/sanitize.sh $O > $O/sanitize.log 2>&1; echo "sanitize rc=$?"; grep -h "rc=" $O/sanitize.log
fi
ls -la $O
