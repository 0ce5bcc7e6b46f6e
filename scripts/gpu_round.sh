#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), per-kernel times, ncu launch list + one `--set full` capture per kernel.
# usage (from the repo root, on the box): bash scripts/gpu_round.sh [tag] [stages]   stages = subset of "tbkcln" (default all)
TAG=${1:-r01}
ST=${2:-tbkcln}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [[ $ST == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
fi
if [[ $ST == *b* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 3000 $O/bench.json
  timeout 600 python bench.py --workload c2-natural --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_natural.json 2> $O/bench_natural.err; echo "bench natural rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json
fi
if [[ $ST == *k* ]]; then
  timeout 900 python scripts/kbench.py c2-dense128 --k1 > $O/kbench.log 2>&1; echo "kbench rc=$?"; cat $O/kbench.log
  timeout 300 python scripts/k5bench.py --dbg --timeline > $O/k5bench.log 2>&1; echo "k5bench rc=$?"; grep -v Warning $O/k5bench.log | head -12
  timeout 300 python scripts/e2e_breakdown.py > $O/e2e_breakdown.log 2>&1; echo "e2e_breakdown rc=$?"
  timeout 600 python scripts/profile_step.py > $O/profile_step.log 2>&1; echo "profile_step rc=$?"
fi
if [[ $ST == *c* ]]; then
  timeout 600 python scripts/c3_preprocess.py > $O/c3_preprocess.json 2> $O/c3_preprocess.err; echo "c3 rc=$?"; cat $O/c3_preprocess.json
fi
if [[ $ST == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-report --no-cuda-graph --loader-workers 0 > $O/launches.log 2>&1; echo "ncu launches rc=$?"
fi
if [[ $ST == *n* ]]; then
  for spec in "k2_bias_fwd_kernel:k2_fwd" "k2_bias_bwd_kernel:k2_bwd" "k3_attn_fwd:k3_fwd" "k3_attn_bwd:k3_bwd" "k1_apsp_kernel:k1" \
              "k4_:k4" "k5_head_kernel:k5"; do
    k=${spec%%:*}; n=${spec##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 4 -o $O/$n -f \
        python scripts/kbench.py c2-dense128 --iters=1 > $O/$n.log 2>&1; echo "ncu $n rc=$?"
  done
fi
ls -la $O
