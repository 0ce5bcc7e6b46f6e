O=gpurun_out/r3s2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or attn or k3 or canonical or bucketed or golden or many_items or k7 or loss" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
timeout 300 python scripts/step_kernels.py c2-natural > $O/step_nat.txt 2>&1; grep "k3\|workload" $O/step_nat.txt
for w in c2-natural c4-gowalla256; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_$w.json 2> $O/bench_$w.err; python -c "
import json;d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]);print('$w',d['value'],d['ms_per_step'],d['e2e']['value'])"
done
