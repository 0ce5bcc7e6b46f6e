O=gpurun_out/r3v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or attn or k3 or canonical or bucketed or golden or cuda_graph or many_items or single_box" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
timeout 300 python scripts/kbench.py c2-dense128 > $O/kbench.log 2>&1; grep "k3_" $O/kbench.log
for w in c2-dense128 c2-natural c4-gowalla256; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
done
