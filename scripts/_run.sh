O=gpurun_out/r3t; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for w in c2-dense128 c2-natural c4-gowalla256; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'])"
done
