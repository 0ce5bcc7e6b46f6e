O=gpurun_out/r3r; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or attn or k3 or canonical or bucketed or golden or cuda_graph" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
timeout 600 python scripts/kbench.py c2-dense128 > $O/kbench.log 2>&1; echo "kbench rc=$?"; grep "k3_" $O/kbench.log
for w in c2-dense128 c2-natural c4-dense256; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 $O/bench_$w.json
done
