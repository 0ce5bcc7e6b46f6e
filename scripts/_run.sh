O=gpurun_out/r3c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "in_place or cuda_graph or canonical or bucketed or entry_cli or k6 or layer_norm or linear or ffn or k10 or model_parity or loss_and_gradients" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_sub.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench rc=$?"; cat $O/bench_c2.json | cut -c1-600
timeout 300 python scripts/step_kernels.py c2-dense128 > $O/step_kernels_c2.txt 2>&1; echo rc=$?
