O=gpurun_out/r3u; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "k10 or k7 or canonical or in_place or golden or model_parity or loss_and_gradients" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-report > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'])"
timeout 300 python scripts/kbench.py c2-dense128 > $O/kbench.log 2>&1; grep "k10\|lib_ffn" $O/kbench.log
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/sanitize_cases.py k3 > $O/san_k3.txt 2>&1; echo "san rc=$?"; tail -5 $O/san_k3.txt
