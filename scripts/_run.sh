O=gpurun_out/r3k5; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "embed or segment or k4 or canonical or golden or in_place or bucketed" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_sub.log
timeout 300 python scripts/step_kernels.py c2-natural > $O/step_nat.txt 2>&1; grep "k4_segsum\|workload" $O/step_nat.txt
timeout 300 python scripts/step_kernels.py c2-dense128 > $O/step_c2.txt 2>&1; grep "k4_segsum\|workload" $O/step_c2.txt
