O=gpurun_out/r3m; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "k5 or head or eval" > $O/pytest_k5.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_k5.log
timeout 300 python scripts/k5bench.py > $O/k5bench.log 2>&1; echo "k5bench rc=$?"; grep -v Warn $O/k5bench.log | grep -v "no cluster" | tail -8
