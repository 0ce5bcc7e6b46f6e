O=gpurun_out/r3s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "many_items" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_sub.log
