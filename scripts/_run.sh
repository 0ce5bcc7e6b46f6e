O=gpurun_out/r3q; mkdir -p $O
for nt in 128 256 512; do
MOBGT_K1_NT=$nt timeout 600 python scripts/kbench.py c2-dense128 --k1 > $O/kbench_$nt.log 2>&1; echo "nt=$nt rc=$?"; grep "^k1 n= 128\|^k1 n=  64\|^k1 n= 256" $O/kbench_$nt.log
done
