O=gpurun_out/r3y; mkdir -p $O
for sg in 1 0; do
MOBGT_SPLIT_GRAPH=$sg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$sg bench.py --gpus 8 --steps 30 --warmup 6 --no-cpu-baseline --no-kernel-report > $O/bench_n8_split$sg.json 2> $O/bench_n8_split$sg.err; echo "bench n8 split=$sg rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_n8_split$sg.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
done
