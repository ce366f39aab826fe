set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err
tail -c 800 gpurun_out/r2k_bench_n8.err
python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_n8.json'))
for k in ('value','n_gpus','e2e','pipeline','heart4x','generation'):
    print(k, json.dumps(d.get(k))[:2600])
"
for v in 0 -1 50000 25000; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 tools/bench_heart4x.py --mode separable --sharded-automaton --visits-per-round $v > gpurun_out/r2k_heart_sharded_n8_v$v.json 2>> gpurun_out/r2k_heart.err
python -c "
import json
d=json.load(open('gpurun_out/r2k_heart_sharded_n8_v$v.json'))
print('visits_per_round $v', d['sharded_automaton'], 'replicated ms', d['automaton_ms'], 'sim ms', d['ms_per_sim'])
"
done
