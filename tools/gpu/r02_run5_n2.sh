set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -p no:cacheprovider -k "two_gpus or devices or sharded" > gpurun_out/r2e_pytest_n2.log 2>&1
tail -8 gpurun_out/r2e_pytest_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
tail -c 1200 gpurun_out/r2e_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_bench_n2.json'))
for k in ('value','n_gpus','pipeline','heart4x','generation'):
    print(k, json.dumps(d.get(k))[:2500])
"
for v in 0 100000 400000; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/bench_heart4x.py --mode separable --sharded-automaton --visits-per-round $v > gpurun_out/r2e_heart_sharded_v$v.json 2>> gpurun_out/r2e_heart.err
cat gpurun_out/r2e_heart_sharded_v$v.json
done
