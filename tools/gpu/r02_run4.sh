set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
tail -25 gpurun_out/r2d_pytest.log
python tools/bench_fit.py > gpurun_out/r2d_fit.json 2> gpurun_out/r2d_fit.err; cat gpurun_out/r2d_fit.json; tail -3 gpurun_out/r2d_fit.err
python tools/time_single.py > gpurun_out/r2d_single.json 2> gpurun_out/r2d_single.err; cat gpurun_out/r2d_single.json
python bench.py --steps 5 --warmup 3 --no-heart --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
for k in ('value','single_sim','separable_path','pipeline'):
    print(k, json.dumps(d.get(k))[:1500])
"
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum
for w in fit256 separable256 separable1 fit1; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --metrics $M -o gpurun_out/r02_${w}_v2 -f python tools/profile_kernels.py $w > gpurun_out/r2d_ncu_$w.log 2>&1
  tail -2 gpurun_out/r2d_ncu_$w.log
done
