set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider -k "fit or golden_single or options or devices or heart2x_ecg" > gpurun_out/r2b_pytest.log 2>&1
tail -25 gpurun_out/r2b_pytest.log
python tools/time_single.py > gpurun_out/r2b_single_default.json 2> gpurun_out/r2b_single.err; cat gpurun_out/r2b_single_default.json
python tools/bench_fit.py > gpurun_out/r2b_fit.json 2> gpurun_out/r2b_fit.err; cat gpurun_out/r2b_fit.json; tail -3 gpurun_out/r2b_fit.err
python bench.py --steps 5 --warmup 3 --no-heart > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
for k in ('value','single_sim','separable_path','pipeline','parity_max_err_of_peak'):
    print(k, json.dumps(d.get(k))[:1200])
"
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum
for w in direct256 fit256 separable256 direct1 hoisted256; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --metrics $M -o gpurun_out/r02_$w -f python tools/profile_kernels.py $w > gpurun_out/r2b_ncu_$w.log 2>&1
  tail -2 gpurun_out/r2b_ncu_$w.log
done
ls -la gpurun_out/*.ncu-rep
