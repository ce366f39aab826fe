set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
tail -25 gpurun_out/r2f_pytest.log
python tools/time_single.py > gpurun_out/r2f_single.json 2> gpurun_out/r2f_single.err; cat gpurun_out/r2f_single.json; tail -3 gpurun_out/r2f_single.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 600 gpurun_out/r2f_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
for k in ('value','ms_per_step','roofline','single_sim','separable_path','fast_path','pipeline','heart4x','parity_max_err_of_peak'):
    print(k, json.dumps(d.get(k))[:1200])
"
