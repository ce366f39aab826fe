set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1
tail -6 gpurun_out/r2j_pytest.log
python bench.py --steps 10 --warmup 3 --ref-full > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 300 gpurun_out/r2j_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
for k in ('value','ms_per_step','e2e','single_sim','separable_path','pipeline','heart4x','generation','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:900])
"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; cat gpurun_out/r2j_bench_reference.json | cut -c1-600
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum
for w in direct256 fit256 separable256 direct1 hoisted256 separable1 fit1 automaton; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --metrics $M -o gpurun_out/r02f_$w -f python tools/profile_kernels.py $w > gpurun_out/r2j_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2j_ncu_$w.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-heart > gpurun_out/r2j_bench_under_ncu.json 2> gpurun_out/r2j_bench_under_ncu.err
wc -l gpurun_out/r02_launches.csv
