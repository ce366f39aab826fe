set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
python tools/time_single.py > gpurun_out/r2a_single_default.json 2> gpurun_out/r2a_single.err
for sub in 64 256; do EKGSIM_B200_ECG_SUB=$sub python tools/time_single.py 1 > gpurun_out/r2a_single_sub$sub.json 2>> gpurun_out/r2a_single.err; done
EKGSIM_B200_ECG_SLICES=1 python tools/time_single.py 1 > gpurun_out/r2a_single_noslice.json 2>> gpurun_out/r2a_single.err
EKGSIM_B200_MOMENT_FUSED=0 python tools/time_single.py 1 > gpurun_out/r2a_single_nofuse.json 2>> gpurun_out/r2a_single.err
cat gpurun_out/r2a_single_*.json
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2a_bench.json'))
for k in ('value','ms_per_step','e2e','roofline','single_sim','separable_path','pipeline','heart4x','generation','automaton','parity_max_err_of_peak'):
    print(k, json.dumps(d.get(k))[:900])
"
