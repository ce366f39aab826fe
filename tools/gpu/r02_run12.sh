set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_heart.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1
tail -4 gpurun_out/r2l_pytest.log
python tools/time_single.py > gpurun_out/r2l_single.json 2> gpurun_out/r2l_single.err; python -c "
import json
d=json.load(open('gpurun_out/r2l_single.json')); print({k:(float('%.4g'%v) if isinstance(v,float) else v) for k,v in d.items() if k!='env'})"
