set -x
mkdir -p gpurun_out
for cfg in "1 25000" "1 0" "1 50000" "1 12000" "2 25000" "4 25000"; do set -- $cfg; EKGSIM_B200_ECG_WAVES=$1 EKGSIM_B200_ECG_STAGGER=$2 python tools/time_single.py > gpurun_out/r2h_single_w$1_s$2.json 2>> gpurun_out/r2h_single.err; python -c "
import json
d=json.load(open('gpurun_out/r2h_single_w$1_s$2.json'))
print('waves $1 stagger $2', {k:float('%.4g'%v) for k,v in d.items() if k.endswith('_ms') or k.endswith('peak')})"; done
