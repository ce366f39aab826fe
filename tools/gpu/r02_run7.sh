set -x
mkdir -p gpurun_out
for w in 1 0.75 2 4 8 16 64; do EKGSIM_B200_ECG_WAVES=$w python tools/time_single.py > gpurun_out/r2g_single_w$w.json 2>> gpurun_out/r2g_single.err; python -c "
import json
d=json.load(open('gpurun_out/r2g_single_w$w.json'))
print('waves $w', {k:float('%.4g'%v) for k,v in d.items() if k.endswith('_ms')})"; done
