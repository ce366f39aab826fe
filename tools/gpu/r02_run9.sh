mkdir -p gpurun_out
for w in 1 4; do EKGSIM_B200_ECG_DEBUG=1 EKGSIM_B200_ECG_WAVES=$w EKGSIM_B200_ECG_STAGGER=0 python tools/time_single.py 256 2>&1 >/dev/null | grep "ecg debug" | sort | uniq -c | sort -rn | head -8; done
