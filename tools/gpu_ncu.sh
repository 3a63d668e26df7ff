#!/bin/bash
# ncu --set full of the main kernels of the second decode of tools/decode_once.py (after the plain run exited 0): gpurun -- 'bash tools/gpu_ncu.sh TAG'
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
timeout 300 python tools/decode_once.py 60 > $O/${TAG}_decode_once.log 2>&1 || { echo "plain run failed"; tail $O/${TAG}_decode_once.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_viterbi|k_ofdm_demod|k_gather|k_vit_trace|k_fine_sync' \
    -s 12 -c 12 -f -o $O/${TAG}_full python tools/decode_once.py 60 > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc $?"; tail -5 $O/${TAG}_ncu_full.log
ls -la $O
