#!/bin/bash
# Final evidence of a round on ONE box: bench (default), launch list, ncu --set full of the workload's kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh TAG'
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 400 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; echo "reference arm rc $?"
timeout 100 python tools/timeline.py 60 > $O/${TAG}_timeline.txt 2>&1; echo "timeline rc $?"
export B200DAB_TRACE_PDL=0      # under ncu every kernel is serialised anyway; keep the launches plain
timeout 100 python tools/decode_once.py 60 > $O/${TAG}_decode_once.log 2>&1 || { echo "decode_once failed"; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python tools/decode_once.py 60 > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_viterbi|k_ofdm_demod|k_gather|k_vit_trace|k_fine_sync|k_block_power|k_coarse|k_sf_decode|k_fire_align|k_au_list|k_tii' \
    -s 21 -c 21 -f -o $O/${TAG}_full python tools/decode_once.py 60 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la $O | tail -12
