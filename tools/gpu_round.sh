#!/bin/bash
# One gpurun call's worth of evidence: gpu tests, bench, kernel timeline, ncu launch list, ncu --set full of the top kernels.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh TAG [tests|notests]'
# Every ncu pass runs only after the same command has exited 0 without ncu.  Outputs go to gpurun_out/TAG_*.
TAG=${1:-r02a}
TESTS=${2:-tests}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
if [ "$TESTS" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > $O/${TAG}_pytest.log 2>&1
  echo "pytest rc $?" >> $O/${TAG}_pytest.log
  tail -30 $O/${TAG}_pytest.log
fi
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc $?"; cut -c1-1500 $O/${TAG}_bench.json
timeout 300 python tools/timeline.py 60 > $O/${TAG}_timeline.txt 2>&1
echo "timeline rc $?"; tail -60 $O/${TAG}_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python tools/decode_once.py 60 > $O/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_viterbi|k_ofdm_demod|k_gather|k_vit_trace|k_fine_sync' \
    -s 30 -c 14 -o $O/${TAG}_full python tools/decode_once.py 60 > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc $?"
ls -la $O
