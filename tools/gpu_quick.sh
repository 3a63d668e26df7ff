#!/bin/bash
# quick GPU check of a kernel change: the stage / whole-path parity tests, then the kernel timeline of the bench workload
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh TAG [full]'
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
if [ "$2" = full ]; then SEL="tests"; else SEL="tests/test_gpu_parity.py tests/test_gpu_paths.py"; fi
timeout 800 python -m pytest $SEL -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
echo "pytest rc $?"; tail -4 $O/${TAG}_pytest.log
timeout 200 python tools/timeline.py 60 > $O/${TAG}_timeline.txt 2>&1
tail -24 $O/${TAG}_timeline.txt
