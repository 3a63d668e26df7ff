"""Host-side timeline (B200DAB_TRACE) of device-resident decodes of the bench workload, profiling events on as in bench.py:
python tools/step_trace.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["B200DAB_TRACE"] = "1"
import torch
import bench
from abracadabra_b200 import pipeline
raw, res, nf = bench.make_recording(60.0, 0)
rx = pipeline.Receiver(0)
rx.set_profiling(True)
d = torch.from_numpy(raw).cuda()
n = raw.size // 2
for i in range(4):
    t = time.perf_counter()
    rx.decode_raw(d.data_ptr(), n, "u8", True, None, pipeline.OPT_TII)
    print("device-buffer decode", round((time.perf_counter() - t) * 1e3, 3), "ms", file=sys.stderr)
print(sorted(rx.kernel_times(), key=lambda x: -x[1]), file=sys.stderr)
