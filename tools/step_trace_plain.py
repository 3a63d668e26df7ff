"""Host-side timeline (B200DAB_TRACE) of device-resident decodes of the bench workload with profiling events OFF (as ab_step.py times
them): where a step's time goes on the host -- submission, the wait for the GPU, the assembly of the result.
python tools/step_trace_plain.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["B200DAB_TRACE"] = "1"
import torch
import bench
from abracadabra_b200 import pipeline
raw, res, nf = bench.make_recording(60.0, 0)
rx = pipeline.Receiver(0)
d = torch.from_numpy(raw).cuda()
n = raw.size // 2
for i in range(6):
    torch.cuda.synchronize()
    t = time.perf_counter()
    rx.decode_raw(d.data_ptr(), n, "u8", True, None, pipeline.OPT_TII)
    print("device-buffer decode", round((time.perf_counter() - t) * 1e3, 3), "ms", file=sys.stderr)
