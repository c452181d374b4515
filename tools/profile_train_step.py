"""Where the host time of the training step goes (torch.profiler, CPU + CUDA).  Usage: python tools/profile_train_step.py [c3|c5]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wsovod_b200 import steps, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda", 0)
w = synth.workload(name)
st = steps.make(name, w, dev)
for _ in range(5):
    st()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    st()
b.record()
torch.cuda.synchronize()
print("ms/step", a.elapsed_time(b) / 10)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        st()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=35, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
