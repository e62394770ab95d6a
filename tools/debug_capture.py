#!/usr/bin/env python
"""Find the call that invalidates a stream capture of the student step (development aid): every C-ABI call is followed by
cudaStreamIsCapturing on the current stream; the first call after which the status is 'invalidated' is reported."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import linkless_link_prediction_b200 as L  # noqa: E402
from linkless_link_prediction_b200 import _native as N  # noqa: E402
from linkless_link_prediction_b200 import main as student  # noqa: E402
from linkless_link_prediction_b200 import ops, shims  # noqa: E402
from linkless_link_prediction_b200.data import synthetic_dataset  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ops.set_compute_dtype(mode)
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
try:
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
except OSError:
    rt = None


def status():
    if rt is None:
        return -1
    st = ctypes.c_int(0)
    rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return st.value


lib = N.load()
seen = {"bad": None}
for name in N.PROTOTYPES:
    fn = getattr(lib, name)

    def make(fn=fn, name=name):
        def wrapped(*a):
            before = status()
            r = fn(*a)
            after = status()
            if seen["bad"] is None and before == 1 and after == 2:
                seen["bad"] = name
                print("CAPTURE INVALIDATED BY", name, "rc", r, flush=True)
            return r
        return wrapped
    setattr(lib, name, make())

import tempfile

from linkless_link_prediction_b200 import train_teacher_gnn as teacher  # noqa: E402

work = os.path.join(tempfile.mkdtemp(), "src")
os.makedirs(work)
os.chdir(work)
common = ["--datasets=cora", "--encoder=sage", "--transductive=transductive", "--runs=1", "--epochs=2", "--synthetic_scale=0.2",
          "--precision=" + mode]
teacher.main(common + ["--hidden_channels=256", "--batch_size=512"])
print("teacher done; status", status(), "invalidated by:", seen["bad"], flush=True)
student.main(common + ["--hidden_channels=256", "--link_batch_size=512", "--LLP_D=1", "--LLP_R=1", "--True_label=1",
                       "--dropout=0.0", "--rw_step=2", "--hops=2", "--ns_rate=1"])
print("done; invalidated by:", seen["bad"])
