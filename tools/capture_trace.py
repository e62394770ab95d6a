"""pytest plugin (development aid): find the source line that invalidates a CUDA stream capture.

    python -m pytest tests/test_gpu_e2e.py -m gpu -q -p tools.capture_trace -s

While a ``torch.cuda.graph`` context is open, every executed line of this package (all threads, the autograd worker
included) is followed by ``cudaStreamIsCapturing`` on the capturing stream; the first line after which the status reads
'invalidated' is printed together with the Python stack.  Nothing here is imported by the product."""
import ctypes
import os
import sys
import threading
import traceback

import torch

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "linkless_link_prediction_b200")
_rt = None
_state = {"stream": None, "reported": False, "last": {}, "depth": 0}


def _cudart():
    global _rt
    if _rt is None:
        for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12"):
            try:
                _rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
    return _rt


def _status():
    st = ctypes.c_int(0)
    _cudart().cudaStreamIsCapturing(ctypes.c_void_p(_state["stream"]), ctypes.byref(st))
    return st.value


def _check(where):
    if _state["reported"] or _state["stream"] is None:
        return
    if _status() == 2:
        _state["reported"] = True
        tid = threading.get_ident()
        print("\n==== CAPTURE INVALIDATED; detected %s" % where, flush=True)
        print("==== last traced line of this thread: %s" % (_state["last"].get(tid),), flush=True)
        for t, loc in _state["last"].items():
            if t != tid:
                print("==== last traced line of thread %d: %s" % (t, loc), flush=True)
        traceback.print_stack()


def _local(frame, event, arg):
    if event in ("line", "return"):
        _check("before %s:%d (%s)" % (frame.f_code.co_filename, frame.f_lineno, event))
        _state["last"][threading.get_ident()] = (frame.f_code.co_filename, frame.f_lineno)
    return _local


def _global(frame, event, arg):
    if _state["stream"] is None:
        return None
    if frame.f_code.co_filename.startswith(PKG):
        return _local
    return None


_enter, _exit = torch.cuda.graph.__enter__, torch.cuda.graph.__exit__


def _traced_enter(self):
    r = _enter(self)
    _state["depth"] += 1
    _state["stream"] = torch.cuda.current_stream().cuda_stream
    _state["reported"] = False
    _state["last"] = {}
    threading.settrace_all_threads(_global)
    sys.settrace(_global)
    # frames already on the stack (the caller of the with-block) need a local tracer too
    f = sys._getframe(1)
    while f is not None:
        if f.f_code.co_filename.startswith(PKG):
            f.f_trace = _local
        f = f.f_back
    return r


def _traced_exit(self, *exc):
    _check("at the end of the capture")
    _state["depth"] -= 1
    threading.settrace_all_threads(None)
    sys.settrace(None)
    _state["stream"] = None
    return _exit(self, *exc)


torch.cuda.graph.__enter__ = _traced_enter
torch.cuda.graph.__exit__ = _traced_exit
