"""Run a few SpMM launches at the collab size (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linkless_link_prediction_b200 import ops
from linkless_link_prediction_b200.data import undirected_graph
dev = torch.device("cuda:0")
n = 235868
g = ops.Graph(undirected_graph(n, 1179052, 0, True, unique=False).to(dev), n)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for dt, F in ((torch.bfloat16, 256), (torch.bfloat16, 128), (torch.float32, 256)):
    x = torch.randn(n, F, device=dev).to(dt)
    for _ in range(2):
        flush.zero_()
        y = g.spmm(x)
torch.cuda.synchronize()
print("ok")
