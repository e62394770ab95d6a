#!/bin/bash
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from linkless_link_prediction_b200 import ops
dev = torch.device("cuda:0")
n = 34493
g = torch.Generator(device="cpu").manual_seed(0)
anch = torch.randperm(n, generator=g)[:5362]
a = anch[:, None].expand(-1, 20).reshape(-1)
c = torch.randint(0, n, (a.numel(),), generator=g)
e_u = torch.randint(0, n, (131072,), generator=g); e_v = torch.randint(0, n, (131072,), generator=g)
u = torch.cat([a, e_u]).to(dev); v = torch.cat([c, e_v]).to(dev)
for _ in range(3): ops.EdgePlan(u, v, n)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.EdgePlan(u, v, n); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort(); print("edge plan, physics-student-like batch (5362 anchors x 20 + 131072 edges): median %.1f us" % ts[10])
PY
