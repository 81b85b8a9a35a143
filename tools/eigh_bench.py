"""Times torch.linalg.eigh (the one library call KFAC keeps, reference kfac.py:213-217) on the ACKTR factor sizes."""
import json
import torch
dev = "cuda:0"
sizes = [257, 513, 577, 1569, 32, 64, 32, 512, 513, 1, 513, 6]
mats = []
for n in sizes:
    a = torch.randn(n, n, device=dev)
    mats.append(a @ a.t() / n + 0.01 * torch.eye(n, device=dev))
def run():
    return [torch.linalg.eigh(m) for m in mats]
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
per = {}
for n, m in zip(sizes, mats):
    e0.record(); torch.linalg.eigh(m); e1.record(); torch.cuda.synchronize()
    per[n] = e0.elapsed_time(e1)
print(json.dumps({"total_ms": tot, "per_size_ms": per}))
