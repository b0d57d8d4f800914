"""Timeline of the end-to-end host path (process_stacks_host): per chunk, when its copy-in, compute and copy-out ran.
    python tools/e2e_trace.py [stacks] [workers] [schedule as comma-separated slices | -] [quiet]
Prints one line per chunk (ms since the first copy-in) and the GPU-idle gaps between consecutive compute spans."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stacks_host, tapered_schedule  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1024
ops = get_ops()
base = np.stack([synth.ct_slice(1000 + z, z / 64) for z in range(64)])
stack = np.concatenate([base] * (n // 64), 0)
pin = torch.from_numpy(stack.view(np.int16)).pin_memory()
outs = [torch.empty((n, 512, 512), dtype=torch.float32, pin_memory=True) for _ in range(2)]
plan = synth.plan_full()
sched = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 and sys.argv[3] != '-' else tapered_schedule(n, W)
quiet = len(sys.argv) > 4


def run(trace=None):
    return process_stacks_host([stack] * K, plan, ops=ops, pinned_ins=[pin] * K,
                               pinned_outs=[outs[i % 2] for i in range(K)], workers=W, schedule=sched, trace=trace)


run(); run()
torch.cuda.synchronize()
trace = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(trace)
e1.record()
torch.cuda.synchronize()
print(f"{K} stacks, {W} workers, schedule {sched}: {e0.elapsed_time(e1):.2f} ms total, {e0.elapsed_time(e1) / K:.2f} per stack")
rows = []
for wk, k, j, cnt, ev in trace:
    t = [e0.elapsed_time(x) for x in ev]
    rows.append((t[2], wk, k, j, cnt, t))
rows.sort()
if not quiet:
    print(" wk stack chunk slices   in0     in1   comp0   comp1    out1   (ms since start)  comp ms  us/slice")
for _, wk, k, j, cnt, t in ([] if quiet else rows):
    print(f" {wk:2d} {k:5d} {j:5d} {cnt:6d} {t[0]:7.2f} {t[1]:7.2f} {t[2]:7.2f} {t[3]:7.2f} {t[4]:7.2f}"
          f"   {t[3] - t[2]:7.2f}  {(t[3] - t[2]) / cnt * 1e3:7.1f}")
# union of compute spans -> time with no chunk computing
spans = sorted((t[2], t[3]) for _, _, _, _, _, t in rows)
busy, cur0, cur1 = 0.0, spans[0][0], spans[0][1]
for a, b in spans[1:]:
    if a > cur1:
        busy += cur1 - cur0
        cur0, cur1 = a, b
    else:
        cur1 = max(cur1, b)
busy += cur1 - cur0
print(f"some chunk computing for {busy:.2f} ms of {e0.elapsed_time(e1):.2f}; first compute starts at {spans[0][0]:.2f} ms, "
      f"last compute ends at {max(b for _, b in spans):.2f} ms, last copy-out ends at {max(t[4] for *_, t in rows):.2f} ms")
