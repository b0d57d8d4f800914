"""Host<->device copy bandwidth with pinned buffers of the bench's sizes (0.5 GiB in, 1 GiB out).
    python tools/pcie_probe.py                                            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
                                                                          # all GPUs of the box at the same time
"""
import os

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
hin = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
hout = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
din = torch.empty_like(hin, device=dev)
dout = torch.empty_like(hout, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                      # every rank starts its copies together
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


h2d = t(lambda: din.copy_(hin, non_blocking=True))
d2h = t(lambda: hout.copy_(dout, non_blocking=True))


def both():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


bi = t(both)
print(f"rank {rank}/{world}: H2D 0.5 GiB {h2d:.1f} ms ({0.537 / h2d * 1e3:.1f} GB/s); D2H 1 GiB {d2h:.1f} ms ({1.074 / d2h * 1e3:.1f} GB/s); "
      f"both concurrently {bi:.1f} ms")
if world > 1:
    tt = torch.tensor([h2d, d2h, bi], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"slowest rank, {world} GPUs copying at once: H2D {0.537 / tt[0].item() * 1e3 * world:.0f} GB/s aggregate, "
              f"D2H {1.074 / tt[1].item() * 1e3 * world:.0f} GB/s aggregate, both directions {1.611 / tt[2].item() * 1e3 * world:.0f} GB/s "
              f"aggregate ({tt[2].item():.1f} ms for one bench step's 1.6 GB per GPU)")
    dist.destroy_process_group()
