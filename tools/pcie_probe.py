"""Host<->device copy bandwidth with pinned buffers of the bench's sizes (0.5 GiB in, 1 GiB out)."""
import torch

dev = torch.device("cuda:0")
hin = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
hout = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
din = torch.empty_like(hin, device=dev)
dout = torch.empty_like(hout, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
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
print(f"H2D 0.5 GiB {h2d:.1f} ms ({0.537 / h2d * 1e3:.1f} GB/s); D2H 1 GiB {d2h:.1f} ms ({1.074 / d2h * 1e3:.1f} GB/s); "
      f"both concurrently {bi:.1f} ms")
