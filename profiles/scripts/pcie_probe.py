"""Host<->device copy bandwidth of this box (pinned memory): one stream, two concurrent streams,
both directions at once.  Gives the PCIe bound of the end-to-end (host pointer) numbers."""
import time
import torch

dev = torch.device("cuda", 0)
N = 1 << 27  # 1 GiB of doubles
h = [torch.empty(N, dtype=torch.float64, pin_memory=True) for _ in range(2)]
d = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(2)]
s = [torch.cuda.Stream() for _ in range(2)]
for t in h:
    t.zero_()


def run(fn, nbytes, label):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("%-34s %.1f GB/s" % (label, nbytes / dt / 1e9))


B = N * 8
run(lambda: d[0].copy_(h[0], non_blocking=True), B, "H2D one stream")
run(lambda: h[0].copy_(d[0], non_blocking=True), B, "D2H one stream")


def two_h2d():
    for i in range(2):
        with torch.cuda.stream(s[i]):
            d[i].copy_(h[i], non_blocking=True)


def bidir():
    with torch.cuda.stream(s[0]):
        d[0].copy_(h[0], non_blocking=True)
    with torch.cuda.stream(s[1]):
        h[1].copy_(d[1], non_blocking=True)


run(two_h2d, 2 * B, "H2D two streams (aggregate)")
run(bidir, 2 * B, "H2D + D2H concurrent (aggregate)")
