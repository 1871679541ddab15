"""HBM ceilings by access mix on this GPU (context for the roofline fractions): write-only
(fill), read-only (sum), copy (read + write).  Plumbing-level torch ops, CUDA-event timed."""
import torch


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = 1 << 29                      # 4 GiB of float64
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    gb = n * 8 / 1e9
    print("write-only fill_  : %.0f GB/s" % (gb / timed(lambda: y.fill_(2.0)) * 1e3))
    print("write-only memset : %.0f GB/s" % (gb / timed(lambda: y.zero_()) * 1e3))
    print("read-only  sum    : %.0f GB/s" % (gb / timed(lambda: x.sum()) * 1e3))
    print("copy (r+w bytes)  : %.0f GB/s" % (2 * gb / timed(lambda: y.copy_(x)) * 1e3))
    z = torch.empty_like(x)
    print("2 reads + 1 write : %.0f GB/s" % (3 * gb / timed(lambda: torch.mul(x, y, out=z)) * 1e3))


if __name__ == "__main__":
    main()
