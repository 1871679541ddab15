"""Device -> pinned-host copy rates: one contiguous block vs the strided row copy the pipeline
uses (rows of one chunk into a [rows][B] host buffer, jt_copy_rows / cudaMemcpy2DAsync)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import torch  # noqa: E402
from junctiontree import _native  # noqa: E402


def main():
    rows, chunk, B = 2092, 8192, 65536
    dev = torch.empty((rows, chunk), dtype=torch.float64, device="cuda")
    host2d = torch.empty((rows, B), dtype=torch.float64).pin_memory()
    host1d = torch.empty((rows, chunk), dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    nbytes = rows * chunk * 8
    t1 = timed(lambda: host1d.copy_(dev, non_blocking=True))
    t2 = timed(lambda: _native.copy_rows(host2d.data_ptr(), B * 8, dev.data_ptr(), chunk * 8, chunk * 8, rows, True,
                                         stream))
    print("contiguous %.1f MB: %.3f ms, %.1f GB/s" % (nbytes / 1e6, t1, nbytes / t1 / 1e6))
    print("strided rows (%d x %d KB, pitch %d KB): %.3f ms, %.1f GB/s" % (rows, chunk * 8 // 1024, B * 8 // 1024, t2,
                                                                           nbytes / t2 / 1e6))


if __name__ == "__main__":
    main()
