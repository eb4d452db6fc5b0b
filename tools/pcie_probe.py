"""Measures what the host link gives the host-vector SpMV path (development
aid): pinned H2D, D2H and both at once, for the 64 MB vectors of config 2."""
import sys

import torch


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
    hx = torch.empty(n, dtype=torch.float64).pin_memory()
    hy = torch.empty(n, dtype=torch.float64).pin_memory()
    dx = torch.empty(n, dtype=torch.float64, device="cuda")
    dy = torch.empty(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    mb = n * 8 / 1e6

    t = timed(lambda: dx.copy_(hx, non_blocking=True))
    print("H2D %.0f MB: %.3f ms  %.1f GB/s" % (mb, t, mb / t))
    t = timed(lambda: hy.copy_(dy, non_blocking=True))
    print("D2H %.0f MB: %.3f ms  %.1f GB/s" % (mb, t, mb / t))

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s2):
            hy.copy_(dy, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    t = timed(both)
    print("H2D + D2H concurrently: %.3f ms  %.1f GB/s each way" % (t, mb / t))
    for chunks in (4, 16, 64):
        c = n // chunks

        def chunked():
            for k in range(chunks):
                dx[k * c:(k + 1) * c].copy_(hx[k * c:(k + 1) * c],
                                            non_blocking=True)

        t = timed(chunked)
        print("H2D in %d chunks: %.3f ms  %.1f GB/s" % (chunks, t, mb / t))


if __name__ == "__main__":
    main()
