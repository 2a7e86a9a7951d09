"""Replays the host eigen-checks of a recorded GPU solve on the CPU.

  RBL_DUMP_T=gpurun_out/config2_T.bin python tools/profile_solve.py      # on the GPU box: records A_i / B_i
  python tools/replay_dump.py gpurun_out/config2_T.bin [tracker_stride]   # here: main checker + emulated tracker

The main checker sees every check (every 4th block step once i*b > k) exactly as the solver's does; the
"tracker" checker is fed every `tracker_stride`-th snapshot with force_full (what the background thread does).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rbl_b200


def load(path):
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:32].view(np.int64)
    m, B, b, final_i = (int(x) for x in hdr)
    body = raw[32:].view(np.float64)
    hA = body[: m * B * B].reshape(m, B, B)
    hB = body[m * B * B: 2 * m * B * B].reshape(m, B, B)
    return m, B, b, final_i, hA, hB


def band(hA, hB, b, it):
    """LAPACK lower band (b+1) x N of T after `it` block steps (B_it not yet applied, common.jl:113)."""
    N = it * b
    ab = np.zeros((b + 1, N))
    for j in range(it):
        A = hA[j][:b, :b]
        for r in range(b):
            for c in range(r + 1):
                ab[r - c, j * b + c] = A[r, c]
        if j > 0:
            Bm = hB[j - 1][:b, :b]          # upper triangular, couples blocks j-1 and j
            for mm in range(b):
                for cc in range(mm, b):
                    row, col = j * b + mm, (j - 1) * b + cc
                    ab[row - col, col] = Bm[mm, cc]
    return ab


if __name__ == "__main__":
    path = sys.argv[1]
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    k = int(os.environ.get("K", "100"))
    threads = int(os.environ.get("THREADS", str(os.cpu_count())))
    m, B, b, final_i, hA, hB = load(path)
    print(f"blocks {m}, b {b}, accepted at {final_i}")
    main = rbl_b200.Checker(threads=threads)
    tracker = rbl_b200.Checker(threads=threads)
    steps = [i for i in range(2, final_i + 1) if i * b > k and i % 4 == 0]
    tw = tt = 0.0
    for n, it in enumerate(steps):
        ab = band(hA, hB, b, it)
        Bi = hB[it - 1][:b, :b]
        t0 = time.perf_counter()
        r = main.check(ab, k, Bi)
        dt = time.perf_counter() - t0
        tw += dt
        line = f"it={it} N={it * b} main: {dt * 1e3:7.1f} ms fac={r['factorizations']} full={int(r['full'])} conv={int(r['converged'])}"
        if n % stride == stride - 1 and it * b >= 2 * k:
            t0 = time.perf_counter()
            q = tracker.check(ab, k, np.zeros((b, b)), tol=0.0, force_full=True)
            dt = time.perf_counter() - t0
            tt += dt
            line += f" | tracker: {dt * 1e3:7.1f} ms fac={q['factorizations']} all={int(q['have_all'])}"
        print(line, flush=True)
    print(f"main checks {tw:.2f} s, tracker {tt:.2f} s")
