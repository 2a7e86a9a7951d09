"""Virtual-time replay of a solve's host eigen-checks against a device that produces one block step every STEP_MS.

  python tools/check_timeline_sim.py <T dump> <step_ms> [threads]       # dump: RBL_DUMP_T of a GPU solve, or tools/make_T_dump.py

What it is for: the accepting host check is the serial tail of a row-sharded solve (8 GPUs: a block step takes 0.8 ms, a
full check of T tens of milliseconds), and whether it is cheap depends on how FRESH the background tracker's Ritz pairs are
when it starts.  That interplay cannot be seen in a CPU replay that runs every check back to back (tools/replay_dump.py),
and GPU time is too scarce to tune it there.  Here the main checker and the tracker are the real `BandTopK` objects run on
the recorded T snapshots; only the clock is virtual: each call's measured duration advances the timeline of its own actor
(device / main check thread / tracker thread), following the solver's non-waiting check-point protocol
(`csrc/solver.cu`, Run::cycle: one check in flight, at most 16 speculative steps, the tracker always takes the newest
snapshot, its pairs are handed to the main checker at the start of the next check).  The adaptive check cadence is not
modelled (a check starts at every check point at which none is in flight).  Test / tuning infrastructure only.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rbl_b200
from tools.replay_dump import band, load

if os.environ.get("RBL_LIB"):        # compare builds: RBL_LIB=/path/to/other/librbl_b200.so
    rbl_b200.load_library(os.environ["RBL_LIB"])

MAX_AHEAD = 16
CHECK_PERIOD = 4


def simulate(path, step_ms, threads, k=100, verbose=True, last_step=None):
    m, B, b, final_i, hA, hB = load(path)
    last_step = last_step or m
    main = rbl_b200.Checker(threads=threads)
    tracker = rbl_b200.Checker(threads=max(1, threads - 1))
    zeros = np.zeros((b, b))
    requests = []            # (time posted, step) snapshots offered to the tracker, newest wins
    results = []             # (time done, step, D, S) finished tracker passes
    st = dict(tracker_free=0.0, consumed=0, handed=-1, gifts=[])
    log = []

    def advance_tracker(until):
        while st["consumed"] < len(requests):
            first = requests[st["consumed"]][0]
            start = max(st["tracker_free"], first)
            if start > until:
                break
            cand = [j for j in range(st["consumed"], len(requests)) if requests[j][0] <= start]
            it = requests[cand[-1]][1]
            st["consumed"] = cand[-1] + 1
            ab = band(hA, hB, b, it)
            gifts = [g for g in st["gifts"] if g[0] <= start]          # full checks of the main checker seed the tracker too
            if gifts:
                tracker.set_seeds(gifts[-1][1], gifts[-1][2])
                st["gifts"] = [g for g in st["gifts"] if g[0] > start]
            t0 = time.perf_counter()
            r = tracker.check(ab, k, hB[it - 1][:b, :b], tol=0.0, force_full=True)
            d = (time.perf_counter() - t0) * 1e3
            st["tracker_free"] = start + d
            if r["have_all"]:
                results.append((start + d, it, r["D"], r["S"], r["resid"]))
            log.append(("tracker", it, start, d, r["factorizations"]))

    def run_main(it, now):
        advance_tracker(now)
        done = [j for j, r in enumerate(results) if r[0] <= now]
        seed_it = None
        if done and done[-1] > st["handed"]:
            st["handed"] = done[-1]
            _, seed_it, D, S, rho = results[done[-1]]
            main.set_seeds(D, S, rho if os.environ.get("SIM_NO_SEED_WITNESSES") is None else None)
        ab = band(hA, hB, b, it)
        waited = [0.0, 0.0]

        def need_seeds(N):
            # the solver's hook: a full check without usable seeds waits for the tracker's pass in flight
            tw0 = time.perf_counter()
            elapsed = (tw0 - t0) * 1e3
            advance_tracker(now + elapsed)                 # passes that start before this moment (runs the one in flight)
            cand = [j for j, r_ in enumerate(results) if j > st["handed"] and r_[1] * b >= 0.75 * N]
            if cand:
                j = cand[-1]
                st["handed"] = j
                waited[0] = max(0.0, results[j][0] - (now + elapsed))      # virtual wait until that pass is done
                main.set_seeds(results[j][2], results[j][3], results[j][4])
            waited[1] = (time.perf_counter() - tw0) * 1e3                  # real time spent emulating: not the check's

        if os.environ.get("SIM_NO_WAIT") is None:
            main.set_need_seeds(need_seeds)
        t0 = time.perf_counter()
        r = main.check(ab, k, hB[it - 1][:b, :b])
        d = (time.perf_counter() - t0) * 1e3 - waited[1] + waited[0]
        log.append(("main", it, now, d, r["factorizations"], int(r["full"]), int(r["converged"]), seed_it))
        return r, d

    t_dev, idle = 0.0, 0.0
    in_flight = None         # (step, time done, result)
    accepted = None
    i = 1
    started_tracker = False
    while i < last_step and accepted is None:
        i += 1
        t_dev += step_ms
        if not (i * b > k and i % CHECK_PERIOD == 0):
            continue
        if in_flight and (in_flight[1] <= t_dev or i - in_flight[0] >= MAX_AHEAD):
            if in_flight[1] > t_dev:                       # bounded speculation: the device waits for the check
                idle += in_flight[1] - t_dev
                t_dev = in_flight[1]
            it0, tdone, r = in_flight
            in_flight = None
            if r["converged"]:
                accepted = (it0, t_dev)
                break
        if in_flight is None:
            r, d = run_main(i, t_dev)
            in_flight = (i, t_dev + d, r)
            if not r["converged"] and i * b >= 2 * k:
                requests.append((t_dev + d, i))            # the solver posts the snapshot when the check returns
                if r["have_all"]:
                    st["gifts"].append((t_dev + d, r["D"], r["S"]))
    if accepted is None and in_flight:
        it0, tdone, r = in_flight
        idle += max(0.0, tdone - t_dev)
        t_dev = max(t_dev, tdone)
        if r["converged"]:
            accepted = (it0, t_dev)
    if verbose:
        for e in log:
            if e[0] == "main" and (e[5] or e[3] > 0.75 * MAX_AHEAD * step_ms):
                print(f"  main    step {e[1]:4d} at {e[2]:8.1f} ms: {e[3]:7.1f} ms fac={e[4]} full={e[5]} conv={e[6]} seeds_from={e[7]}")
        ntr = [e for e in log if e[0] == "tracker"]
        if ntr:
            print(f"  tracker passes: {len(ntr)}, last five (step, start ms, ms): " + ", ".join(f"({e[1]}, {e[2]:.0f}, {e[3]:.0f})" for e in ntr[-5:]))
    nmain = [e for e in log if e[0] == "main"]
    return dict(accepted_step=accepted[0] if accepted else None, t_total_ms=accepted[1] if accepted else None,
                device_ms=(accepted[0] if accepted else i) * step_ms, idle_ms=idle, checks=len(nmain),
                full_checks=sum(e[5] for e in nmain), main_ms=sum(e[3] for e in nmain))


if __name__ == "__main__":
    path = sys.argv[1]
    step_ms = float(sys.argv[2])
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count()
    out = simulate(path, step_ms, threads, k=int(os.environ.get("K", "100")))
    print(out)
