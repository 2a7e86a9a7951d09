#!/usr/bin/env python
"""bench.py - time-to-k-eigenpairs of the RBL hot path on BASELINE.json configs[1].

Workload (N=1 and N>1, "strong" scaling): 3-D 7-point Laplacian 100^3 (n = 10^6, nnz = 6.94e6), the 100 lowest
eigenpairs as the 100 largest of 12*I - A, block size b = 16, fp64 SpMM / 3-term / QR with an fp32 Krylov buffer
and fp32 re-orthogonalisation ("mixed", the reference's README.md:69 split).  One "step" is one complete solve
(random start block -> Lanczos iteration with all convergence checks -> Ritz vectors).

    value   seconds per solve with A, Omega and V resident in HBM          (rbl_solve_device)
    e2e     seconds per solve through the reference-facing call RBL_gpu(A,k,b) with HOST buffers: upload of A
            (CSR) and Omega, download of V inside the timed region         (rbl_create + rbl_solve + rbl_destroy)
    roofline   the dominant kernel family (re-orthogonalisation Gram / update against the Krylov buffer):
            algorithmic HBM bytes / CUDA-event time inside the library's stream, against MEASURED_PEAKS.json
    cpu_baseline   the oracle restatement of RBL.jl on the box's host cores: a bounded sample (first block
            steps of the same problem) extrapolated with the measured per-phase costs - a reported baseline.

`--impl reference` times only that CPU restatement (the reference itself is Julia and cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GRID = 100
K_WANTED = 100
BLOCK = 16
SIGMA = 12.0
MAX_KRYL = 9600
SEED = 20260
METRIC = "time_to_k_eigenpairs"


def problem():
    from oracle import matrices
    L = matrices.laplacian_3d(GRID).tocsr()
    L.sort_indices()
    return L


def omega(n, b):
    return np.random.default_rng(SEED).standard_normal((n, b))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: only samples taken after this call are reported."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s in sm if s > 0.5 * max(mx + [1])] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline
# block steps the fixed workload (seeded Omega) needs to converge: measured by the GPU arm (oracle-identical iteration
# rule); the reference arm cannot afford to run the CPU solve to the end (~2e4 s) to find out
KNOWN_BLOCK_STEPS = 492


def cpu_reference(iterations_needed: int | None, budget_steps: int = 20):
    """Oracle restatement of RBL.jl (fp64 as shipped) on the host cores: first `budget_steps` block steps of the
    same problem, then extrapolation to `iterations_needed` steps with the measured per-phase costs."""
    from oracle import matrices, rbl_oracle
    from scipy.linalg import lapack
    L = problem()
    A = matrices.shifted(L, SIGMA)
    n = A.shape[0]
    Om = omega(n, BLOCK)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    _, _, det = rbl_oracle.RBL(A, K_WANTED, BLOCK, Om, max_kryl_sz=MAX_KRYL, max_iterations=budget_steps,
                               return_details=True)
    t_sample = time.perf_counter() - t0
    st = det["stats"]
    m0 = st.iterations
    sec = st.seconds
    # dsbev('V') cost model c*N^3, measured at N=640 (bandwidth 16), RBL.jl:107 / common.jl:36-48
    N_e = 640
    rng = np.random.default_rng(0)
    ab = rng.standard_normal((BLOCK + 1, N_e))
    te = time.perf_counter()
    lapack.dsbev(np.asfortranarray(ab), compute_v=1, lower=1)
    c_eig = (time.perf_counter() - te) / N_e ** 3
    m = iterations_needed or KNOWN_BLOCK_STEPS
    per_step = (sec["A*Q"] + sec["3-term"] + sec["QR"] + sec["Loc reorth"]) / m0
    blocks0 = sum(i - 2 for i in range(2, m0 + 1, 2))
    c_reorth = sec["Part reorth"] / max(blocks0, 1)
    blocks = sum(i - 2 for i in range(2, m + 1, 2))
    t_eig = sum(c_eig * (i * BLOCK) ** 3 for i in range(4, m + 1, 4) if i * BLOCK > K_WANTED)
    t_ritz = 2.0 * n * m * BLOCK * K_WANTED / 5e10  # dgemm at ~50 GFLOP/s
    est_no_eig = m * per_step + c_reorth * blocks + t_ritz
    est = est_no_eig + t_eig
    return {
        "value": est, "unit": "s", "cores": cores, "kind": "port",
        "sample": (f"oracle restatement of RBL.jl (NumPy/SciPy-OpenBLAS, fp64 as shipped, {cores} threads): first {m0} block "
                   f"steps of the same problem measured ({t_sample:.1f} s: per-step {per_step:.3f} s, part-reorth "
                   f"{c_reorth * 1e3:.1f} ms per stored block), dsbev('V') measured at N={N_e} ({c_eig * N_e ** 3:.2f} s, "
                   f"cubic model); extrapolated to the {m} block steps the solve needs: "
                   f"{est_no_eig:.0f} s without the host eigensolves + {t_eig:.0f} s of dsbev"),
        "measured_sample_s": t_sample, "estimate_without_dsbev_s": est_no_eig, "estimate_dsbev_s": t_eig,
        "block_steps_extrapolated_to": m,
    }


def run_reference(args, rank, out=sys.stdout):
    if rank != 0:
        return
    cb = cpu_reference(None, budget_steps=args.ref_steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["value"] * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


def config_dict(ngpu):
    return {"workload": f"configs[1]: 3D 7-point Laplacian {GRID}^3 (n=1e6, nnz=6.94e6), {K_WANTED} lowest eigenpairs via "
                        f"{SIGMA:g}I-A, b={BLOCK}, fp64 SpMM/QR + fp32 Krylov buffer/reorth (mixed), tol=1e-7, full solve",
            "n": GRID ** 3, "k": K_WANTED, "b": BLOCK, "precision": "mixed", "max_kryl_sz": MAX_KRYL,
            "row_shards": ngpu, "l2": "working set (Krylov buffer, tens of GB) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--verbose", type=int, default=0)
    args = ap.parse_args()
    # keep stdout to the ONE JSON line: libraries (NCCL prints its version banner) write to fd 1, so point fd 1
    # at stderr for the whole run and print the result line on the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    import torch
    import rbl_b200
    from rbl_b200 import binding as B
    if not torch.cuda.is_available() or rbl_b200.lib().rbl_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: rbl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    L = problem()
    n = L.shape[0]
    Om = omega(n, BLOCK)
    cores = os.cpu_count() or 1
    host_threads = cores  # only rank 0 evaluates the host eigen-checks, so it may use every core of the box
    opt_kw = dict(max_kryl_sz=MAX_KRYL, precision=B.PRECISION_MIXED if args.precision == "mixed" else B.PRECISION_FP64,
                  op=B.OP_SHIFT_MINUS_A, sigma=SIGMA, device=local_rank, async_check=1, host_threads=host_threads,
                  verbose=args.verbose)

    if world > 1:
        rs = rbl_b200.partition_rows(n, world)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        def fresh_uid():
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                import ctypes
                buf = ctypes.create_string_buffer(128)
                assert rbl_b200.lib().rbl_nccl_unique_id(buf) == 0, rbl_b200.lib().rbl_last_error()
                uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            return bytes(uid.cpu().numpy().tobytes())
        Lloc = L[r0:r1, :]
        Lloc.sort_indices()

        def make_solver():
            return B.Solver(options=B.default_options(**opt_kw),
                            shard=dict(n=n, row0=r0, rowptr=Lloc.indptr.astype(np.int64), colidx=Lloc.indices.astype(np.int64),
                                       vals=Lloc.data, rank=rank, world=world, uid=fresh_uid()))
        Om_loc = np.asfortranarray(Om[r0:r1])
        nloc = r1 - r0
    else:
        def make_solver():
            return B.Solver(L, options=B.default_options(**opt_kw))
        Om_loc = np.asfortranarray(Om)
        nloc = n

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: everything resident in HBM ------------------------------------------------------------
    solver = make_solver()
    om_dev = torch.from_numpy(np.ascontiguousarray(Om_loc.T)).to(dev)          # (b, nloc) row-major == nloc x b column-major
    v_dev = torch.empty((K_WANTED, nloc), dtype=torch.float64, device=dev)    # nloc x k column-major
    stats = None
    # the nvidia-smi poller is started before the last warm-up solve: its start-up (NVML initialisation, device
    # enumeration) was measured to stall the first solve after it by 0.2-0.6 s; it keeps sampling every 200 ms
    # through the timed region and only those samples are reported.  Rank 0 samples its own GPU.
    sampler = ClockSampler(local_rank)
    for w in range(args.warmup):
        if w == args.warmup - 1 and rank == 0:
            sampler.start()
        D, stats = solver.solve_device(K_WANTED, BLOCK, om_dev.data_ptr(), v_dev.data_ptr())
    if args.warmup == 0 and rank == 0:
        sampler.start()
    barrier()
    sampler.mark()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    agg = {}
    for _ in range(args.steps):
        tw = time.perf_counter()
        D, stats = solver.solve_device(K_WANTED, BLOCK, om_dev.data_ptr(), v_dev.data_ptr())
        if args.verbose:
            print(f"[bench] solve wall {time.perf_counter() - tw:.3f} s, library t_total {stats.t_total:.3f} s", file=sys.stderr)
        for f, v in stats.as_dict().items():
            agg[f] = agg.get(f, 0) + v
    e1.record()
    barrier()
    clocks = sampler.stop()
    t_dev = e0.elapsed_time(e1) * 1e-3
    if dist is not None:
        tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    sec_per_solve = t_dev / args.steps
    solver.close()

    # ---- e2e: the reference-facing call with host buffers ------------------------------------------------
    om_pin = torch.from_numpy(np.asfortranarray(Om_loc).T.copy()).pin_memory()   # pinned host Omega (column-major view)
    v_pin = torch.empty((K_WANTED, nloc), dtype=torch.float64).pin_memory()
    import ctypes as C

    def e2e_once():
        s = make_solver()                                                        # uploads A (rbl_create)
        Dh = np.zeros(K_WANTED)
        st = B.RblStats()
        rc = rbl_b200.lib().rbl_solve(s._h, K_WANTED, BLOCK, C.cast(om_pin.data_ptr(), C.POINTER(C.c_double)),
                                      Dh.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(v_pin.data_ptr()), C.byref(st))
        if rc != 0:
            raise RuntimeError(rbl_b200.lib().rbl_last_error().decode())
        s.close()
        return Dh, st
    for _ in range(max(1, args.warmup)):     # same warm-up count as the device-resident arm (first calls grow the workspace)
        e2e_once()
    barrier()
    f0 = torch.cuda.Event(enable_timing=True)
    f1 = torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        Dh, st_e2e = e2e_once()
    f1.record()
    barrier()
    t_e2e = f0.elapsed_time(f1) * 1e-3
    if dist is not None:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    nnz_loc = (Lloc.nnz if world > 1 else L.nnz)
    h2d = 4 * (nloc + 1) + 12 * nnz_loc + 8 * nloc * BLOCK
    d2h = 8 * nloc * K_WANTED
    if dist is not None:
        hb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(hb)
        h2d, d2h = int(hb[0].item()), int(hb[1].item())

    # ---- checks on the result (eigenvalues against the analytic spectrum) -----------------------------------
    from oracle import matrices
    exact = SIGMA - matrices.laplacian_eigs(GRID, 3, K_WANTED)
    eig_err = float(np.max(np.abs(D - exact) / exact))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    steps = args.steps
    g_t, u_t = agg["t_reorth_gram"], agg["t_reorth_update"]
    g_b, u_b = agg["bytes_reorth_gram"], agg["bytes_reorth_update"]
    g_n, u_n = max(agg["launches_reorth_gram"], 1), max(agg["launches_reorth_update"], 1)
    dom = "reorth_gram" if g_t >= u_t else "reorth_update"
    d_t, d_b, d_n = (g_t, g_b, g_n) if dom == "reorth_gram" else (u_t, u_b, u_n)
    achieved = d_b / d_t / 1e9 if d_t > 0 else 0.0
    peak = float(peaks["hbm_gbs"])
    # DRAM bytes of the dominant kernel from the committed ncu --set full capture: the captured launch's
    # (read + write) / algorithmic ratio applied to this run's average launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        cap = tj[dom]
        ratio = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap["algorithmic_bytes"]
        traffic = ratio * d_b / d_n
        traffic_src = f"ncu capture of one launch (m={cap['m']}): DRAM read+write = {ratio:.4f} x algorithmic bytes, scaled to the average launch"
    except Exception:
        pass
    roof = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
        "avg_launch_ms": d_t / d_n * 1e3, "algorithmic_bytes_per_launch": d_b / d_n, "launches": int(d_n),
        "share_of_step": d_t / (sec_per_solve * steps),
        "others": {
            "reorth_gram": {"GB/s": g_b / g_t / 1e9 if g_t else 0, "s_per_solve": g_t / steps},
            "reorth_update": {"GB/s": u_b / u_t / 1e9 if u_t else 0, "s_per_solve": u_t / steps},
            "spmm": {"GB/s": agg["bytes_spmm"] / agg["t_spmm"] / 1e9 if agg["t_spmm"] else 0, "s_per_solve": agg["t_spmm"] / steps},
            "ritz": {"TFLOP/s": agg["flops_ritz"] / agg["t_ritz_kernel"] / 1e12 if agg["t_ritz_kernel"] else 0,
                     "GB/s": agg["bytes_ritz"] / agg["t_ritz_kernel"] / 1e9 if agg["t_ritz_kernel"] else 0,
                     "s_per_solve": agg["t_ritz_kernel"] / steps},
            "3term_s_per_solve": agg["t_3term"] / steps, "qr_s_per_solve": agg["t_qr"] / steps,
            "loc_reorth_s_per_solve": agg["t_loc_reorth"] / steps,
            "host_eig_s_per_solve": agg["t_eig"] / steps, "host_eig_wait_s_per_solve": agg["t_eig_wait"] / steps,
        },
    }
    iterations = int(round(agg["iterations"] / steps))
    if args.no_cpu_baseline or world > 1:
        cb = {"value": None, "unit": "s", "cores": cores, "kind": "port", "sample": "skipped (N>1 or --no-cpu-baseline)"}
    else:
        cb = cpu_reference(iterations, budget_steps=args.ref_steps)
    line = {
        "metric": METRIC, "value": sec_per_solve, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_solve * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.precision != "mixed" else "f64+f32", "data": "synthetic", "config": config_dict(world),
        "e2e": {"value": t_e2e / steps, "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(agg["kernel_launches"]),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cb,
        "solve": {"block_steps": iterations, "kryl_sz": iterations * BLOCK, "block_steps_run": int(round(agg["iterations_run"] / steps)),
                  "checks": int(round(agg["checks"] / steps)), "full_checks": int(round(agg["full_checks"] / steps)),
                  "max_rel_eig_err_vs_analytic": eig_err, "host_cores": cores},
    }
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
