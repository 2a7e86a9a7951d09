"""BASELINE.json configs 3, 4 and 5 AT FULL SIZE through the drop-in entry (restart + Chebyshev filter, N1).

    python tools/run_config.py 3 [--n 10000000] [--degree 9] [--k 50] [--b 32]            ER, 1 GPU, Krylov slab bounded by HBM
    python tools/run_config.py 4 [--side 4096] [--degree 40] [--k 64] [--b 16]           image-grid graph Laplacian, 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/run_config.py 5 [--grid 400] [--degree 16]   400^3 Laplacian

Prints one JSON line per run with the three north-star gates measured on the result (eigenvalue accuracy where an
analytic spectrum exists, max ||A v - lambda v|| / ||A||, ||V'V - I||) and the solve statistics.  Matrices are generated
with fast vectorised builders (the scipy.sparse.kron / coo paths of oracle/matrices.py take minutes at these sizes);
tests/test_config_generators.py checks them against oracle/matrices.py at small sizes.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp


# ---- fast generators ------------------------------------------------------------------------------------------------
def er_sym_fast(n, nnz_per_row=32, seed=0):
    """Same distribution as oracle.matrices.erdos_renyi_sym (n*nnz_per_row/2 uniform (i,j) pairs, N(0,1) weights,
    U = triu(.,1), A = U + U'), built with one sort instead of COO -> CSR conversions."""
    rng = np.random.default_rng(seed)
    m = n * nnz_per_row // 2
    i = rng.integers(0, n, m)
    j = rng.integers(0, n, m)
    w = rng.standard_normal(m)
    keep = i < j                       # triu(., 1) of the COO matrix: entries below / on the diagonal are dropped
    i, j, w = i[keep], j[keep], w[keep]
    rows = np.concatenate([i, j])
    cols = np.concatenate([j, i])
    vals = np.concatenate([w, w])
    key = rows * np.int64(n) + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    first = np.concatenate([[True], key[1:] != key[:-1]])
    idx = np.flatnonzero(first)
    vals = np.add.reduceat(vals, idx)    # duplicate pairs add up, as in coo -> csr
    key = key[idx]
    rows = key // n
    cols = key - rows * n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.bincount(rows, minlength=n)
    rowptr = np.cumsum(rowptr)
    return sp.csr_matrix((vals, cols, rowptr), shape=(n, n))


def laplacian_3d_rows(N, r0, r1):
    """Rows [r0, r1) of the 7-point Dirichlet Laplacian on an N^3 grid (index = (z*N + y)*N + x) as local CSR with
    global columns - what each rank of a row-sharded solve builds for itself."""
    r = np.arange(r0, r1, dtype=np.int64)
    x = r % N
    y = (r // N) % N
    z = r // (N * N)
    offs = np.array([-N * N, -N, -1, 0, 1, N, N * N], dtype=np.int64)
    valid = np.stack([z > 0, y > 0, x > 0, np.ones_like(x, dtype=bool), x < N - 1, y < N - 1, z < N - 1], axis=1)
    cols = r[:, None] + offs[None, :]
    vals = np.where(offs[None, :] == 0, 6.0, -1.0) * np.ones_like(cols, dtype=np.float64)
    rowptr = np.concatenate([[0], np.cumsum(valid.sum(axis=1))]).astype(np.int64)
    return rowptr, cols[valid], vals[valid]


def image_laplacian_fast(H, W, seed=0, sigma2=0.05):
    """oracle.matrices.image_graph_laplacian (8-neighbour weights exp(-(dI)^2/sigma2), L = D - W) built directly as CSR."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = 0.5 + 0.25 * np.sin(2 * np.pi * xx / max(W, 1) * 3) * np.cos(2 * np.pi * yy / max(H, 1) * 2)
    img += 0.25 * ((xx > W // 2) ^ (yy > H // 3))
    img += 0.02 * rng.standard_normal((H, W))
    n = H * W
    flat = img.ravel()
    idx = np.arange(n, dtype=np.int64)
    y = idx // W
    x = idx % W
    nb = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]   # sorted by column offset
    cols = np.empty((n, 9), dtype=np.int64)
    vals = np.zeros((n, 9))
    valid = np.zeros((n, 9), dtype=bool)
    for c, (dy, dx) in enumerate(nb):
        ok = (y + dy >= 0) & (y + dy < H) & (x + dx >= 0) & (x + dx < W)
        j = idx + dy * W + dx
        cols[:, c] = j
        valid[:, c] = ok
        if (dy, dx) != (0, 0):
            jj = np.where(ok, j, idx)
            wgt = np.exp(-((flat - flat[jj]) ** 2) / sigma2)
            vals[:, c] = np.where(ok, -wgt, 0.0)
    vals[:, 4] = -vals.sum(axis=1)          # degree on the diagonal
    rowptr = np.concatenate([[0], np.cumsum(valid.sum(axis=1))]).astype(np.int64)
    return sp.csr_matrix((vals[valid], cols[valid], rowptr), shape=(n, n))


def gates(A, D, V, norm_a):
    R = A @ V - V * D[None, :]
    G = V.T @ V
    return {"max_ritz_residual_over_normA": float(np.max(np.linalg.norm(R, axis=0)) / norm_a),
            "v_orthonormality_2norm": float(np.linalg.norm(G - np.eye(V.shape[1]), 2))}


def stats_dict(st):
    keep = ("iterations", "kryl_sz", "iterations_run", "converged", "checks", "restarts", "locked", "buffer_blocks", "spilled_blocks",
            "filter_degree", "filter_two_sided", "filter_cut", "max_residual", "t_total", "t_spmm", "t_3term", "t_qr", "t_part_reorth",
            "t_loc_reorth", "t_eig", "t_ritz", "t_eig_wait", "launches_spmm", "kernel_launches", "bytes_spmm")
    return {k: getattr(st, k) for k in keep}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int, choices=[3, 4, 5])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--pairs-per-row", type=int, default=64, help="config 3: n*this/2 (i,j) draws; triu keeps half -> ~this/2 nnz per row")
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--grid", type=int, default=400)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--b", type=int, default=0)
    ap.add_argument("--degree", type=int, default=-1)
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--max-kryl", type=int, default=0)
    ap.add_argument("--verbose", type=int, default=1)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    import rbl_b200
    from rbl_b200 import binding as B
    from oracle import matrices
    prec = B.PRECISION_MIXED if a.precision == "mixed" else B.PRECISION_FP64
    out = {"config": a.config, "precision": a.precision}
    t_gen = time.perf_counter()
    if a.config == 3:
        k, b = a.k or 50, a.b or 32
        A = er_sym_fast(a.n, a.pairs_per_row, seed=3)
        n = A.shape[0]
        out.update(workload=f"configs[2]: symmetric Erdos-Renyi n={n}, nnz={A.nnz}, {k} extreme eigenpairs, b={b}, 1 GPU", n=n, nnz=int(A.nnz))
        opts = B.default_options(max_kryl_sz=a.max_kryl or 100000, precision=prec, restart=1, filter_degree=a.degree if a.degree >= 0 else 9,
                                 verbose=a.verbose)
        norm_a, sigma = None, None
    elif a.config == 4:
        k, b = a.k or 64, a.b or 16
        L = image_laplacian_fast(a.side, a.side, seed=0)
        n = L.shape[0]
        dmax = float(L.diagonal().max())
        sigma = 2.0 * dmax                      # Gershgorin: lambda_max(L) <= 2 max degree
        A = L
        out.update(workload=f"configs[3]: image-grid graph Laplacian {a.side}x{a.side} (n={n}, nnz={L.nnz}), {k} smallest eigenpairs "
                            f"via {sigma:.3f}I - L, b={b}, 1 GPU", n=n, nnz=int(L.nnz), sigma=sigma)
        opts = B.default_options(max_kryl_sz=a.max_kryl or 100000, precision=prec, restart=1, op=B.OP_SHIFT_MINUS_A, sigma=sigma,
                                 filter_degree=a.degree if a.degree >= 0 else 40, verbose=a.verbose)
        norm_a = sigma
    else:
        return config5(a, rbl_b200, B, prec)
    out["generate_s"] = time.perf_counter() - t_gen
    Om = np.random.default_rng(1).standard_normal((n, b))
    t0 = time.perf_counter()
    with B.Solver(A, options=opts) as s:
        out["create_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        D, V, st = s.solve(k, b, Om, allow_not_converged=True)
        out["solve_s"] = time.perf_counter() - t1
    out["stats"] = stats_dict(st)
    out["D_first_last"] = [float(D[0]), float(D[-1])]
    if not a.no_check:
        Aop = A if sigma is None else (sigma * sp.identity(n, format="csr") - A).tocsr()
        out["gates"] = gates(Aop, D, V, norm_a if norm_a is not None else float(np.max(np.abs(D))))
        out["descending_abs"] = bool(np.all(np.abs(D)[:-1] >= np.abs(D)[1:] * (1 - 1e-12)))
    print(json.dumps(out), flush=True)


def config5(a, rbl_b200, B, prec):
    """400^3 Laplacian row-sharded over the ranks of torchrun (one process per GPU)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from oracle import matrices
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = a.grid
    n = N ** 3
    k, b = a.k or 100, a.b or 16
    rs = rbl_b200.partition_rows(n, world)
    r0, r1 = int(rs[rank]), int(rs[rank + 1])
    t0 = time.perf_counter()
    rowptr, cols, vals = laplacian_3d_rows(N, r0, r1)
    t_gen = time.perf_counter() - t0
    uid = None
    if world > 1:
        u = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = C.create_string_buffer(128)
            assert rbl_b200.lib().rbl_nccl_unique_id(buf) == 0
            u.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(u, 0)
        uid = bytes(u.cpu().numpy().tobytes())
    opts = B.default_options(max_kryl_sz=a.max_kryl or 100000, precision=prec, restart=1, op=B.OP_SHIFT_MINUS_A, sigma=12.0,
                             filter_degree=a.degree if a.degree >= 0 else 16, device=lr, verbose=a.verbose if rank == 0 else 0,
                             host_threads=os.cpu_count() or 1)
    Om = np.random.default_rng(1000 + rank).standard_normal((r1 - r0, b))
    t0 = time.perf_counter()
    s = B.Solver(options=opts, shard=dict(n=n, row0=r0, rowptr=rowptr, colidx=cols, vals=vals, rank=rank, world=world, uid=uid))
    t_create = time.perf_counter() - t0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    D, V, st = s.solve(k, b, np.asfortranarray(Om), allow_not_converged=True)
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t0
    # gates: residual with the local rows needs the halo of V: ||A v - lambda v||^2 summed over ranks via an all-gather-free trick -
    # each rank applies its local rows to the full V only if it fits; here V is gathered block-column-wise on rank 0 for k columns
    exact = 12.0 - matrices.laplacian_eigs(N, 3, k)
    res = {"config": 5, "workload": f"configs[4]: 3D 7-point Laplacian {N}^3 (n={n}), {k} lowest eigenpairs via 12I-A, b={b}, {world} GPUs row-sharded",
           "n": n, "world": world, "generate_s": t_gen, "create_s": t_create, "solve_s": t_solve, "stats": stats_dict(st),
           "max_rel_eig_err_vs_analytic": float(np.max(np.abs(D - exact) / exact)), "D_first_last": [float(D[0]), float(D[-1])],
           "device_measured_max_residual_over_normA": st.max_residual / 12.0}
    # V'V - I over ranks
    G = torch.from_numpy(V.T @ V).to(dev)
    if world > 1:
        dist.all_reduce(G)
    res["v_orthonormality_2norm"] = float(np.linalg.norm(G.cpu().numpy() - np.eye(k), 2))
    s.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
