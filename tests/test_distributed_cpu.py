"""world_size-2 gloo test (CPU): the row partition + halo plan drive a distributed SpMM and the small
all-reduces exactly as the NCCL path does (same plan arrays, same exchange pattern), and reproduce the
single-process result.  The device kernels are not involved (no GPU here); this covers the host-side logic
of SURVEY.md 8(e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.sparse as sp
    import rbl_b200
    from oracle import matrices
    A = matrices.laplacian_3d(10).tocsr()
    A.sort_indices()
    n, b = A.shape[0], 4
    Q = np.random.default_rng(0).standard_normal((n, b))
    rs = rbl_b200.partition_rows(n, world)
    r0, r1 = int(rs[rank]), int(rs[rank + 1])
    Al = A[r0:r1]
    halo, optr, loc = rbl_b200.halo_plan(n, world, rs, rank, Al.indptr, Al.indices)
    # every rank tells each owner which rows it needs (same handshake as handle_create)
    need = [halo[optr[p]:optr[p + 1]] for p in range(world)]
    counts = torch.tensor([len(x) for x in need], dtype=torch.int64)
    allc = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    give = [None] * world
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        give[p] = torch.zeros(int(allc[p][rank]), dtype=torch.int64)
        if len(need[p]):
            reqs.append(dist.isend(torch.from_numpy(need[p].copy()), p))
        if give[p].numel():
            reqs.append(dist.irecv(give[p], p))
    for r in reqs:
        r.wait()
    # halo exchange of Q rows, then the local SpMM on [own | halo]
    Ql = Q[r0:r1]
    recv = {p: torch.zeros((int(optr[p + 1] - optr[p]), b), dtype=torch.float64) for p in range(world) if p != rank}
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        if give[p].numel():
            reqs.append(dist.isend(torch.from_numpy(Ql[give[p].numpy() - r0].copy()), p))
        if recv[p].numel():
            reqs.append(dist.irecv(recv[p], p))
    for r in reqs:
        r.wait()
    Qext = np.vstack([Ql] + [recv[p].numpy() for p in range(world) if p != rank])
    # halo rows are ordered by owner, i.e. by global index (sorted halo list)
    Aloc = sp.csr_matrix((Al.data, loc, Al.indptr), shape=(r1 - r0, Qext.shape[0]))
    Ul = Aloc @ Qext
    # Gram all-reduce (A_i = Q' U)
    G = torch.from_numpy(Ql.T @ Ul)
    dist.all_reduce(G)
    np.save(os.path.join(out_dir, f"U{rank}.npy"), Ul)
    np.save(os.path.join(out_dir, f"G{rank}.npy"), G.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_halo_spmm_and_gram(tmp_path, rbl, world):
    sys.path.insert(0, ROOT)
    from oracle import matrices
    port = 29650 + os.getpid() % 200
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    A = matrices.laplacian_3d(10).tocsr()
    Q = np.random.default_rng(0).standard_normal((A.shape[0], 4))
    U = A @ Q
    Ucat = np.vstack([np.load(tmp_path / f"U{r}.npy") for r in range(world)])
    assert np.max(np.abs(Ucat - U)) < 1e-13
    for r in range(world):
        assert np.max(np.abs(np.load(tmp_path / f"G{r}.npy") - Q.T @ U)) < 1e-10
