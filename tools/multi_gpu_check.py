"""Run under torchrun (one rank per GPU): row-sharded solve == single-GPU solve == analytic spectrum.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import rbl_b200
from rbl_b200 import binding as B
from oracle import matrices


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    def fresh_uid():
        """One ncclUniqueId per communicator: created on rank 0, broadcast by the host plumbing."""
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert rbl_b200.lib().rbl_nccl_unique_id(buf) == 0, rbl_b200.lib().rbl_last_error()
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())
    ok = True
    for (name, L, sigma, k, b, prec) in (
            ("lap3d-20 fp64", matrices.laplacian_3d(20), 12.0, 20, 16, "fp64"),
            ("lap3d-24 mixed", matrices.laplacian_3d(24), 12.0, 30, 16, "mixed"),
            ("er-4000 fp64", matrices.erdos_renyi_sym(4000, 16, seed=1), None, 8, 8, "fp64")):
        L = L.tocsr(); L.sort_indices()
        n = L.shape[0]
        Om = np.random.default_rng(5).standard_normal((n, b))
        rs = rbl_b200.partition_rows(n, world)
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        kw = dict(max_kryl_sz=4000, precision=B.PRECISION_MIXED if prec == "mixed" else B.PRECISION_FP64,
                  op=B.OP_SHIFT_MINUS_A if sigma is not None else B.OP_A, sigma=sigma or 0.0, device=lr)
        D, V, st = rbl_b200.rbl_solve_sharded(L[r0:r1], n, r0, k, b, rank=rank, world=world, uid=fresh_uid(),
                                              Omega_local=Om[r0:r1], **kw)
        # gather V on rank 0 and compare with the single-GPU solve
        Vt = torch.from_numpy(np.ascontiguousarray(V)).to(dev)
        parts = [torch.empty((int(rs[p + 1] - rs[p]), k), dtype=torch.float64, device=dev) for p in range(world)]
        dist.all_gather(parts, Vt)
        if rank == 0:
            Vall = torch.cat(parts).cpu().numpy()
            A = matrices.shifted(L, sigma) if sigma is not None else L
            res = np.max(np.linalg.norm(A @ Vall - Vall * D[None, :], axis=0)) / np.max(np.abs(D))
            with B.Solver(L, options=B.default_options(**kw)) as s1:
                D1, V1, st1 = s1.solve(k, b, Om)
            rel = np.max(np.abs(D - D1) / np.abs(D1))
            good = rel < 1e-8 and res < 1e-6 and st.converged
            ok &= good
            print(f"[{name}] world={world} iterations {st.iterations} (1-GPU {st1.iterations}) "
                  f"max rel eig diff vs 1-GPU {rel:.2e}, residual {res:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
