"""rbl_b200 - B200-native randomized block Lanczos behind the reference's RBL / RBL_gpu entry points.

Host-side mirror (Python, ctypes) of the Julia interface of Iasonaspg/GPU-Randomized-Block-Lanczos:
``RBL_gpu(A, k, b) -> (D, V)`` (Julia/RBL_gpu.jl:205-221).  All arithmetic happens in the C-ABI library
``lib/librbl_b200.so`` (hand-written sm_100a kernels + C++ host driver); there is no CPU fallback.
"""
from .binding import (RblError, RblOptions, RblStats, Solver, lib, lib_path, load_library,  # noqa: F401
                      band_eig_topk, band_count_below, Checker, halo_plan, partition_rows, microbench,
                      k_spmm, k_gram, k_block_qr, k_reorth, k_ritz, load_matrix_market, load_matrix)
from .rbl import RBL, RBL_gpu, rbl_solve_sharded  # noqa: F401
