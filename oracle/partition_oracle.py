"""NumPy twin of the row partition + SpMM halo plan (TEST INFRASTRUCTURE; bit-exact checker of
csrc/partition.cpp).  The reference has no multi-GPU path; this is new design."""
import numpy as np


def partition_rows(n: int, world: int) -> np.ndarray:
    return np.array([(n * p) // world for p in range(world + 1)], dtype=np.int64)


def halo_plan(n, world, row_starts, rank, rowptr, colidx):
    r0, r1 = int(row_starts[rank]), int(row_starts[rank + 1])
    colidx = np.asarray(colidx, dtype=np.int64)
    ext = np.unique(colidx[(colidx < r0) | (colidx >= r1)])
    owner_ptr = np.searchsorted(ext, row_starts, side="left").astype(np.int64)
    owner_ptr[0] = 0
    loc = np.where((colidx >= r0) & (colidx < r1), colidx - r0, (r1 - r0) + np.searchsorted(ext, colidx)).astype(np.int32)
    return ext.astype(np.int64), owner_ptr, loc
