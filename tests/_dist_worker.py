"""world_size-2 worker (gloo, CPU): host-side planning of the row-partitioned path.
The compute stand-ins here are scipy/numpy (test infrastructure); the product's kernels
need a GPU and are covered by tests/test_dist_gpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from krypy_b200 import problems  # noqa: E402
from krypy_b200.dist import HaloPlan, RowPartition, extend_vector, local_rows  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for name, A in (("lap2d", problems.laplace2d(13)), ("convdiff", problems.convdiff2d(11)),
                    ("poisson3d", problems.poisson3d(5))):
        N = A.shape[0]
        part = RowPartition(N, world, rank)
        assert part.block % 32 == 0 and part.block * world >= N
        rows = local_rows(A, part)
        # the direct row-block builder must agree with slicing the global matrix
        if name == "lap2d":
            R2 = problems.laplace2d(13, rows=(part.lo, part.hi))
            assert (abs(R2 - rows)).nnz == 0 and np.array_equal(R2.indices, rows.indices)
        plan = HaloPlan(rows, part)
        x = np.random.default_rng(7).standard_normal(N)
        # exchange: all ranks publish their padded block (what the peer-mapped basis rows hold)
        mine = torch.zeros(part.block, dtype=torch.float64)
        mine[: part.nloc] = torch.from_numpy(x[part.lo:part.hi])
        blocks = [torch.zeros(part.block, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(blocks, mine)
        xb = [b.numpy()[: RowPartition(N, world, r).nloc] for r, b in enumerate(blocks)]
        xe = extend_vector(plan, part, xb)
        y = plan.local_matrix() @ xe
        ref = (A @ x)[part.lo:part.hi]
        assert np.array_equal(y, ref), name          # same entries in the same order: bit-identical
        # halo is minimal: exactly the distinct remote columns
        cols = rows.indices
        remote = np.unique(cols[(cols < part.lo) | (cols >= part.hi)])
        assert plan.nhalo == remote.size and np.array_equal(plan.halo_cols, remote)
        assert np.all(plan.halo_peer != rank) and np.all(plan.halo_off < part.block)
        # global reductions: partial dot + all-reduce == global dot (to round-off)
        w = np.random.default_rng(8).standard_normal(N)
        part_dot = torch.tensor([float(w[part.lo:part.hi] @ x[part.lo:part.hi])], dtype=torch.float64)
        dist.all_reduce(part_dot)
        assert abs(part_dot.item() - w @ x) < 1e-12 * np.abs(w @ x) + 1e-12
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
