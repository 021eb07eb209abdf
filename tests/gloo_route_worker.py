"""One rank of the CPU (gloo) check of the per-step routing layer (hymd_b200.pm.ParticleMesh.route_in /
route_back, HYMD_B200_AUTO_ROUTE): the two methods are written against ParticleMesh.migrate only, so
here migrate is replaced by a gloo all-to-all with the semantics include/hymd_b200.h documents for
hymd_migrate_plan / hymd_migrate_apply (destination = slab of the routing position's x; rows that stay
first, in their original order, then the arrivals from rank 0, 1, ...).  Checked: every particle reaches
the owner of its slab, and a per-particle result computed on the working set comes back to the right
particle in the caller's order.  Launched by tests/test_gloo_md.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    from hymd_b200.pm import ParticleMesh

    class GlooPM(ParticleMesh):
        def __init__(self):                       # no CUDA context: only what route_in / route_back touch
            self.rank, self.world_size = rank, P
            self.device = torch.device("cpu")
            self.dtype = torch.float64
            self.BoxSize = np.array([8.0, 5.0, 6.0])
            self._route = None

        def __del__(self):
            pass

        def migrate(self, positions, *arrays, routing_positions=None):
            arrays = (positions,) + tuple(arrays)
            route = positions if routing_positions is None else routing_positions
            slab = self.BoxSize[0] / P
            x = torch.remainder(torch.as_tensor(route)[:, 0].double(), float(self.BoxSize[0]))
            dest = torch.clamp((x / slab).long(), max=P - 1)
            out = []
            for a in arrays:
                a = torch.as_tensor(a)
                send = [a[dest == q].contiguous() for q in range(P)]
                counts = torch.tensor([len(t) for t in send])
                all_counts = [torch.zeros(P, dtype=torch.long) for _ in range(P)]
                dist.all_gather(all_counts, counts)
                recv = [torch.empty((int(all_counts[q][rank]),) + tuple(a.shape[1:]), dtype=a.dtype) for q in range(P)]
                _all_to_all(recv, send, rank, P)
                out.append(torch.cat([recv[rank]] + [recv[q] for q in range(P) if q != rank]))
            return tuple(out)

    pm = GlooPM()
    rng = np.random.default_rng(100 + rank)
    n = 500 + 37 * rank
    # every rank holds particles from all over the box (guests) -- like beads of straddling molecules
    pos = torch.tensor(rng.random((n, 3)) * pm.BoxSize)
    types = torch.tensor(rng.integers(0, 3, size=n).astype(np.int32))
    charges = torch.tensor(rng.normal(size=n))
    wpos, wtypes, wnone, wq = pm.route_in(pos, types, None, charges)
    assert wnone is None
    slab = pm.BoxSize[0] / P
    assert ((wpos[:, 0] >= rank * slab) & (wpos[:, 0] < (rank + 1) * slab)).all()       # everybody is at home
    total = torch.tensor([len(wpos)])
    dist.all_reduce(total)
    want = torch.tensor([n])
    dist.all_reduce(want)
    assert int(total) == int(want)
    # a "force" that identifies the particle: f(position, type, charge)
    f_w = torch.stack([wpos[:, 0] * 3 + wtypes.double(), wpos[:, 1] - wq, wpos[:, 2] * wq], dim=1)
    f = pm.route_back(f_w)
    ref = torch.stack([pos[:, 0] * 3 + types.double(), pos[:, 1] - charges, pos[:, 2] * charges], dim=1)
    assert f.shape == (n, 3) and torch.equal(f, ref)
    # a second routing of the same positions gives the same working order (what sort() relies on when
    # charges are attached later)
    wpos2, wq2 = pm.route_in(pos, charges)
    assert torch.equal(wpos2, wpos) and torch.equal(wq2, wq)
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()


def _all_to_all(recv, send, rank, P):
    """gloo has no all_to_all for CPU tensors in every build: pairwise isend / irecv."""
    reqs = []
    for q in range(P):
        if q == rank:
            recv[q].copy_(send[q])
            continue
        reqs.append(dist.isend(send[q], q))
        reqs.append(dist.irecv(recv[q], q))
    for r in reqs:
        r.wait()


if __name__ == "__main__":
    main()
