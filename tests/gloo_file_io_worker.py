"""One rank of the CPU (gloo) check of hymd_b200.file_io over real process ranks: each rank reads its share of the
fixture (distribute_input), writes its own file through the in-memory h5py stand-in, and checks its rows, the bond
offsets and the rank-summed observables against the golden tree the reference's own file_io.py wrote
(tests/golden/file_io_golden.npz).  Covers the torch.distributed path of hymd_b200._world (the all-reduces behind
the bond offsets and the momenta).  Launched by tests/test_file_io.py through torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    import fake_h5
    from hymd_b200 import file_io as fio
    from test_file_io import G, Cfg
    fio.set_backend(fake_h5)
    ind, pos, vel, mol = G["in|indices"], G["in|positions"], G["in|velocities"], G["in|molecules"]
    b1, b2 = G["in|bonds_2_atom1"], G["in|bonds_2_atom2"]
    cfg = Cfg(str(G["A|config_str"]), n_steps=100, n_print=1, mass=72.0)
    rr, flag = fio.distribute_input({"molecules": mol, "indices": ind}, rank, P, cfg.n_particles, 6)
    lo, hi = rr[0], rr[-1] + 1
    keep = (b1 >= lo) & (b1 < hi)
    out = fio.OutDataset("/tmp", cfg)
    assert out.file.filename.endswith(f"sim.hdf5-{rank:6d}-of-{P:6d}"), out.file.filename
    fio.store_static(out, rr, G["in|names"][lo:hi], G["in|types"][lo:hi], ind[lo:hi], cfg, b1[keep] - lo, b2[keep] - lo,
                     molecules=mol[lo:hi], charges=True, plumed_out=True)
    t = torch.from_numpy
    fio.store_data(out, 0, 0, t(ind[lo:hi].copy()), t(pos[lo:hi].copy()), t(vel[lo:hi].copy()), t(pos[lo:hi].copy()),
                   np.array([10.0, 10.0, 10.0]), 300., 1., 1., 2., 3., 4., 5., 6., 7., 0.02, cfg, charge_out=True,
                   plumed_out=True)
    f = out.file
    ref_pos = G["A|/particles/all/position/value"][0]
    got = f["particles/all/position/value"][0]
    assert np.array_equal(got[lo:hi], ref_pos[lo:hi]) and not got[:lo].any() and not got[hi:].any()
    assert np.array_equal(f["particles/all/species"][lo:hi], G["A|/particles/all/species"][lo:hi])
    assert np.array_equal(f["parameters/vmd_structure/resid"][lo:hi], G["A|/parameters/vmd_structure/resid"][lo:hi])
    # bonds: every rank fills its own stretch of the global arrays; together they are the reference's
    mine = np.stack([f["parameters/vmd_structure/bond_from"][:], f["parameters/vmd_structure/bond_to"][:]]).astype(np.int64)
    total = torch.from_numpy(mine.copy())
    dist.all_reduce(total)
    want = np.stack([G["A|/parameters/vmd_structure/bond_from"], G["A|/parameters/vmd_structure/bond_to"]])
    assert np.array_equal(total.numpy(), want), (total.numpy(), want)
    nz = np.nonzero(mine[0])[0]
    assert len(nz) == int(keep.sum()) and (len(nz) == 0 or nz[-1] - nz[0] + 1 == len(nz))
    for name in ("total_momentum", "angular_momentum", "torque"):
        np.testing.assert_allclose(f[f"observables/{name}/value"][0], G[f"A|/observables/{name}/value"][0], rtol=2e-6,
                                   atol=1e-8)
    assert out.last_log == str(G["A|log"][0])
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
