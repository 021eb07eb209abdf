"""One rank of the CPU (gloo) check of the slab decomposition: runs oracle/slab_oracle.py on this
rank's particles and compares with the single-rank oracle evaluated on all particles.
Launched by tests/test_gloo_slabs.py through torch.distributed.run."""
import argparse
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs=3, default=[16, 12, 10])
    ap.add_argument("--particles", type=int, default=3000)
    ap.add_argument("--pme", action="store_true")
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()

    from conftest import make_config
    from oracle import field_oracle as fo
    from oracle.hamiltonian_oracle import OracleHamiltonian
    from oracle.slab_oracle import SlabComm, SlabCycle, owner_of

    rng = np.random.default_rng(77)                       # same stream on every rank
    n = args.particles
    box = np.asarray([4.0, 5.0, 6.0], dtype=np.float32)
    pos = rng.uniform(0, 1, size=(n, 3)) * box
    # particles on the slab faces and on the box faces (ownership and halo edge cases)
    nx = args.mesh[0]
    pos[:P, 0] = np.arange(P) * (nx // P) * float(box[0]) / nx
    pos[P:2 * P, 0] = np.nextafter(pos[:P, 0] + float(box[0]) / P, 0)
    names = [("A", "B", "C")[i % 3] for i in range(n)]
    chi = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623), ("B", "C", 0.3852001771213374)]
    cfg = make_config(names, n, args.mesh, box, chi=chi, dtype=np.float64,
                      coulombtype="PIC_Spectral" if args.pme else None,
                      dielectric_const=80.0 if args.pme else None)
    types = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    q = rng.choice([-1.0, 0.0, 1.0], size=n)
    q -= q.mean()

    # single-rank oracle on everything
    import copy
    c1 = copy.deepcopy(cfg)
    h1 = OracleHamiltonian(c1)
    st = fo.FieldState(c1, np.float64)
    fo.update_field(st, h1, pos, types, c1, compute_potential=True)
    f_ref = fo.compute_field_force(st, pos, types, c1.n_types)
    ef_ref = fo.update_field_force_q(st, h1, q, pos, c1) if args.pme else None
    e_ref = fo.compute_field_and_kinetic_energy(st, h1, np.zeros((n, 3)), c1)[0]

    # this rank's slab
    own = owner_of(pos, args.mesh, box, P)
    counts = np.bincount(own, minlength=P)
    assert counts.sum() == n and (counts > 0).all()
    mine = own == rank
    c2 = copy.deepcopy(cfg)
    cyc = SlabCycle(c2, OracleHamiltonian(c2), SlabComm(dist), np.float64)
    f = cyc.field_forces(pos[mine], types[mine])
    scale = np.abs(f_ref).max()
    err = np.abs(f - f_ref[mine]).max() / scale
    assert err < 1e-11, f"rank {rank}: sharded forces differ from the single-rank oracle: {err:.3e}"
    # filtered densities of the owned planes
    for t in range(cfg.n_types):
        ref = st.phi[t][rank * cyc.nxl:(rank + 1) * cyc.nxl]
        assert np.abs(cyc.phi[t] - ref).max() <= 1e-11 * max(np.abs(st.phi[t]).max(), 1.0)
    e = cyc.field_energy()
    assert abs(e - e_ref) <= 1e-11 * max(abs(e_ref), 1.0), (e, e_ref)
    if args.pme:
        ef = cyc.pme_forces(pos[mine], q[mine])
        errq = np.abs(ef - ef_ref[mine]).max() / np.abs(ef_ref).max()
        assert errq < 1e-11, f"rank {rank}: sharded PME forces differ: {errq:.3e}"
        ref = st.psi[rank * cyc.nxl:(rank + 1) * cyc.nxl]
        assert np.abs(cyc.psi - ref).max() <= 1e-11 * np.abs(st.psi).max()
    dist.barrier()
    if rank == 0:
        print("OK", f"P={P} mesh={args.mesh} max rel force err {err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
