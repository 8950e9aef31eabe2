"""Generates the golden vectors under tests/golden/ from the C oracle (golden build).

PROVENANCE: these are outputs of OUR restatement (oracle/cpml_oracle.c, gcc -O2
-ffp-contract=off), not of the Fortran reference, which cannot be compiled in this
image (no Fortran compiler).  They pin the oracle against accidental change and give
the GPU tests fixed targets; the independent numpy restatement is checked against
them in tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as O  # noqa: E402
import refcfg  # noqa: E402


def main():
    O.build()
    # cfg1: 2-D second order exactly as shipped
    o = O.run_2d(**refcfg.cfg2d(2))
    np.savez_compressed(os.path.join(HERE, "cpml2d_second_default.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                        energy_kinetic=o["energy_kinetic"], energy_potential=o["energy_potential"])
    # 2-D fourth order exactly as shipped
    o = O.run_2d(**refcfg.cfg2d(4))
    np.savez_compressed(os.path.join(HERE, "cpml2d_fourth_default.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                        energy_kinetic=o["energy_kinetic"], energy_potential=o["energy_potential"])
    # 2-D heterogeneous (layered) medium, both orders, reduced grid
    for order in (2, 4):
        o = O.run_2d(**refcfg.cfg2d(order, nx=83, ny=131, nstep=600, npml=8, material="layered",
                                    ydeb=600.0, yfin=200.0))
        np.savez_compressed(os.path.join(HERE, f"cpml2d_layered_order{order}.npz"), sisvx=o["sisvx"],
                            sisvy=o["sisvy"], energy_kinetic=o["energy_kinetic"],
                            energy_potential=o["energy_potential"])
    # 3-D isotropic, reduced grid, two emulated slabs
    o = O.run_3d_iso(**refcfg.cfg3d(), nproc=2, want_planes=True)
    np.savez_compressed(os.path.join(HERE, "cpml3d_iso_small.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                        total_energy=o["total_energy"], plane_vx=o["plane_vx"], plane_vy=o["plane_vy"])
    # 3-D with K_MAX_PML = 3 (exercises the /K path, quirk-free)
    o = O.run_3d_iso(**refcfg.cfg3d(nx=30, ny=34, nz=32, nstep=120, npml=5, k_max=3.0), nproc=2)
    np.savez_compressed(os.path.join(HERE, "cpml3d_iso_kmax3.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                        total_energy=o["total_energy"])
    # 3-D viscoelastic (fourth order, N_SLS = 2, Carcione 1993 relaxation times), reduced grid,
    # 4 emulated slabs like the reference's NPROC = 4 (quirk B6 applies at the 3 interfaces)
    o = O.run_3d_visco(**refcfg.cfgv3d(), nproc=4)
    np.savez_compressed(os.path.join(HERE, "cpml3d_visco_small.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                        total_energy=o["total_energy"], energy_kinetic=o["energy_kinetic"],
                        energy_potential=o["energy_potential"])
    # 2-D viscoelastic (N_SLS = 3), both orders, layered medium, reduced grid
    for order in (2, 4):
        o = O.run_2d_visco(**refcfg.cfgv2d(order=order, material="layered", nstep=300))
        np.savez_compressed(os.path.join(HERE, f"cpml2d_visco_order{order}.npz"), sisvx=o["sisvx"], sisvy=o["sisvy"],
                            sispressure=o["sispressure"])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
