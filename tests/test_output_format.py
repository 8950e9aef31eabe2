"""Output fidelity: the text files are what the gfortran build of the reference writes.

The reference uses list-directed output (`write(unit,*)`, 3D-iso :1219-1229, :1254-1256, :1349), whose format is the
compiler's.  gfortran (the reference Makefile's compiler) formats REAL(4) as 1PG16.9E2 and REAL(8) as 1PG25.17E3 with 9 /
17 significant digits in F and E editing alike, INTEGER(4) in 11 columns, one leading blank per record and one blank
before each further item.  No Fortran compiler exists in this image, so the expected strings below are hand-derived from
those rules (the decimal expansions of the single-precision values are exact arithmetic, e.g. sngl(1.6d-3) = 1.59999996E-03)."""
import os

import numpy as np
import pytest

from seismic_cpml_b200 import lib as L


@pytest.mark.parametrize("value,kind,text", [
    (1.0, 4, "  1.00000000    "),
    (0.0, 4, "  0.00000000    "),
    (1.6e-3, 4, "  1.59999996E-03"),
    (-2.5, 4, " -2.50000000    "),
    (0.1, 4, " 0.100000001    "),
    (123456.7, 4, "  123456.703    "),
    (1.0e10, 4, "  1.00000000E+10"),
    (-3.2e-7, 4, " -3.19999998E-07"),
    (999999999.0, 4, "  1.00000000E+09"),          # rounds to 10 digits: E editing
    (123456789.0, 4, "  123456792.    "),
    (1.0, 8, "  1.0000000000000000     "),
    (1.0e-3, 8, "  1.0000000000000000E-003"),
    (0.0, 8, "  0.0000000000000000     "),
    (0.1, 8, " 0.10000000000000001     "),
    (-2.5, 8, " -2.5000000000000000     "),
    (1.0e25, 8, "  1.0000000000000001E+025"),
    (0.25, 8, " 0.25000000000000000     "),
])
def test_gfortran_list_directed_reals(value, kind, text):
    assert L.host_format_real(value, kind) == text
    assert len(text) == (16 if kind == 4 else 25)


def test_seismogram_energy_and_timestamp_files_are_list_directed(tmp_path):
    lib = L.load()
    nt, nrec, dt = 4, 1, 1.6e-3
    sx = np.array([[0.0, 1.5e-5, -0.25, 3.0]])
    sy = sx * 2
    assert lib.cpml_host_write_seismograms(str(tmp_path).encode(), L._d(sx), L._d(sy), nt, nrec, dt) == 0
    lines = open(tmp_path / "Vx_file_001.dat").read().split("\n")
    # write(11,*) sngl(dble(it-1)*DELTAT),' ',sngl(sisvx(it,irec))   (:1349)
    assert lines[0] == "   0.00000000         0.00000000    "
    assert lines[1] == "   1.59999996E-03     1.49999996E-05"
    assert lines[2] == "   3.19999992E-03   -0.250000000    "
    assert lines[3] == "   4.80000023E-03     3.00000000    " and lines[4] == ""
    e = np.array([0.0, 2.0e8, 1.0 / 3.0, 12.5])
    assert lib.cpml_host_write_energy_3d(str(tmp_path / "energy.dat").encode(), L._d(e), nt, dt) == 0
    lines = open(tmp_path / "energy.dat").read().split("\n")
    # write(20,*) sngl(dble(it-1)*DELTAT),total_energy(it)   (:1254-1256)
    assert lines[0] == "   0.00000000       0.0000000000000000     "
    assert lines[1] == "   1.59999996E-03   200000000.00000000     "
    assert lines[2] == "   3.19999992E-03  0.33333333333333331     "
    assert lib.cpml_host_write_energy_2d(str(tmp_path / "energy2.dat").encode(), L._d(e), L._d(e), nt, dt) == 0
    assert open(tmp_path / "energy2.dat").read().split("\n")[3] == "   4.80000023E-03   12.5000000       12.5000000       25.0000000    "
    assert lib.cpml_host_write_timestamp(str(tmp_path).encode(), 100, dt, 0.5, 2.0e8, 3725.5) == 0
    t = open(tmp_path / "timestamp000100").read().split("\n")
    assert t[0] == " Time step #          100"
    assert t[1] == " Time:   0.158399999      seconds"
    assert t[2] == " Max norm velocity vector V (m/s) =   0.50000000000000000     "
    assert t[3] == " Total energy =    200000000.00000000     "
    assert t[5] == " Elapsed time in hh:mm:ss =    1 h 02 m 05 s"


def test_gnuplot_scripts_match_the_reference_text(tmp_path):
    lib = L.load()
    assert lib.cpml_host_write_gnuplot_scripts(str(tmp_path).encode(), 0) == 0
    pe = open(tmp_path / "plot_energy").read().split("\n")
    assert pe[0] == " # set term x11" and pe[2] == "" and pe[6] == ' set output "CPML3D_total_energy_semilog.eps"'
    assert pe[8] == " plot \"energy.dat\" t 'Total energy' w l lc 1"
    pg = open(tmp_path / "plotgnu").read()
    assert pg.startswith(" set term x11\n # set term postscript landscape monochrome dashed \"Helvetica\" 22\n\n")
    for rec in ("001", "002"):
        for c in ("Vx", "Vy", "Vz"):                   # :1275-1303 (the Vz files are this library's extension, quirk B7)
            assert f' set output "v_sigma_{c}_receiver_{rec}.eps"\n plot "{c}_file_{rec}.dat" t \'{c} C-PML\' w l lc 1\n' in pg
    assert not os.path.exists(tmp_path / "plot_comparison")
    assert lib.cpml_host_write_gnuplot_scripts(str(tmp_path).encode(), 1) == 0
    assert "Vz_file" not in open(tmp_path / "plotgnu").read()
    assert "../collino/energy.dat" in open(tmp_path / "plot_comparison").read()       # 2D-2nd :773-774
    assert "us 1:3  t 'Ep' w l lc 3" in open(tmp_path / "plot_energy").read()
    assert lib.cpml_host_write_gnuplot_scripts(str(tmp_path).encode(), 2) != 0
