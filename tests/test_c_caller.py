"""A plain C program written against include/roki_fd/rkfd_b200.h in the style of the reference's example callers
(reference example/chain/boxdrop_test.c): it must compile as C99 with gcc, link against librokifd_b200.so only, and -
on a GPU box - reproduce the oracle's trajectory of the same world for each of the three solver plugins."""
import os
import subprocess

import numpy as np
import pytest

import rokifd_b200  # noqa: F401
from rokifd_b200 import chains as ch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_callers", "boxdrop.c")
LIBDIR = os.path.join(ROOT, "roki-fd_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def build(tmp_path):
    exe = str(tmp_path / "boxdrop")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
                           "-L" + LIBDIR, "-lrokifd_b200", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


def test_c_caller_compiles_and_links_without_cuda_headers(tmp_path):
    exe = build(tmp_path)
    needed = subprocess.check_output(["readelf", "-d", exe], text=True)
    assert "librokifd_b200.so" in needed and "libcuda" not in needed and "libcudart" not in needed
    src = open(SRC).read()
    assert "cuda" not in src.lower().replace("no cuda in sight", "")


@pytest.mark.gpu
@pytest.mark.parametrize("solver,contacts,z0,cube", [("Volume", "-", 0.049, "cube.ztk"), ("Vert", "-", 0.049, "cube.ztk"),
                                                     ("MLCP", "contacts.ztk", 0.049, "cube.ztk"),
                                                     ("Volume", "-", 0.049, "cube_poly.ztk")])     # polyhedron-described box
def test_c_caller_matches_oracle(tmp_path, oracle, solver, contacts, z0, cube):
    exe = build(tmp_path)
    nsteps = 60
    args = [exe, os.path.join(GOLD, cube), os.path.join(GOLD, "rigidfloor.ztk"),
            "-" if contacts == "-" else os.path.join(GOLD, contacts), solver, str(nsteps), repr(z0)]
    out = subprocess.check_output(args, text=True)
    final = [l for l in out.splitlines() if l.startswith("final:")][0].split()
    got = np.array([float(x) for x in final[1:7]])
    t = float(final[7].split("=")[1])
    assert abs(t - nsteps * 0.001) < 1e-12
    ci = [] if contacts == "-" else [ch.ContactInfo("soft", "body", "elastic", E=1000.0, V=10.0, SF=0.5, KF=0.3),
                                     ch.ContactInfo("ground", "body", "rigid", K=1000.0, L=0.0001, SF=0.6, KF=0.4)]
    w = ch.World(chains=[ch.box(), ch.floor()], contact_info=ci, solver=solver)
    e = oracle.OracleWorld(w).env()
    q = np.zeros(6); q[2] = z0; q[3] = 0.05; q[4] = 0.02
    qd = np.zeros(6); qd[0] = 0.3
    e.set_state(q, qd); e.set_motor_input(np.zeros(w.nl)); e.update_init()
    for _ in range(nsteps):
        e.update()
    ref = e.get_state()[0]
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), (got, ref)
