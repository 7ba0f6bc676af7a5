"""Restart files in the reference's format (write_restart_files.f90 / readfiles.f90): the Python
module and the C++ host mirror read and write the same bytes.  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest

from freecappuccino_b200 import restart

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "restart_copy")


def sample(n=37, F=80, nt=61, seed=3):
    rng = np.random.default_rng(seed)
    f = {k: rng.standard_normal(nt) for k in ("u", "v", "w", "p", "vis", "uo", "vo", "wo")}
    f["flmass"] = rng.standard_normal(F)
    for k in ("te", "ed", "teo", "edo"):
        f[k] = np.zeros(nt)       # allocated unconditionally by the reference (allocate.f90:75-80)
    return f                      # t and the Reynolds stresses: not allocated in a laminar run -> empty records


def test_layout_is_fortran_unformatted_sequential(tmp_path):
    path = os.path.join(str(tmp_path), "restart")
    f = sample()
    restart.write_restart(path, 17, 0.125, f)
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 12                      # write(3) itime,time: 4 + 8 bytes
    assert struct.unpack("<id", raw[4:16]) == (17, 0.125)
    assert struct.unpack("<i", raw[16:20])[0] == 12
    assert struct.unpack("<i", raw[20:24])[0] == 8 * f["flmass"].size  # write(3) flmass
    back = restart.read_restart(path)
    assert back["itime"] == 17 and back["time"] == 0.125
    for k in restart.ORDER:
        assert np.array_equal(back[k], f.get(k, np.zeros(0))), k
    assert back["t"].size == 0 and back["uu"].size == 0


def test_const_mflux_adds_the_gradpcmf_record(tmp_path):
    path = os.path.join(str(tmp_path), "restart")
    restart.write_restart(path, 3, 1.5, sample(), const_mflux=True, gradpcmf=0.37)
    back = restart.read_restart(path, const_mflux=True)
    assert back["gradpcmf"] == 0.37
    with pytest.raises(ValueError):
        restart.read_restart(path, const_mflux=False)


def test_cpp_mirror_reads_and_writes_the_same_bytes(tmp_path):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "restart_copy"])
    n, F, nt = 37, 80, 61
    src, dst = os.path.join(str(tmp_path), "restart.in"), os.path.join(str(tmp_path), "restart.out")
    restart.write_restart(src, 42, 2.5e-3, sample(n, F, nt))
    out = subprocess.run([EXE, str(n), str(F), str(nt), src, dst], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split()[:2] == ["itime", "42"]
    assert open(src, "rb").read() == open(dst, "rb").read()
    # a file of the wrong mesh is rejected
    bad = subprocess.run([EXE, str(n), str(F), str(nt + 1), src, dst], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "size of" in bad.stderr
