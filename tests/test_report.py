"""The printout contract (reference tiny_mc.c:37-40,55-66): given the reference's own tallies
our formatter must print the reference's own bytes."""
import ctypes as C
import json
import os
import tempfile

import numpy as np
from conftest import GOLDEN, ROOT


def render(heat, heat2, photons, shells=101, mps=50.0, mu_s=20.0, mu_a=2.0, backend="x"):
    lib = C.CDLL(str(ROOT / "tiny_mc_b200" / "lib" / "libtmc_report.so"))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lib.tmc_report_heading.argtypes = [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_uint64]
    lib.tmc_report_timing.argtypes = [C.c_void_p, C.c_double, C.c_uint64]
    lib.tmc_report_table.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.txt")
        f = libc.fopen(path.encode(), b"w")
        lib.tmc_report_heading(f, backend.encode(), mu_s, mu_a, photons)
        lib.tmc_report_timing(f, 0.25, photons)
        lib.tmc_report_table(f, shells, mps, photons, heat.ctypes.data, heat2.ctypes.data)
        libc.fclose(f)
        return open(path).read().splitlines()


def test_table_is_byte_identical_to_reference_headless():
    gold = (GOLDEN / "headless_asshipped.txt").read_text().splitlines()
    g = json.loads((GOLDEN / "ref_float_tallies.json").read_text())["headless"]
    heat = np.array(g["heat_bits"], np.uint32).view(np.float32)
    heat2 = np.array(g["heat2_bits"], np.uint32).view(np.float32)
    mine = render(heat, heat2, g["photons"], backend=gold[2][2:])
    assert len(mine) == len(gold) == 3 + 4 + 2 + 2 + 100 + 1
    for i, (a, b) in enumerate(zip(mine, gold)):
        if "seconds" in b or "K photons per second" in b:      # timing lines: same format, other numbers
            assert a.split()[0] == "#" and a.split()[2:] == b.split()[2:]
            continue
        assert a == b, (i, a, b)
    assert mine[7] == "# 0.250000 seconds" and mine[8] == "# 131.072000 K photons per second"


def test_photon_count_is_64_bit_with_the_reference_field_width():
    heat = np.ones(101, np.float32)
    lines = render(heat, heat, 1 << 32)
    assert lines[5] == "# Photons    = 4294967296"
    lines = render(heat, heat, 32768)
    assert lines[5] == "# Photons    =    32768"
