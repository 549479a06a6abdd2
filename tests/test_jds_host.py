"""Chunk-local JDS layout of the lane-per-row SpMV path (abip_b200/csrc/order_host.h: jds_sort_chunks; device side:
lp_device.cuh, Csr::jds).  tests/tools/jds_check.cpp emulates a warp with the device's position arithmetic on the host and
compares the row sums bit for bit with a plain CSR product (replaces the inner loop of the reference's _accum_by_Atrans,
linsys/common.c:598-639, for matrices of short rows)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_jds_chunks_match_device_addressing(tmp_path):
    exe = str(tmp_path / "jds_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "tools", "jds_check.cpp"), "-o", exe],
                   check=True)
    for rows, maxlen, seed in ((5000, 12, 1), (3000, 5, 2), (700, 40, 3), (64, 1, 4), (1, 7, 5), (4097, 16, 6)):
        r = subprocess.run([exe, str(rows), str(maxlen), str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, (rows, maxlen, seed, r.stdout)
        assert r.stdout.startswith("ok")
