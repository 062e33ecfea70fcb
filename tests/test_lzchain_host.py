"""CPU check of the K4 match finder's per-position logic (libagmv_b200/csrc/lzchain_core.h, the state machines the CUDA
kernels drive) against the reference's brute-force search (src/agmv_encode.c:119-143): tests/lzchain_host_check.cpp compiles
that header for the host and compares, for EVERY position of synthetic buffers (random alphabets, byte runs longer than the
window, periodic zones, repeats at the window edge, frames back to back, block-record-like data), the longest match length
and the earliest start among the longest."""
import os
import subprocess
import tempfile

from agmv_testlib import ROOT


def test_chain_logic_matches_brute_force():
    exe = os.path.join(tempfile.mkdtemp(prefix="lzchain_"), "lzchain_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "libagmv_b200", "csrc"),
                    os.path.join(ROOT, "tests", "lzchain_host_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert "all ok" in r.stdout and "MISMATCH" not in r.stdout
