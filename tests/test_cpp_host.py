"""Builds and runs the C++ host-template test (tests/cpp/host_api_test.cpp) that mirrors the reference's
own Boost tests for the hot-path entities on top of the C ABI."""
import os
import subprocess

import pytest

from oracle import fields, fri, hashes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_test")


def _build():
    from crypto3_zk_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")
    deps = [src, os.path.join(ROOT, "crypto3_zk_b200", "host", "zkb_crypto3.hpp"), os.path.join(ROOT, "include", "zkb200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        lib = os.path.join(ROOT, "crypto3_zk_b200")
        subprocess.check_call(["g++", "-O2", "-std=c++17", src, "-o", EXE, "-L" + lib, "-lzkb200", "-Wl,-rpath," + lib])
    return EXE


def test_host_templates_compile_and_link():
    exe = _build()
    out = subprocess.run([exe, "compile-only"], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "compiled" in out


@pytest.mark.gpu
def test_host_templates_on_gpu():
    exe = _build()
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout
    root = [l.split()[1] for l in r.stdout.splitlines() if l.startswith("ROOT ")][0]
    F = fields.PALLAS_FP
    polys = [[(p + 1) * 1000 + i for i in range(16)] for p in range(3)]
    levels, _ = fri.precommit(polys, F, 64, 2, hashes.keccak256)
    assert root == levels[-1][0].hex()
