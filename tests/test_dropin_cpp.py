"""Runs the C++ drop-in binary (tests/cpp/dropin_test.cpp): the unmodified reference and the host mirror
(hobbit_b200/host) are linked into ONE process and called with the same arguments and the same libc RNG state;
commit_standard / open_standard front half / Elastic commit / the sumcheck provers must agree field by field."""
import os
import subprocess

import pytest

from helpers import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_test not prebuilt (needs /root/reference at build time)")
def test_dropin_cpp():
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "DROPIN: all identical" in p.stdout
