"""Opening recursion: runs tests/cpp/open_test.cpp — the unmodified reference and the host mirror in ONE process, same
libc RNG state — in both builds: against the CPU emulation of the C ABI (host logic, no GPU needed) and against the CUDA library."""
import os
import subprocess

import pytest

from helpers import ROOT

EMUL = os.path.join(ROOT, "oracle", "_ref", "open_test_emul")
GPU = os.path.join(ROOT, "oracle", "_ref", "open_test")


def _run(binary, *args, timeout=300):
    p = subprocess.run([binary, *args], capture_output=True, text=True, timeout=timeout)
    tail = "\n".join(l for l in p.stdout.splitlines() if l.startswith(("ok:", "FAIL", "OPEN", " ")))
    print(tail[-6000:])
    assert p.returncode == 0, tail[-3000:] + p.stderr[-2000:]
    assert "OPEN: all identical" in p.stdout


@pytest.mark.skipif(not os.path.exists(EMUL), reason="oracle/_ref/open_test_emul not prebuilt (needs /root/reference at build time)")
def test_open_recursion_host_logic_vs_reference():
    _run(EMUL)


@pytest.mark.skipif(not os.path.exists(EMUL), reason="oracle/_ref/open_test_emul not prebuilt (needs /root/reference at build time)")
def test_deep_product_tree_host_logic_vs_reference():
    """layers > distance: batched streaming sumchecks + commit_layers / open_layers (sumcheck.cpp:983-1011, 1871-1911)."""
    _run(EMUL, "deep")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPU), reason="oracle/_ref/open_test not prebuilt")
def test_deep_product_tree_gpu_vs_reference():
    _run(GPU, "deep")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPU), reason="oracle/_ref/open_test not prebuilt (needs /root/reference at build time)")
def test_open_recursion_gpu_vs_reference():
    _run(GPU)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPU), reason="oracle/_ref/open_test not prebuilt")
def test_open_standard_2e20_ps_kat():
    """test_PC(2^20, 4, 32): commit + open, ps == 3839.078125 KB, identical to the reference run in the same process."""
    _run(GPU, "big")
