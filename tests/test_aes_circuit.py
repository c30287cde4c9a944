"""8f.4, AES circuit evaluator — CPU-side pins (no GPU):
  * the closed-form records the CUDA kernel computes (hobbit_b200/csrc/aes_circuit.cuh, host+device) evaluated on the host by
    tools/aes_records_check.cu and pushed through hb_trace_push == the gate-by-gate restatement in the C-ABI emulation;
  * the reference-free prove_circuit sequence on the emulation backend (oracle/prove_emul = tools/mlp_prove.cpp) reproduces the
    reference's full-run proof size for `pigeon 5 19 8 1`: Ps = 1135.046875 KB.
The emulation's gate-by-gate evaluator itself is pinned to the unmodified reference by tests/test_circuit_cpp.py (circ_test_emul 19 aes 4 1)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_LIB = os.path.join(ROOT, "oracle", "libhb_emul.so")
PROVE = os.path.join(ROOT, "oracle", "prove_emul")


@pytest.mark.skipif(shutil.which("nvcc") is None or not os.path.exists(EMUL_LIB), reason="needs nvcc (host compile) and oracle/libhb_emul.so")
def test_closed_form_records_equal_gate_by_gate(tmp_path):
    exe = str(tmp_path / "aes_records_check")
    subprocess.run(["nvcc", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tools", "aes_records_check.cu"),
                    "-L" + os.path.join(ROOT, "oracle"), "-lhb_emul", "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")],
                   check=True, capture_output=True, timeout=600, cwd=ROOT)
    for blocks in (1, 2, 17):
        p = subprocess.run([exe, str(blocks)], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and "every derived stream identical" in p.stdout, p.stdout[-1000:] + p.stderr[-1000:]
    for rows in (1, 2, 5, 37, 300):                                  # the SQL range-query circuit (sql_circuit.cuh)
        p = subprocess.run([exe, str(rows), "sql"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and "every derived stream identical" in p.stdout, p.stdout[-1000:] + p.stderr[-1000:]


@pytest.mark.skipif(not os.path.exists(PROVE), reason="oracle/prove_emul not built (__graft_entry__.build())")
def test_reference_free_sql_proof_runs_on_emulation():
    """the SQL flow at a small size (the full-run KAT `pigeon 6 19 17 1` -> 1329.890625 KB is checked on the GPU, tests/test_circuit_cpp.py)"""
    p = subprocess.run([PROVE, "11", "sql", "9", "--reps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ps_kb"] > 0


@pytest.mark.skipif(not os.path.exists(PROVE), reason="oracle/prove_emul not built (__graft_entry__.build())")
def test_reference_free_aes_proof_size_kat_on_emulation():
    p = subprocess.run([PROVE, "19", "aes", "8", "--reps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ps_kb"] == 1135.046875
