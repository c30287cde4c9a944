"""Circuit streams + the prove_circuit flow on an MLP (BASELINE config 3): tests/cpp/circ_test.cpp runs the unmodified
reference (live producer thread, one circuit per process) next to the host mirror fed with ONE pass of the trace resident in HBM;
every named stream, the witness commitment, the wiring product tree, gate consistency and the opening must agree (ps, RNG state)."""
import os
import subprocess

import pytest

from helpers import ROOT

EMUL = os.path.join(ROOT, "oracle", "_ref", "circ_test_emul")
GPU = os.path.join(ROOT, "oracle", "_ref", "circ_test")


def _run(binary, *args, timeout=300):
    p = subprocess.run([binary, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    tail = "\n".join(l for l in p.stdout.splitlines() if l.startswith(("ok:", "FAIL", "CIRC", " ", "{", "circuit_size")))
    print(tail[-6000:])
    assert p.returncode == 0, tail[-3000:] + p.stderr[-2000:]
    assert "CIRC: all identical" in p.stdout


@pytest.mark.skipif(not os.path.exists(EMUL), reason="oracle/_ref/circ_test_emul not prebuilt (needs /root/reference at build time)")
@pytest.mark.parametrize("args", [(12, 64, 32, 16), (11, 128, 16, 8), (19, "aes", 4, 1), (11, "sql", 9, 1), (14, "pruned", 20, 1, 1)])
def test_mlp_circuit_host_logic_vs_reference(args):
    _run(EMUL, *args)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GPU), reason="oracle/_ref/circ_test not prebuilt (needs /root/reference at build time)")
@pytest.mark.parametrize("args", [(12, 64, 32, 16), (11, 128, 16, 8), (14, 256, 64, 64, 16), (19, "aes", 4, 1), (11, "sql", 9, 1), (14, "pruned", 20, 1, 1)])
def test_mlp_circuit_gpu_vs_reference(args):
    _run(GPU, *args)


@pytest.mark.gpu
def test_mlp_prove_reference_free_ps_kat():
    """Reference-free flow (GPU evaluator -> streams -> commit, product tree, gate consistency, open) at the libc position of
    `pigeon 9 18 18 1 4 1024 256 256 16`: the proof-size counter must be the reference's 559.000000 KB (SURVEY §9 full-run KAT)."""
    import json
    binary = os.path.join(ROOT, "hobbit_b200", "mlp_prove")
    if not os.path.exists(binary):
        pytest.skip("hobbit_b200/mlp_prove not built")
    p = subprocess.run([binary, "18", "1024", "256", "256", "16", "--reps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(line)
    assert d["ps_kb"] == 559.0


@pytest.mark.gpu
def test_aes_prove_reference_free_ps_kat():
    """Same for the AES circuit with lookups (`pigeon 5 19 8 1`): GPU evaluator -> witness / wiring / lookup streams -> both commitments,
    both product trees, prove_gate_consistency_lookups, both opens: Ps must be the reference's 1135.046875 KB."""
    import json
    binary = os.path.join(ROOT, "hobbit_b200", "mlp_prove")
    if not os.path.exists(binary):
        pytest.skip("hobbit_b200/mlp_prove not built")
    p = subprocess.run([binary, "19", "aes", "8", "--reps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(line)
    assert d["ps_kb"] == 1135.046875


@pytest.mark.gpu
def test_sql_prove_reference_free_ps_kat():
    """Same for the SQL range query (`pigeon 6 19 17 1`, 2^17 rows, 2^22 gates): Ps must be the reference's 1329.890625 KB."""
    import json
    binary = os.path.join(ROOT, "hobbit_b200", "mlp_prove")
    if not os.path.exists(binary):
        pytest.skip("hobbit_b200/mlp_prove not built")
    p = subprocess.run([binary, "19", "sql", "17", "--reps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(line)
    assert d["ps_kb"] == 1329.890625


@pytest.mark.gpu
def test_prove_with_resident_levels_same_ps():
    """hobbit::commit_levels_on_host = false (big Merkle levels stay in HBM, only the levels of <= 1024 digests are filled): the prover only
    uses the level sizes, so the proof-size counter of the reference-free AES run is unchanged."""
    import json
    binary = os.path.join(ROOT, "hobbit_b200", "mlp_prove")
    if not os.path.exists(binary):
        pytest.skip("hobbit_b200/mlp_prove not built")
    p = subprocess.run([binary, "19", "aes", "8", "--reps", "1", "--resident-levels"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ps_kb"] == 1135.046875
