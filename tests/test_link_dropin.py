"""Link-level drop-in: the reference's OWN main() (main.cpp:1171-1237 -> prove_circuit), test_PC and test_Elastic_PC, compiled from the
reference's sources unchanged, with the hot-path entry points (commit_standard / open_standard / commit / open /
prove_multiplication_tree_stream_shallow / prove_gate_consistency[_lookups]) weakened by objcopy and re-defined by
hobbit_b200/host/hobbit_adapter.cpp in the global namespace (oracle/build_pigeon_gpu.sh).  The proof sizes the reference prints must be the
values the unmodified CPU reference prints (SURVEY §9 / BASELINE.md)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _run(binary, *args, timeout=600):
    path = os.path.join(REFDIR, binary)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time: oracle/build_pigeon_gpu.sh)" % binary)
    p = subprocess.run([path] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.mark.parametrize("args,ps", [((9, 18, 18, 1, 4, 1024, 256, 256, 16), 559.0),          # MLP_test.sh
                                     ((5, 19, 8, 1), 1135.046875),                           # test_aes.sh
                                     ((6, 19, 17, 1), 1329.890625)])                         # sql_test.sh
def test_reference_main_on_the_gpu_backend(args, ps):
    out = _run("pigeon_gpu", *args)
    m = re.search(r"Ps : ([0-9.]+) KB", out)
    assert m, out[-1500:]
    assert float(m.group(1)) == ps


def test_reference_test_PC_on_the_gpu_backend():
    out = _run("ref_pc_gpu", "pc", 20, 4, 32)                      # test_PC(2^20, 4, 32): prints "ps,vt" (Our_PC.cpp:826)
    m = re.findall(r"^([0-9.]+),([0-9.]+)$", out, re.M)
    assert m and float(m[-1][0]) == 3839.078125, out[-1500:]


def test_reference_test_Elastic_PC_on_the_gpu_backend():
    out = _run("ref_pc_gpu", "elastic", 22, 18, 1)                 # test_Elastic_PC(2^22, 1) with BUFFER_SPACE 2^18: open prints "PC : ps = .."
    m = re.search(r"PC : ps = ([0-9.]+)", out)
    assert m, out[-1500:]
    # the same run through the reference-free tool (host mirror only) must report the same proof size
    tool = os.path.join(ROOT, "hobbit_b200", "pc_prove")
    p = subprocess.run([tool, "elastic", "22", "18", "1", "--reps", "0"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, WORLD_SIZE="1"))
    assert p.returncode == 0, p.stdout + p.stderr
    import json
    assert float(m.group(1)) == json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])["ps_kb"]
