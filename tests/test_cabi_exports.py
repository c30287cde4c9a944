"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol that
include/hobbit_b200.h declares, and refuses (loudly) to create a context when no CUDA device exists."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hobbit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import hobbit_b200
    lib = hobbit_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "include/hobbit_b200.h declares %s but libhobbit_b200.so does not export it" % n


def test_host_mirror_library_links():
    so = os.path.join(ROOT, "hobbit_b200", "libhobbit_host.so")
    if not os.path.exists(so):
        pytest.skip("host library not built")
    ctypes.CDLL(os.path.join(ROOT, "hobbit_b200", "libhobbit_b200.so"), mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(so)
    # C++ symbols with the reference's names in namespace hobbit
    for mangled in ("_ZN6hobbit15commit_standardERSt6vectorINS_1FESaIS1_EERNS_5_hashERS0_IS0_IS5_SaIS5_EESaIS8_EERS0_IS0_IS3_SaIS3_EESaISD_EEi",
                    "_ZN6hobbit19expander_init_storeExi", "_ZN6hobbit15init_commitmentEb"):
        assert hasattr(L, mangled), mangled


def test_no_cpu_fallback():
    import torch
    import hobbit_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hobbit_b200.HobbitError):
        hobbit_b200.Context(0)


def test_product_never_imports_oracle():
    """The product (package + csrc + host + include) must not reference anything under oracle/."""
    bad = []
    for base in ("hobbit_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    for line in txt.splitlines():
                        s = line.strip()
                        if ("oracle/" in s or "hobbit_oracle" in s or "libhobbit_ref" in s) and not s.startswith(("//", "#", "*", "/*")):
                            bad.append((f, s[:100]))
    assert not bad, bad
