"""The C-ABI library must exist in-tree, load, and export every symbol declared in include/b200sim.h (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200sim.h")).read()
    return sorted(set(re.findall(r"\b(b200sim_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gymnasium_robotics_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200sim.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_create_without_cuda_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gymnasium_robotics_b200 import _lib
    from gymnasium_robotics_b200.fetch import FETCH_TASKS, make_task_struct
    from gymnasium_robotics_b200.models import load_model

    L = _lib.lib()
    m = load_model("fetch_reach")
    task = make_task_struct(m, FETCH_TASKS["FetchReach"], "sparse")
    blob = m.to_blob()
    h = ctypes.c_void_p()
    rc = L.b200sim_create(blob, len(blob), None, None, ctypes.byref(task), 4, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in L.b200sim_last_error(None)


def test_vector_env_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gymnasium_robotics_b200.fetch import FetchVectorEnv

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FetchVectorEnv("FetchReach", num_envs=2)


def test_product_package_does_not_import_the_oracle():
    """The product path must never route through the CPU oracle (or the test-only host emulation)."""
    pkg = os.path.join(ROOT, "gymnasium_robotics_b200")
    bad = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|liboracle|oracle_sim|libhostsim|#include\s+\"[^\"]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f
