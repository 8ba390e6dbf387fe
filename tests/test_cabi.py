"""CPU: the C-ABI library loads and exports every entry point include/addk.h declares; the ctypes mirrors of the
header structs have the C layout; the product refuses to run without CUDA (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "addk.h")


def _declared():
    with open(HEADER) as f:
        src = f.read()
    return re.findall(r"^\s*(?:int|long long|const char\*)\s+(addk_\w+)\s*\(", src, flags=re.M)


def test_library_exports_every_declared_symbol():
    from add_gym_b200 import _lib
    names = _declared()
    assert len(names) >= 24
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.addk_version() >= 100


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct, computed by gcc from include/addk.h, against the ctypes mirrors."""
    from add_gym_b200 import _lib
    pairs = {"addk_task": _lib.AddkTask, "addk_motion_lib": _lib.AddkMotionLib, "addk_sim_state": _lib.AddkSimState,
             "addk_env_buffers": _lib.AddkEnvBuffers, "addk_exp_row": _lib.AddkExpRow, "addk_gemm_args": _lib.AddkGemmArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "addk.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = dict(l.split() for l in out if l.strip())
    for cname, cls in pairs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_update_ctx_field_table_parses():
    from add_gym_b200 import _lib
    ptrs, ints, f64s = _lib.ctx_field_names()
    assert "params" in ptrs and "slabs" in ptrs and "mb_rows" in ints and "lr" in f64s
    assert len(set(ptrs + ints + f64s)) == len(ptrs) + len(ints) + len(f64s)
    assert _lib.lib().addk_update_ctx_size() == 8 * (len(ptrs) + len(ints) + len(f64s))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_refuses_to_run_without_cuda():
    from add_gym_b200 import _lib, config
    from add_gym_b200.add_agent import ADDAgent
    with pytest.raises(_lib.AddkError):
        ADDAgent(config.default_config(num_envs=4))
    with pytest.raises(_lib.AddkError):
        _lib.ptr(torch.zeros(3))
