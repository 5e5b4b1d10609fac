"""The C-ABI library: it loads on a machine without a GPU, exports every symbol that
include/toygpu.h declares, and refuses to do any work without a CUDA device (no CPU path)."""
import ctypes
import os
import re

import pytest

import toycluster_b200 as tc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "toygpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z_A-Z0-9]+)\s*\(", text)))


def test_header_declares_the_operator_entry_points():
    names = _declared()
    for must in ("tg_create", "tg_destroy", "tg_upload", "tg_download", "tg_find_sph_quantities",
                 "tg_regularise", "tg_bfld_from_rotA", "tg_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(tc.LIB_PATH):
        tc.build()
    lib = ctypes.CDLL(tc.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert set(tc.EXPORTS) <= set(_declared())


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors have the sizes gcc gives the structs of include/toygpu.h."""
    import subprocess
    from toycluster_b200 import api
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "toygpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(tg_config), sizeof(tg_stats), sizeof(tg_halo), sizeof(tg_bfield),'
                   'sizeof(tg_exchange));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    got = [ctypes.sizeof(c) for c in (api._Config, api.Stats, api._Halo, api._BField, api._Exchange)]
    assert got == want, (got, want)
    assert want[0] == 88 and want[2] == 72


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(tc.ToyGpuError, match="no CUDA device|CUDA"):
        tc.HotPath(1024, 1000.0, 1.0, 1e5, [[0, 0, 0, 1e-6, 0.54, 100, 1000, 0, 1.0]])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "toycluster_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libtoyoracle" not in src and "libtoyref" not in src, f
