"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/wsi_hgnn.h declares (no compute calls - there is no GPU here), and the product path
refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wsi_hgnn.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wsi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from wsi_hgnn_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wsi_hgnn.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in wsi_hgnn_b200/_lib.py"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header declarations differ"
    assert lib.wsi_abi_version() == _lib.ABI_VERSION


def test_abi_version_macro_matches():
    from wsi_hgnn_b200 import _lib
    m = re.search(r"#define\s+WSI_ABI_VERSION\s+(\d+)", open(HEADER).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_head_perm_is_a_head_preserving_permutation():
    # pure host function of the library: safe without a GPU
    from wsi_hgnn_b200 import ops
    for D, H in [(128, 4), (256, 8), (512, 4), (512, 1), (1024, 32), (384, 2)]:
        p = ops.head_perm(D, H)
        assert sorted(p.tolist()) == list(range(D))
        dk, G = D // H, 32 // H
        for pos, col in enumerate(p.tolist()):
            lane = (pos // 4) % 32
            assert col // dk == lane // G, "a lane must only hold columns of its own head"
    assert ops.head_perm(200, 4) is None and ops.head_perm(64, 4) is None and ops.head_perm(512, 3) is None


def test_error_channel():
    from wsi_hgnn_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_int32 * 64)()
    rc = lib.wsi_head_perm(64, 4, buf)
    assert rc == -3 and b"lane-grouped" in lib.wsi_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(rc, "wsi_head_perm")


def test_no_cpu_fallback():
    """A CPU graph must make the product path fail loudly (never route through the oracle)."""
    import helpers
    fx, G, m = helpers.golden_setup("heat4_rand_T3", helpers.build_ours)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(G)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wsi_hgnn_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle/"


def test_unknown_pooling_raises():
    from wsi_hgnn_b200.models import HEATNet4
    with pytest.raises(NotImplementedError):
        HEATNet4(8, 16, 2, 1, 4, {"0": 0}, 0.1, graph_pooling_type="att")


def test_struct_layouts_match_the_header(tmp_path):
    """the ctypes mirrors of the header's structs (wsi_heat_graph / wsi_heat_params / wsi_slide_desc) have the C layout:
    a C program compiled against include/wsi_hgnn.h prints sizeof / offsetof, compared with ctypes"""
    import shutil
    import subprocess
    from wsi_hgnn_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"wsi_heat_graph": _lib.HeatGraph, "wsi_heat_params": _lib.HeatParams, "wsi_slide_desc": _lib.SlideDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = os.path.join(tmp_path, "layout.c")
    open(src, "w").write("\n".join(lines))
    exe = os.path.join(tmp_path, "layout")
    subprocess.run(["gcc", "-std=c11", "-o", exe, src], check=True)
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
