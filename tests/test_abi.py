"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from izpi_b200 import cuda
from izpi_b200.build import build as build_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build_lib()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(izpi_[a-z0-9_]+)\s*\(", src))


def test_exports_match_headers():
    declared = _declared("izpi_cuda.h") | _declared("izpi_host.h") | _declared("izpi_proto.h")
    assert declared == set(cuda.EXPORTS)
    L = C.CDLL(cuda.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_sizes_match_header():
    from izpi_b200 import scene as S
    assert S.PRIM_DTYPE.itemsize == 168 and S.NODE_DTYPE.itemsize == 128
    assert C.sizeof(S.MaterialSpec) == 64 and C.sizeof(S.TextureSpec) == 48 and C.sizeof(S.CameraSpec) == 128
    assert C.sizeof(cuda.RenderConfig) == 88 and C.sizeof(cuda.TraceStats) == 32 and C.sizeof(cuda.RenderStats) == 64
    from izpi_b200 import proto
    # sizeof() of the same structs compiled from include/izpi_proto.h / izpi_scene.h with gcc
    assert C.sizeof(proto.ProtoOptions) == 72 and C.sizeof(proto.ProtoImage) == 24 and C.sizeof(proto.ProtoSPD) == 32
    assert C.sizeof(S.SceneSpecC) == 200


def test_headers_compile_as_plain_c(tmp_path):
    """The boundary is a C ABI: every header must compile as C (no C++-isms), and the struct sizes the Python binding assumes
    are the compiler's."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "izpi_cuda.h"\n#include "izpi_host.h"\n#include "izpi_proto.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(izpi_render_stats), sizeof(izpi_proto_options), sizeof(izpi_scene_spec), '
                   'sizeof(izpi_prim_rec), sizeof(izpi_tri_attr), sizeof(izpi_bvh4_node), sizeof(izpi_render_config), sizeof(izpi_prim_spec), '
                   'sizeof(izpi_material_spec));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(x) for x in out] == [64, 72, 200, 80, 176, 128, 88, 168, 64]


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(cuda.IzpiError) as e:
        cuda.Context(0)
    assert e.value.code == cuda.ECUDA
    assert "no CPU fallback" in str(e.value)


def test_null_arguments_rejected():
    L = cuda.lib()
    assert L.izpi_ctx_create(1, None, None) == cuda.EINVAL
    assert L.izpi_scene_upload(None, None) == cuda.EINVAL
    assert L.izpi_trace_closest(None, 1, None, None, 0.0, 1.0, 0, None, None, None) == cuda.EINVAL
    assert b"bad argument" in L.izpi_last_error()
