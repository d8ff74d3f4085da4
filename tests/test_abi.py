"""The C-ABI library loads and exports every symbol include/shapes_b200.h declares;
without a GPU it fails loudly instead of falling back.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "shapes_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shapes_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported(product_lib):
    from shapes_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    for n in names:
        assert getattr(product_lib, n) is not None


def test_struct_layout_matches_header(product_lib):
    from shapes_b200 import _lib
    # pointers and int64 only before the stats block: 2 + 1 + 5 + 5 + 7 + 6 + 7 + 2 + 6 = 41 words, then stats
    # pairs 3, n_contacts 1, keys 4, flip 1, contact 5, j_np 6, b_np 1, radii/normal 6, j_f 6, b_f 1,
    # inverse effective masses 2, warm start 3, aabb 4, world 2 words; then the stats block
    assert C.sizeof(_lib.FrameOut) == 8 * (3 + 1 + 4 + 1 + 5 + 6 + 1 + 6 + 6 + 1 + 2 + 3 + 4 + 2) + 8 + 4 + 4 + 8 + 4 + 4
    assert _lib.FrameOut.n_contacts.offset == 24
    assert product_lib.shapes_version().startswith(b"shapes_b200")


def test_every_struct_field_offset_matches_the_header(tmp_path):
    """gcc compiles a C program against include/shapes_b200.h that prints sizeof / offsetof of every field of the five
    public structs; the ctypes mirror must agree field by field (a reordered or resized field would otherwise read
    garbage silently)."""
    import subprocess
    from shapes_b200 import _lib
    structs = {"shapes_frame_out": _lib.FrameOut, "shapes_device_view": _lib.DeviceView, "shapes_step_config": _lib.StepConfig,
               "shapes_step_stats": _lib.StepStats, "shapes_frame_info": _lib.FrameInfo}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "shapes_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line:
            continue
        cname, field, value = line.split()
        cls = structs[cname]
        if field == "size":
            assert C.sizeof(cls) == int(value), (cname, C.sizeof(cls), value)
        else:
            assert getattr(cls, field).offset == int(value), (cname, field, getattr(cls, field).offset, value)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_no_silent_fallback_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU boxes")
    ctx = C.c_void_p()
    rc = product_lib.shapes_create(C.byref(ctx), 0, 16, 64, 64, 128)
    assert rc == -2 and not ctx.value          # SHAPES_E_CUDA
    assert len(product_lib.shapes_last_error(None)) > 0


def test_bad_arguments_rejected(product_lib):
    ctx = C.c_void_p()
    assert product_lib.shapes_create(C.byref(ctx), 0, -1, 1, 1, 1) == -1
    assert product_lib.shapes_create_ranked(C.byref(ctx), 0, 2, 2, None, 1, 1, 1, 1) == -1
    assert product_lib.shapes_set_hulls(None, 0, None, None, None, None, None, None) == -1


def test_product_does_not_touch_the_oracle():
    """Nothing under shapes_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for base in ("shapes_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                    with open(os.path.join(d, fn)) as f:
                        if re.search(r"oracle", f.read(), flags=re.I):
                            bad.append(os.path.join(d, fn))
    assert not bad, bad
