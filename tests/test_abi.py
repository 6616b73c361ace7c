"""The C-ABI library loads without a GPU and exports every function include/distillclip_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "distillclip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from distillclip_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == names, "ctypes signature table and header disagree"
    assert lib.dcb_compiled_arch() == 100
    assert lib.dcb_version() >= 100


def test_argument_errors_are_reported_without_gpu():
    from distillclip_b200 import _lib
    lib = _lib.load()
    assert lib.dcb_finalize(0, None, None, None, None, None, None) != 0
    assert b"n_terms" in lib.dcb_last_error()
    assert lib.dcb_clip_workspace_bytes(4096, 4096) > 0
    assert lib.dcb_clip_grad_splits(4096, 4096, 512) >= 1


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    import pytest
    import torch
    from distillclip_b200.model import HiddenMSE, HardLabel
    with pytest.raises(_err()):
        HiddenMSE()([torch.zeros(2, 3, 4)], [torch.zeros(2, 3, 4)])
    with pytest.raises(_err()):
        HardLabel()(torch.zeros(4, 4))


def _err():
    from distillclip_b200._lib import DistillClipB200Error
    return DistillClipB200Error


def test_build_stamp_is_location_independent(tmp_path):
    """The GPU box unpacks the tree under a different path: the prebuilt .so must be recognised there (a path-dependent
    stamp made every rank of a torchrun launch rebuild and overwrite the library at the same time)."""
    import shutil
    from distillclip_b200 import build as b
    srcs = b.sources()[:3]
    copies = []
    for s in srcs:
        dst = tmp_path / os.path.basename(s)
        shutil.copy(s, dst)
        copies.append(str(dst))
    assert b._digest(srcs) == b._digest(copies)
    stamp = os.path.join(b.OBJ, "stamp.txt")
    if os.path.exists(b.LIB) and os.path.exists(stamp):          # a built tree: build() must be a no-op
        before = os.path.getmtime(b.LIB)
        b.build()
        assert os.path.getmtime(b.LIB) == before
