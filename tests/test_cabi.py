"""CPU checks of the C-ABI boundary: the library builds for sm_100a, loads, exports every symbol that
include/srggnn.h declares, the ctypes prototypes cover the header, and compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "srggnn.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from situation_recognition_b200 import _lib
    return _lib.load()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsrggnn.so does not export %s" % n


def test_ctypes_prototypes_cover_header():
    from situation_recognition_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_param_struct_layout():
    from situation_recognition_b200 import _lib
    assert ctypes.sizeof(_lib.SrgParams) == 18 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.SrgGrads) == 20 * ctypes.sizeof(ctypes.c_void_p)


def test_sass_is_blackwell_native():
    """The shipped kernels must contain tcgen05 MMA / TMEM loads / TMA, not the legacy mma.sync path."""
    import shutil
    import subprocess
    so = os.path.join(ROOT, "situation_recognition_b200", "libsrggnn.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass          # tcgen05.mma kind::f16
    assert "LDTM" in sass             # tcgen05.ld
    assert "UTMALDG" in sass and "UTMASTG" in sass   # TMA load / store
    assert "HMMA.16816" not in sass   # no mma.sync fallback


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from situation_recognition_b200 import _lib
    import situation_recognition_b200 as S
    from situation_recognition_b200.synthetic import make_train_json
    h = ctypes.c_void_p()
    rc = lib.srg_create(ctypes.byref(h), 0, 2048, 6, 4, 504, 190, 2001)
    assert rc != 0 and lib.srg_last_error()          # no device -> error code + message, never a silent fallback
    enc = S.imsitu_encoder(make_train_json(seed=0, images_per_verb=1), verbose=False)
    m = S.FCGGNN(enc, 256, backbone=None)
    with pytest.raises(_lib.SrgError):
        m(torch.zeros(2, 256), torch.zeros(2, dtype=torch.long))
    with pytest.raises(_lib.SrgError):
        m.ggsnn(torch.zeros(2, 256), verb=True)


def test_state_dict_keys_match_reference_names():
    """Checkpoint compatibility: the non-backbone keys of the reference's model_state_dict."""
    import situation_recognition_b200 as S
    from situation_recognition_b200.synthetic import make_train_json
    enc = S.imsitu_encoder(make_train_json(seed=0, images_per_verb=1), verbose=False)
    m = S.FCGGNN(enc, 256, backbone=None)
    keys = set(m.state_dict())
    expect = {"role_emb.weight", "verb_emb.weight", "verb_classifier.1.weight", "verb_classifier.1.bias",
              "nouns_classifier.1.weight", "nouns_classifier.1.bias"}
    for n in ["W_p", "W_z", "U_z", "W_r", "U_r", "W_h", "U_h"]:
        expect |= {"ggsnn.%s.weight" % n, "ggsnn.%s.bias" % n}
    assert keys == expect
    assert m.role_emb.padding_idx == enc.get_num_roles()
    assert m.module is m
