"""CPU tests: the C-ABI library builds, loads and exports every declared symbol; no compute without a GPU."""
import ctypes as C
import os
import re

import libagmv_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(agmvb_[a-z0-9_]+|AGMV_[A-Za-z0-9]+|CreateAGMV|DestroyAGMV)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(libagmv_b200.LIB_PATH)
    names = [n for n in _declared("agmv_b200.h") if n.startswith("agmvb_")]
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(libagmv_b200.SYMBOLS), "python binding and header disagree"


def test_dropin_exports_reference_api():
    dl = C.CDLL(libagmv_b200.DROPIN_PATH)
    for n in ["CreateAGMV", "DestroyAGMV", "AGMV_EncodeAGMV", "AGMV_DecodeAGMV", "AGMV_EncodeFrame", "AGMV_DecodeFrameChunk",
              "AGMV_DecodeHeader", "AGMV_EncodeHeader", "AGMV_EncodeVideo", "AGMV_EncodeFullAGMV"]:
        assert hasattr(dl, n), n


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device context creation must fail loudly (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        libagmv_b200.Context(0)
    except libagmv_b200.AgmvError as e:
        assert e.code == 5
    else:
        raise AssertionError("context creation succeeded without a GPU")
