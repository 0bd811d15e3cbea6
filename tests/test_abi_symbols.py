"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/kmerlr_b200.h
declares, and fails loudly (no fallback) when there is no GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "kmerlr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmerlr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from kmerlr_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from kmerlr_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build the library first: python __graft_entry__.py"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH]).decode()
    exported = set(re.findall(r"\bT (kmerlr_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    L = _lib.lib()
    for s in header_symbols():
        assert hasattr(L, s)
    assert L.kmerlr_version() == 100


def test_library_is_sm100a_only():
    from kmerlr_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "--list-elf", _lib.SO_PATH]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}


def test_pure_host_entry_points():
    """CoeffIndex and the slot count are plain arithmetic (kmerLr_coefficients_index.go:26-54,
    kmerLr_predict_genomic.go:152-156) and need no GPU"""
    import kmerlr_b200 as K
    n = 70
    ci = K.CoeffIndex(n)
    assert ci.Dim() == 2486
    seen = []
    for a in range(n):
        for b in range(a, n):
            j = ci.Ind2Sub(a, b)
            assert ci.Sub2Ind(j - 1) == (a, b)
            seen.append(j)
    assert sorted(seen) == list(range(1, ci.Dim()))
    from kmerlr_b200 import _lib
    L = _lib.lib()
    assert L.kmerlr_window_slots(300, 200, 10) == 11 and L.kmerlr_window_slots(200, 200, 10) == 0


def test_no_gpu_fails_loudly():
    """the product path has no CPU fallback: without a usable GPU every compute call errors"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import kmerlr_b200 as K
    with pytest.raises(K.KmerLrError) as e:
        K.init(0)
    assert e.value.code == 4
    with pytest.raises(K.KmerLrError):
        K.compile_test_data(None, K.NewKmerCounter(1, 4), None, None, True, False, ["ACGT"])
    with pytest.raises(K.KmerLrError):
        K.from_dense([[1.0, 2.0]])


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kmerlr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f
