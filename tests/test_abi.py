"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/matinvent_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "matinvent_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mi_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_header_symbols():
    from matinvent_b200.csrc.build import build
    path = build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n


def test_prototypes_match_header():
    from matinvent_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "matinvent_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for n in _declared():
        if n == "mi_last_error":
            continue
        assert n in _lib.PROTOTYPES, n
        args = re.search(r"\b" + n + r"\s*\(([^;]*?)\)\s*;", hdr, re.S).group(1).strip()
        cnt = 0 if args in ("void", "") else len(args.split(","))
        assert cnt == len(_lib.PROTOTYPES[n]), "%s: header has %d args, ctypes prototype %d" % (n, cnt, len(_lib.PROTOTYPES[n]))
    lib = _lib.load()
    assert lib.mi_version() >= 100


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "matinvent_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(d, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from matinvent_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("MATINVENT_B200_LIB", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except _lib.MatInventLibError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
    finally:
        monkeypatch.setattr(_lib, "_lib", None)
