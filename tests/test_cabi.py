"""The C-ABI shared library: loads without a GPU, exports exactly what include/plaid_b200.h declares,
and the ctypes prototypes agree with the header's parameter counts.  No compute calls here."""
import ctypes
import re

from reranking_multimodal_retrievers_b200 import _lib, build


def _header_decls():
    text = re.sub(r"/\*.*?\*/", "", open(build.HEADER).read(), flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(plaid_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_builds_and_loads():
    path = build.build_library()
    handle = ctypes.CDLL(path)
    assert handle.plaid_abi_version() == 2
    handle.plaid_arch.restype = ctypes.c_char_p
    assert handle.plaid_arch() == b"sm_100a"


def test_every_declared_symbol_is_exported():
    handle = _lib.lib()
    decls = _header_decls()
    assert len(decls) >= 20
    for name in decls:
        assert hasattr(handle, name), f"{name} declared in plaid_b200.h but not exported"


def test_ctypes_prototypes_match_header():
    decls = _header_decls()
    assert set(decls) == set(_lib._SIGNATURES)
    for name, nargs in decls.items():
        assert len(_lib._SIGNATURES[name]) == nargs, f"{name}: header has {nargs} parameters"


def test_errors_are_reported_not_fatal():
    """A bad argument returns PLAID_ERR_ARG and sets plaid_last_error (no abort / exit like
    filter_pids.cpp:47,98-101); argument validation happens before any CUDA call."""
    handle = _lib.lib()
    rc = handle.plaid_prepare_queries(None, 1, 32, 0, 4, 32, None, None, None, None)
    assert rc == -1
    assert b"null pointer" in handle.plaid_last_error()
    try:
        _lib.call("plaid_select_top", None, None, None, 1, 1, 1, None, None, None, 1, None, None)
    except _lib.PlaidError as e:
        assert "plaid_select_top" in str(e)
    else:
        raise AssertionError("expected PlaidError")
