"""Static checks of the cgo binding (go/bn254), which cannot be compiled here (no Go toolchain): every C function the
Go package calls is declared in include/bn254_b200.h with the same number of arguments, every C type and constant it
names exists, and the header itself is plain C99 (cgo compiles the preamble as C, not C++)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bn254_b200.h")
GO_DIR = os.path.join(ROOT, "go")


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _split_args(s):
    """Top-level comma split of an argument list (no outer parentheses)."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def header_prototypes():
    src = _strip_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"\b(bn254_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        protos[name] = 0 if args in ("", "void") else len(_split_args(args))
    return protos, src


def go_calls():
    calls = []
    for dirpath, _, files in os.walk(GO_DIR):
        for fn in files:
            if not fn.endswith(".go"):
                continue
            src = _strip_comments(open(os.path.join(dirpath, fn)).read())
            for m in re.finditer(r"C\.(bn254_[a-z0-9_]+)\s*\(", src):
                i, depth = m.end(), 1
                while depth:
                    depth += {"(": 1, ")": -1}.get(src[i], 0)
                    i += 1
                args = src[m.end():i - 1].strip()
                calls.append((fn, m.group(1), len(_split_args(args)) if args else 0))
    return calls


def test_header_is_plain_c99():
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    prog = '#include "bn254_b200.h"\nint main(void) { return 0; }\n'
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.dirname(HEADER), "-fsyntax-only", "-x", "c", "-"],
                       input=prog, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_cgo_call_matches_a_header_prototype():
    protos, _ = header_prototypes()
    calls = go_calls()
    assert len(calls) >= 40, "the Go package calls the C ABI from every batch entry point"
    bad = [(fn, name, n, protos.get(name)) for fn, name, n in calls if protos.get(name) != n]
    assert not bad, "cgo call sites that disagree with include/bn254_b200.h (file, function, args in Go, args in C): %r" % bad


def test_cgo_types_and_constants_exist_in_the_header():
    _, hdr = header_prototypes()
    names = set()
    for dirpath, _, files in os.walk(GO_DIR):
        for fn in files:
            if fn.endswith(".go"):
                names |= set(re.findall(r"C\.((?:bn254|BN254)_[A-Za-z0-9_]+)\b", _strip_comments(open(os.path.join(dirpath, fn)).read())))
    missing = sorted(n for n in names if not re.search(r"\b%s\b" % re.escape(n), hdr))
    assert not missing, "named in go/ but absent from the header: %r" % missing


def test_python_mirror_device_signatures_match_the_header():
    """Engine.dev() builds its ctypes argument list from a signature string per entry point (bn254.py _DEV_SIGS); a wrong
    count there is a silent stack mismatch, so every string is checked against the header's prototype: context + one
    argument per letter (two for a byte string: pointer and length) + the stream."""
    import sys

    sys.path.insert(0, ROOT)
    from gopairingbasedcryptography_b200 import bn254

    protos, _ = header_prototypes()
    sigs = bn254.Engine._DEV_SIGS
    assert len(sigs) >= 30
    bad = []
    for name, sig in sigs.items():
        want = protos.get("bn254_" + name)
        got = 1 + sum(2 if c == "b" else 1 for c in sig) + 1
        if want != got:
            bad.append((name, sig, got, want))
    assert not bad, "(entry point, signature, arguments built, arguments declared): %r" % bad
