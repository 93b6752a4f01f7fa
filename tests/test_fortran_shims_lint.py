"""Static checks of the Fortran shim modules (fortran/*.f90, *.F90). No Fortran compiler exists in this image or on the
GPU box (profiles/r2_probe_fortran_gpubox.txt), so the constraints `gfortran -std=f2008` (the reference's flag,
Makefile:51) would enforce on the boundary are checked textually:
 * every bind(C) name is a symbol include/tfx.h declares and libtfx.so exports, with the same number of arguments;
 * F2008 C1276: a pure FUNCTION has no intent(out) / intent(inout) dummy;
 * no F2018-only `error stop` inside pure procedures;
 * every procedure called through tfx_c_api is declared in its interface block.
"""
import glob
import os
import re
import subprocess

import tomofastx_b200 as tfx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "fortran", "*.f90")) + glob.glob(os.path.join(ROOT, "fortran", "*.F90")))


def _joined(path):
    """Source with comments stripped and continuation lines joined."""
    out, cur = [], ""
    for raw in open(path):
        line = raw.rstrip("\n")
        # strip comments (no '!' inside the string literals that matter here except messages: keep it simple)
        if "!" in line:
            q = False
            for i, ch in enumerate(line):
                if ch == '"':
                    q = not q
                if ch == "!" and not q:
                    line = line[:i]
                    break
        line = line.strip()
        if not line:
            continue
        if cur:
            line = line[1:].strip() if line.startswith("&") else line
        cur += line
        if cur.endswith("&"):
            cur = cur[:-1].rstrip() + " "
            continue
        out.append(cur)
        cur = ""
    return out


def _c_prototypes():
    src = open(os.path.join(ROOT, "include", "tfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(tfx_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def _interfaces():
    """(kind, pure, name, nargs, body lines) for every bind(C) interface of tfx_c_api.f90."""
    lines = _joined(os.path.join(ROOT, "fortran", "tfx_c_api.f90"))
    res = []
    i = 0
    while i < len(lines):
        m = re.match(r"(pure\s+)?(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", lines[i], re.I)
        if m:
            body = []
            j = i + 1
            while not re.match(r"end (function|subroutine)", lines[j], re.I):
                body.append(lines[j])
                j += 1
            args = [a for a in m.group(4).split(",") if a.strip()]
            res.append((m.group(2).lower(), bool(m.group(1)), m.group(3), m.group(5), len(args), body))
            i = j
        i += 1
    return res


def test_bindings_match_the_header_and_the_library():
    protos = _c_prototypes()
    out = subprocess.check_output(["nm", "-D", "--defined-only", tfx.build()], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    ifs = _interfaces()
    assert len(ifs) > 50
    for kind, pure, fname, cname, nargs, body in ifs:
        assert fname == cname, (fname, cname)
        assert cname in protos, cname + " is bound in tfx_c_api.f90 but not declared in include/tfx.h"
        assert cname in exported, cname
        assert protos[cname] == nargs, (cname, protos[cname], nargs)


def test_pure_functions_have_only_intent_in_or_value_dummies():
    for kind, pure, fname, cname, nargs, body in _interfaces():
        if pure and kind == "function":
            for line in body:
                assert not re.search(r"intent\(\s*(out|inout)\s*\)", line, re.I), (fname, line)      # F2008 C1276


def test_no_error_stop_and_no_impure_calls_in_pure_procedures():
    pure_c = {f for kind, pure, f, c, n, b in _interfaces() if pure}
    for path in FILES:
        lines = _joined(path)
        inside = None
        for line in lines:
            m = re.match(r"(pure|elemental)\s+(function|subroutine)\s+(\w+)", line, re.I)
            if m and "bind(" not in line.lower():
                inside = m.group(3)
                continue
            if inside and re.match(r"end (function|subroutine)", line, re.I):
                inside = None
                continue
            if inside:
                assert "error stop" not in line.lower(), (path, inside)
                for called in re.findall(r"\b(tfx_\w+)\s*\(", line):
                    assert called in pure_c, "%s: pure procedure %s calls impure %s" % (path, inside, called)


def test_every_called_entry_point_is_declared():
    declared = {f for kind, pure, f, c, n, b in _interfaces()} | {"tfx_check", "tfx_c_api", "tfx_setup_comm"}
    for path in FILES:
        if path.endswith("tfx_c_api.f90"):
            continue
        for line in _joined(path):
            for called in re.findall(r"\b(tfx_\w+)\s*\(", line):
                assert called in declared, (os.path.basename(path), called)
