"""The ISO_C_BINDING module a Fortran host uses (pde-odesolver_b200/fortran/pdeode_b200_mod.f90)
against the C header it binds to (include/pdeode_b200.h).

No Fortran compiler exists in this image, so the module cannot be compiled.  What can be done:
read its interface bodies and its bind(C) type by the interoperability rules of the standard
(Fortran 2003 15.2-15.3: a VALUE dummy is passed by value, any other dummy as a pointer to the
interoperable C type; type(c_ptr) is `void *`, here the opaque context handle; a bind(C) derived
type is the C struct with the same components in the same order), write the C prototypes and the
struct those rules give, and let gcc hold them against the real header with _Static_assert /
__builtin_types_compatible_p.  A wrong VALUE attribute, a swapped argument, an int where the
header has long long, or a struct component out of order fails here.  The call sites of
INTEGRATION.md's route 2 are checked the same way against the interfaces (argument counts, the
structure constructor's component count).

CPU only."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MOD = os.path.join(ROOT, "pde-odesolver_b200", "fortran", "pdeode_b200_mod.f90")
INC = os.path.join(ROOT, "include")

CTYPE = {"integer(c_int)": "int", "integer(c_long_long)": "long long", "real(c_double)": "double",
         "type(c_ptr)": "pdeode_ctx *", "type(pdeode_params)": "pdeode_params"}


def statements(path):
    """Free-form source -> list of statements: comments cut, `&` continuations joined."""
    out, cur = [], ""
    for line in open(path):
        line = line.split("!")[0].rstrip()
        if not line.strip():
            continue
        cur += " " + line.strip()
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        out.append(re.sub(r"\s+", " ", cur.strip()))
        cur = ""
    return out


def parse_module(path):
    st = statements(path)
    consts, fields, funcs = {}, [], {}
    k = 0
    while k < len(st):
        s = st[k]
        m = re.match(r"integer\(c_int\), parameter :: (\w+) = (\d+)$", s)
        if m:
            consts[m.group(1)] = int(m.group(2))
        if re.match(r"type, bind\(C\) :: pdeode_params$", s, re.I):
            k += 1
            while not st[k].lower().startswith("end type"):
                t, names = st[k].split(" :: ")
                for nm in names.split(","):
                    fields.append((CTYPE[t.strip()], nm.split("=")[0].strip()))
                k += 1
        m = re.match(r"(\w+\(\w+\)) function (\w+)\((.*?)\) bind\(C, name=\"(\w+)\"\)$", s)
        if m:
            ret, name, args, cname = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(",")], m.group(4)
            assert name == cname
            decl = {}
            k += 1
            while not st[k].lower().startswith("end function"):
                if not st[k].startswith("import"):
                    spec, names = st[k].split(" :: ")
                    parts = [p.strip() for p in re.split(r",\s*(?![^()]*\))", spec)]
                    for nm in names.split(","):
                        decl[nm.strip()] = parts
                k += 1
            funcs[name] = (CTYPE[ret], [(a, decl[a]) for a in args])
        k += 1
    return consts, fields, funcs


def c_param(parts):
    """Interoperability rules: VALUE -> by value; otherwise a pointer (const for intent(in))."""
    base = CTYPE[parts[0]]
    attrs = [p.lower() for p in parts[1:]]
    if "value" in attrs:
        return base
    const = "const " if "intent(in)" in attrs else ""
    return f"{const}{base} *".replace("* *", "**")


def test_module_matches_header(tmp_path):
    consts, fields, funcs = parse_module(MOD)
    assert set(funcs) >= {"pdeode_create", "pdeode_destroy", "pdeode_set_state", "pdeode_get_state",
                          "pdeode_step", "pdeode_step_many", "pdeode_mass", "pdeode_peak"}
    assert len(fields) == 15
    lines = ['#include <stddef.h>', '#include "pdeode_b200.h"']
    for nm, v in consts.items():
        lines.append(f'_Static_assert({nm} == {v}, "{nm}");')
    lines.append("struct f_params { " + " ".join(f"{t} {n};" for t, n in fields) + " };")
    lines.append('_Static_assert(sizeof(struct f_params) == sizeof(pdeode_params), "size");')
    for t, n in fields:
        lines.append(f'_Static_assert(offsetof(struct f_params, {n}) == offsetof(pdeode_params, {n}), "{n}");')
        lines.append(f'_Static_assert(__builtin_types_compatible_p(__typeof__(((pdeode_params *)0)->{n}), {t}), "{n} type");')
    for name, (ret, args) in funcs.items():
        proto = f"{ret} (*)({', '.join(c_param(p) for _, p in args)})"
        # type(c_ptr) says nothing about the pointee: the handle may be `const pdeode_ctx *` in C
        alt = proto.replace("(pdeode_ctx *", "(const pdeode_ctx *", 1)
        lines.append(f'_Static_assert(__builtin_types_compatible_p(__typeof__(&{name}), {proto}) || '
                     f'__builtin_types_compatible_p(__typeof__(&{name}), {alt}), "{name}: {proto}");')
    src = tmp_path / "binding_check.c"
    src.write_text("\n".join(lines) + "\n")
    r = subprocess.run(["gcc", "-std=gnu11", "-fsyntax-only", f"-I{INC}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + "\n" + src.read_text()


def test_check_catches_a_wrong_binding(tmp_path):
    """The same check on a module with one VALUE attribute dropped must fail (the test tests)."""
    txt = open(MOD).read().replace("integer(c_int), value :: n\n", "integer(c_int) :: n\n")
    assert txt != open(MOD).read()
    bad = tmp_path / "bad.f90"
    bad.write_text(txt)
    _, _, funcs = parse_module(str(bad))
    ret, args = funcs["pdeode_step_many"]
    proto = f"{ret} (*)({', '.join(c_param(p) for _, p in args)})"
    src = tmp_path / "bad.c"
    src.write_text(f'#include "pdeode_b200.h"\n_Static_assert(__builtin_types_compatible_p(__typeof__(&pdeode_step_many), {proto}), "x");\n')
    r = subprocess.run(["gcc", "-std=gnu11", "-fsyntax-only", f"-I{INC}", str(src)], capture_output=True, text=True)
    assert r.returncode != 0


def test_integration_call_sites_match_the_interfaces():
    """INTEGRATION.md route 2: every `rc = pdeode_x(...)` passes as many arguments as the
    interface declares, and the structure constructor names every component."""
    _, fields, funcs = parse_module(MOD)
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = "\n".join(re.findall(r"```fortran\n(.*?)```", txt, re.S))
    block = re.sub(r"&\s*\n\s*", " ", block)
    calls = re.findall(r"rc = (pdeode_\w+)\(([^)]*)\)", block)
    assert len(calls) >= 8
    for name, args in calls:
        assert name in funcs, name
        assert len([a for a in args.split(",") if a.strip()]) == len(funcs[name][1]), (name, args)
    ctor = re.search(r"pp = pdeode_params\(([^)]*)\)", block)
    assert ctor and len(ctor.group(1).split(",")) == len(fields)
    # components in the order of the type: names match case-insensitively, the last is `reserved` = 0
    given = [a.strip().lower() for a in ctor.group(1).split(",")]
    assert given[:-1] == [n.lower() for _, n in fields[:-1]] and given[-1] == "0"


def test_patch_tool_calls_match_the_interfaces():
    """Every call pde-odesolver_b200/fortran/patch_reference.py inserts into solveOrder2 names an
    interface of the module and passes as many arguments as it declares."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "patch_reference", os.path.join(ROOT, "pde-odesolver_b200", "fortran", "patch_reference.py"))
    pr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pr)
    _, fields, funcs = parse_module(MOD)
    text = re.sub(r"&\s*\n\s*", " ", "\n".join([pr.CREATE, pr.MASS, pr.PEAK, pr.FRAME, pr.STEP, pr.FINISH]))
    calls = re.findall(r"rc = (pdeode_\w+)\(([^)]*)\)", text)
    assert {c[0] for c in calls} == {"pdeode_create", "pdeode_set_state", "pdeode_mass", "pdeode_peak",
                                     "pdeode_get_state", "pdeode_step", "pdeode_destroy"}
    for name, args in calls:
        assert len([a for a in args.split(",") if a.strip()]) == len(funcs[name][1]), (name, args)
    ctor = re.search(r"pp = pdeode_params\(([^)]*)\)", text)
    assert [a.strip().lower() for a in ctor.group(1).split(",")][:-1] == [n.lower() for _, n in fields[:-1]]
    for decl in ("type(c_ptr) :: ctx", "type(pdeode_params) :: pp", "integer(c_int) :: rc, cnt",
                 "integer(c_long_long) :: nsum, nmax"):
        assert decl in pr.DECLS
