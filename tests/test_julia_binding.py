"""The reference-side binding (julia/PathMatFacB200.jl) against the C ABI, without a Julia toolchain: every `ccall` of the shim
is parsed -- symbol, return type, argument-type tuple, number of values passed -- and compared with the signature the Python
mirror binds for the same symbol (pathmatfac.jl_b200/_lib.py, itself checked against include/pmf.h and the built library in
tests/test_host_logic.py); the three structs the shim passes by reference are compared field by field with the ctypes
structures (Julia lays out a struct of bits types as C does, so equal field lists mean equal offsets)."""
import ctypes as C
import os
import re

import pathmatfac_b200  # noqa: F401
from pathmatfac_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "PathMatFacB200.jl")

STRUCTS = {"PmfDims": _lib.pmf_dims, "PmfFitOpts": _lib.pmf_fit_opts, "PmfHistory": _lib.pmf_history}
SCALARS = {"Int32": C.c_int32, "Cint": C.c_int, "UInt32": C.c_uint32, "Int64": C.c_int64, "Float32": C.c_float,
           "Float64": C.c_double, "UInt8": C.c_uint8, "Handle": C.c_void_p, "Ptr{Cvoid}": C.c_void_p, "Cstring": C.c_char_p}


def jl_type(t):
    t = t.strip()
    if t in SCALARS:
        return SCALARS[t]
    m = re.fullmatch(r"(?:Ptr|Ref)\{(.+)\}", t)
    assert m, f"unmapped Julia type {t!r}"
    inner = m.group(1).strip()
    if inner in STRUCTS:
        return C.POINTER(STRUCTS[inner])
    return C.POINTER(jl_type(inner))


def same_ctype(a, b):
    """ctypes equality up to the aliases of one platform (c_int is c_int32, POINTER types are cached per target)."""
    if a is b:
        return True
    pa, pb = getattr(a, "_type_", None), getattr(b, "_type_", None)
    if isinstance(pa, str) or isinstance(pb, str) or pa is None or pb is None:        # simple types: compare code and size
        return pa == pb and C.sizeof(a) == C.sizeof(b) and isinstance(pa, str)
    return same_ctype(pa, pb)


def top_level_split(s):
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
    return [x.strip() for x in out]


def ccalls(text):
    """(symbol, return type, [argument types], number of values passed, line) of every ccall in the shim."""
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIBPMF\),\s*([\w{}]+),\s*\(", text):
        i = m.end()
        j = text.index(")", i)                                   # Julia type names hold braces, never parentheses
        types = [t for t in top_level_split(text[i:j]) if t]
        depth, k = 1, j + 1                                      # the values: up to the parenthesis that closes the ccall
        while depth:
            depth += {"(": 1, ")": -1}.get(text[k], 0)
            k += 1
        values = top_level_split(text[j + 1:k - 1].lstrip().lstrip(","))
        out.append((m.group(1), m.group(2), types, len(values), text.count("\n", 0, m.start()) + 1))
    return out


def test_every_ccall_of_the_shim_matches_the_abi():
    text = open(SHIM).read()
    calls = ccalls(text)
    assert len(calls) >= 30 and len(calls) == text.count("ccall(")
    for name, ret, types, n_values, line in calls:
        assert name in _lib.SIGNATURES, f"line {line}: {name} is not a symbol of the ABI"
        res, args = _lib.SIGNATURES[name]
        assert same_ctype(jl_type(ret), res), f"line {line}: {name} returns {res}, the shim says {ret}"
        assert len(types) == len(args) == n_values, f"line {line}: {name} takes {len(args)} arguments, the shim declares {len(types)} and passes {n_values}"
        for pos, (t, a) in enumerate(zip(types, args)):
            assert same_ctype(jl_type(t), a), f"line {line}: {name} argument {pos}: ABI {a}, shim {t}"
    used = {c[0] for c in calls}
    # what a fit through the shim needs end to end: lifecycle, data, parameters both ways, regularisers, noise, the fit itself,
    # the statistics passes of the staging functions, the communicator of the sample-sharded path
    for needed in ("pmf_create", "pmf_destroy", "pmf_set_data", "pmf_set_batch_layout", "pmf_set_factors", "pmf_get_factors",
                   "pmf_set_col_params", "pmf_get_col_params", "pmf_set_batch_values", "pmf_get_batch_values", "pmf_set_noise",
                   "pmf_get_thresholds", "pmf_clear_reg", "pmf_set_reg_l2", "pmf_set_reg_group", "pmf_set_reg_sel_l1",
                   "pmf_set_reg_ard", "pmf_set_reg_fsard", "pmf_set_reg_network", "pmf_set_layer_reg_col", "pmf_set_layer_reg_batch",
                   "pmf_set_frozen", "pmf_reset_opt_state", "pmf_fit", "pmf_column_stats", "pmf_link_col_sqerr", "pmf_batch_stats",
                   "pmf_comm_unique_id", "pmf_comm_init_rank", "pmf_comm_destroy", "pmf_last_error"):
        assert needed in used, needed


def test_structs_of_the_shim_have_the_layout_of_the_abi():
    text = open(SHIM).read()
    for jl_name, ct in STRUCTS.items():
        m = re.search(r"struct " + jl_name + r"\b[^\n]*\n(.*?)\nend", text, re.S)
        assert m, jl_name
        body = re.sub(r"#[^\n]*", "", m.group(1))
        fields = [f.strip() for f in re.split(r"[;\n]", body) if f.strip()]
        got = [tuple(f.split("::")) for f in fields]
        want = ct._fields_
        assert [g[0] for g in got] == [w[0] for w in want], (jl_name, got)
        for (fname, jt), (_, wt) in zip(got, want):
            assert same_ctype(jl_type(jt), wt), (jl_name, fname, jt, wt)
        # and the C layout both sides rely on: natural alignment, no packing
        off = 0
        for fname, wt in want:
            al = C.alignment(wt)
            off = (off + al - 1) // al * al
            assert getattr(ct, fname).offset == off, (jl_name, fname)
            off += C.sizeof(wt)


def test_shim_overrides_the_method_the_reference_calls():
    """src/fit.jl:58 calls `mf_fit!(model; opt=..., ...)` with ONE positional argument (missing 3): the shim must define
    that method on the reference's own function, not a function of its own module with a handle argument."""
    text = open(SHIM).read()
    assert re.search(r"function PM\.mf_fit!\(model::PM\.PathMatFacModel;", text), "PM.mf_fit!(model; ...) is not overridden"
    assert not re.search(r"function (?:PM\.)?mf_fit!\(model[^;)]*,\s*h\b", text)


# ---- the other half of the chain: include/pmf.h against the signatures the mirror binds ------------------------------------------------
C_SCALARS = {"int": C.c_int, "int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint16_t": C.c_uint16,
             "uint8_t": C.c_uint8, "float": C.c_float, "double": C.c_double, "pmf_handle": C.c_void_p,
             "pmf_dims": _lib.pmf_dims, "pmf_losses": _lib.pmf_losses, "pmf_fit_opts": _lib.pmf_fit_opts, "pmf_history": _lib.pmf_history}


def c_type(decl, is_return=False):
    """ctypes type of one C parameter declaration (`const float* A_host`) or return type (`const char*`)."""
    d = re.sub(r"\bconst\b", " ", decl).strip()
    d, n_arr = re.subn(r"\[\d*\]", "", d)                         # an array parameter is a pointer to its element type
    stars = d.count("*") + n_arr
    words = d.replace("*", " ").split()
    base = words[0]
    assert len(words) == (1 if is_return else 2) or (not is_return and len(words) == 1), decl
    if base == "void":
        assert stars >= 1, decl
        t, stars = C.c_void_p, stars - 1
    elif base == "char":
        assert stars == 1, decl
        return C.c_char_p
    else:
        t = C_SCALARS[base]
    for _ in range(stars):
        t = C.POINTER(t)
    return t


def header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "pmf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    out = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?\w+\s*\*?)\s*(pmf_\w+)\s*\(([^)]*)\)\s*;", hdr):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def test_header_prototypes_match_the_signatures_the_mirror_binds():
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    for name, (ret, plist) in protos.items():
        res, args = _lib.SIGNATURES[name]
        if ret == "void":
            assert res is None, name
        else:
            assert same_ctype(c_type(ret, is_return=True), res), (name, ret, res)
        assert len(plist) == len(args), (name, plist, args)
        for pos, (p, a) in enumerate(zip(plist, args)):
            assert same_ctype(c_type(p), a), f"{name} argument {pos}: header `{p}`, mirror {a}"


def test_header_structs_have_the_layout_the_mirror_and_the_shim_use(tmp_path):
    """The C compiler's own offsets of every field of pmf.h's structs (a program that includes the header) against ctypes."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"pmf_dims": _lib.pmf_dims, "pmf_losses": _lib.pmf_losses, "pmf_fit_opts": _lib.pmf_fit_opts, "pmf_history": _lib.pmf_history}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pmf.h"', "int main(void) {"]
    for sname, ct in structs.items():
        lines.append(f'  printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    r = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]                       # also: every field the mirror names exists in the header
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for sname, ct in structs.items():
        assert int(got[sname]) == C.sizeof(ct), sname
        for fname, _ in ct._fields_:
            assert int(got[f"{sname}.{fname}"]) == getattr(ct, fname).offset, (sname, fname)
    # and no field of the header is missing from the mirror: the sizes agree, so a dropped trailing field would show above


def test_struct_constructor_calls_of_the_shim_pass_one_value_per_field():
    """`PmfFitOpts(...)`, `PmfHistory(...)` and `PmfDims(...)` are built positionally: a missing or extra value would shift every
    later field (Julia would raise a MethodError only at the first fit)."""
    text = open(SHIM).read()
    seen = set()
    for m in re.finditer(r"(?<![\w{])(PmfFitOpts|PmfHistory|PmfDims)\(", text):
        depth, k = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[k], 0)
            k += 1
        values = top_level_split(text[m.end():k - 1])
        n = 0
        for v in values:
            if v.endswith("..."):                                    # pointer.(losses)...: one pointer per loss component
                assert v == "pointer.(losses)..." and re.search(r"losses = \[zeros\(Float64, cap\) for _ in 1:5\]", text), v
                n += 5
            else:
                n += 1
        assert n == len(STRUCTS[m.group(1)]._fields_), (m.group(1), values)
        seen.add(m.group(1))
    assert seen == set(STRUCTS)
