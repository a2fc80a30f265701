"""The Julia boundary (ext/SeismicWaves_B200BackendExt.jl) checked without a Julia runtime (none exists in this image):

  * every C struct the extension mirrors: field order, offsets and sizes computed from the Julia definitions with the C layout rules
    Julia applies to isbits structs, against what the library itself reports (swb_abi_layout, built with offsetof / sizeof);
  * the same for the ctypes mirror in seismicwaves.jl_b200/_lib.py and for the declarations in include/swb200.h;
  * every ccall of the extension against the prototype in include/swb200.h: symbol, return type, argument count and argument kinds;
  * the members of the four backend tuples against the list the reference's L3/L4 code looks up on a backend module.

No GPU needed: swb_abi_layout never touches a device."""
import ctypes as C
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = os.path.join(ROOT, "ext", "SeismicWaves_B200BackendExt.jl")
HEADER = os.path.join(ROOT, "include", "swb200.h")


@pytest.fixture(scope="module")
def layout():
    import swb200 as S

    return json.loads(S._lib.load().swb_abi_layout().decode())


@pytest.fixture(scope="module")
def jl_src():
    return open(EXT, encoding="utf-8").read()


@pytest.fixture(scope="module")
def header():
    txt = open(HEADER).read()
    return re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)


# ---- C header ----------------------------------------------------------------------------------------------------------
def header_structs(h):
    """{struct name: [member names in declaration order]}"""
    out = {}
    for body, name in re.findall(r"typedef struct\s*\{(.*?)\}\s*(\w+)\s*;", h, flags=re.S):
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(struct\s+)?\w+\s+", "", decl, count=1)  # drop the type
            for part in decl.split(","):
                m = re.search(r"(\w+)\s*(\[\d+\])?\s*$", part.strip().lstrip("*").replace("*const", "").replace("const ", ""))
                names.append(m.group(1))
        out[name] = names
    return out


def header_prototypes(h):
    """{function: (return type, [parameter type strings])}"""
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|int64_t|const char \*|void)\s*(swb_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret.replace(" ", ""), params)
    return protos


def c_kind(param: str) -> str:
    p = param.split("/*")[0].strip()
    if "*" in p or "[" in p:
        return "ptr"
    for t, k in (("int32_t", "i32"), ("int64_t", "i64"), ("size_t", "size"), ("double", "f64")):
        if re.match(rf"(const\s+)?{t}\b", p):
            return k
    raise AssertionError(f"unclassified C parameter: {param!r}")


def test_layout_report_covers_the_header(layout, header):
    hs = header_structs(header)
    assert set(layout) == set(hs), (sorted(layout), sorted(hs))
    for name, members in hs.items():
        assert [f[0] for f in layout[name]["fields"]] == members, name
        end = 0
        for _, off, size in layout[name]["fields"]:
            assert off >= end, (name, "fields overlap")
            end = off + size
        assert layout[name]["size"] >= end


# ---- ctypes mirror -----------------------------------------------------------------------------------------------------
def test_ctypes_structures_match_the_library_layout(layout):
    import swb200 as S

    for name, spec in layout.items():
        cls = getattr(S._lib, name)
        assert C.sizeof(cls) == spec["size"], name
        fields = [(f[0].rstrip("_"), getattr(cls, f[0]).offset, getattr(cls, f[0]).size) for f in cls._fields_]  # `lambda_`: Python keyword
        assert fields == [tuple(f) for f in spec["fields"]], name


# ---- Julia structs -----------------------------------------------------------------------------------------------------
PRIM = {"Int32": (4, 4), "Int64": (8, 8), "UInt64": (8, 8), "Float64": (8, 8), "Float32": (4, 4), "Cdouble": (8, 8), "Csize_t": (8, 8), "UInt8": (1, 1)}


def jl_structs(src):
    """{Julia struct name: (C struct name from the trailing comment, [(field, type string)])} for the immutable C-mirror structs"""
    out = {}
    for m in re.finditer(r"^struct (\w+)[ \t]*#[ \t]*(swb_\w+)[^\n]*\n(.*?)^end", src, flags=re.S | re.M):
        fields = []
        for line in m.group(3).splitlines():
            line = line.split("#")[0]
            for decl in line.split(";"):
                decl = decl.strip()
                if decl:
                    fname, ftype = decl.split("::")
                    fields.append((fname.strip(), ftype.strip()))
        out[m.group(1)] = (m.group(2), fields)
    return out


def jl_size_align(t, structs, cache):
    t = t.strip()
    if t.startswith("Ptr{"):
        return 8, 8
    if t in PRIM:
        return PRIM[t]
    m = re.match(r"NTuple\{(\d+),\s*(.+)\}$", t)
    if m:
        s, a = jl_size_align(m.group(2), structs, cache)
        return int(m.group(1)) * s, a
    if t in structs:
        if t not in cache:
            cache[t] = jl_layout(structs[t][1], structs, cache)
        return cache[t][0], cache[t][1]
    raise AssertionError(f"unknown Julia field type {t!r}")


def jl_layout(fields, structs, cache):
    """C layout rules (what Julia uses for an immutable struct of isbits fields): natural alignment, size rounded up to the largest alignment"""
    off, maxa, out = 0, 1, []
    for fname, ftype in fields:
        s, a = jl_size_align(ftype, structs, cache)
        off = (off + a - 1) // a * a
        out.append((fname, off, s))
        off += s
        maxa = max(maxa, a)
    return (off + maxa - 1) // maxa * maxa, maxa, out


def test_julia_structs_match_the_library_layout(layout, jl_src):
    structs = jl_structs(jl_src)
    mirrored = {c for c, _ in structs.values()}
    assert mirrored == set(layout), f"structs without a Julia mirror: {sorted(set(layout) - mirrored)}; unknown mirrors: {sorted(mirrored - set(layout))}"
    cache = {}
    for jname, (cname, fields) in structs.items():
        size, _, offs = jl_layout(fields, structs, cache)
        assert size == layout[cname]["size"], (jname, cname, size, layout[cname]["size"])
        assert [tuple(f) for f in layout[cname]["fields"]] == offs, (jname, cname)


# ---- ccalls ------------------------------------------------------------------------------------------------------------
def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def jl_ccalls(src):
    """[(symbol, return type, [argument types], number of values passed)]"""
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), lib\),\s*", src):
        i, depth, start = m.end(), 1, m.end()
        while depth:  # find the parenthesis closing `ccall(`
            ch = src[i]
            depth += ch in "({["
            depth -= ch in ")}]"
            i += 1
        parts = split_top(src[start:i - 1])
        ret, argt = parts[0], parts[1]
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        types = split_top(argt[1:-1])
        calls.append((m.group(1), ret, types, len(parts) - 2))
    return calls


def jl_kind(t: str) -> str:
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Int64": "i64", "Csize_t": "size", "Cdouble": "f64"}[t]


def test_every_ccall_matches_its_prototype(jl_src, header):
    protos = header_prototypes(header)
    calls = jl_ccalls(jl_src)
    assert len(calls) >= 40
    rets = {"int32_t": "Int32", "int64_t": "Int64", "constchar*": "Cstring"}
    for sym, ret, types, nvals in calls:
        assert sym in protos, f"ccall of {sym}: not declared in include/swb200.h"
        cret, cparams = protos[sym]
        assert rets[cret] == ret, (sym, ret, cret)
        assert len(types) == len(cparams) == nvals, (sym, types, cparams, nvals)
        assert [jl_kind(t) for t in types] == [c_kind(p) for p in cparams], (sym, types, cparams)
    # struct-typed references point at the right mirror
    structs = jl_structs(jl_src)
    for sym, _, types, _ in calls:
        for t, p in zip(types, protos[sym][1]):
            m = re.match(r"(?:Ref|Ptr)\{(\w+)\}", t)
            if m and m.group(1) in structs:
                assert structs[m.group(1)][0] in p, (sym, t, p)


def test_exports_cover_every_declared_symbol(header):
    import swb200 as S

    protos = header_prototypes(header)
    assert set(protos) == set(S._lib.EXPORTS), (sorted(set(protos) - set(S._lib.EXPORTS)), sorted(set(S._lib.EXPORTS) - set(protos)))
    lib = S._lib.load()
    for name in protos:
        assert hasattr(lib, name), name


# ---- backend-module members -----------------------------------------------------------------------------------
REQUIRED = {
    "Acoustic2D_CD_CPML_B200": {"Data", "zeros", "ones", "forward_onestep_CPML!", "adjoint_onestep_CPML!", "prescale_residuals!", "correlate_gradient!"},
    "Acoustic2D_VD_CPML_B200": {"Data", "zeros", "ones", "forward_onestep_CPML!", "adjoint_onestep_CPML!", "prescale_residuals!", "correlate_gradient_m0!",
                                "correlate_gradient_m1!"},
    "Elastic2D_Iso_CPML_B200": {"Data", "zeros", "ones", "forward_onestep_CPML!", "adjoint_onestep_CPML!", "correlate_gradients!"},
}


def test_backend_tuples_expose_the_members_the_reference_calls(jl_src):
    for name, members in REQUIRED.items():
        m = re.search(rf"const {name} = \(;(.*?)\)\n", jl_src, flags=re.S)
        assert m, name
        have = {p.split("=")[0].strip() for p in split_top(m.group(1))}
        assert members <= have, (name, members - have)
    assert "const Acoustic3D_CD_CPML_B200 = Acoustic2D_CD_CPML_B200" in jl_src
    for kind, n in (("AcousticCDCPMLWaveSimulation", 1), ("AcousticCDCPMLWaveSimulation", 2), ("AcousticCDCPMLWaveSimulation", 3),
                    ("AcousticVDStaggeredCPMLWaveSimulation", 1), ("AcousticVDStaggeredCPMLWaveSimulation", 2), ("ElasticIsoCPMLWaveSimulation", 2)):
        assert re.search(rf"SeismicWaves\.select_backend\(::CPMLBoundaryCondition, ::LocalGrid, ::Type\{{<:{kind}\{{<:FT, {n}\}}\}}, ::Type\{{Val\{{:B200\}}\}}\)", jl_src), (kind, n)
    # the two elastic forward methods differ by the three moment-tensor arguments, as at ela_forward.jl:53-69,127-144
    sigs = re.findall(r"function ela_forward_onestep_CPML!\((.*?)\)\n", jl_src, flags=re.S)
    assert len(sigs) == 2
    counts = sorted(len(split_top(s.split(";")[0])) for s in sigs)
    assert counts == [15, 18], counts


def test_level2_never_finalizes_an_immutable_simulation(jl_src):
    """the reference's simulations are immutable structs (acou_models.jl:68,311; ela_models.jl:177): finalizers may only hang on the
    extension's own mutable objects"""
    targets = re.findall(r"finalizer\([^,]+,\s*(\w+)\)", jl_src)
    assert targets and set(targets) <= {"a", "e"}, targets
    assert "mutable struct EngineField" in jl_src and "mutable struct B200Array" in jl_src
