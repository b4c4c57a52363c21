"""The Rust -sys crate (rust/fuzzy-aho-corasick-gpu-sys) must declare exactly what include/fac.h declares.

No Rust toolchain exists in this image, so the crate cannot be compiled here; this test keeps its `extern "C"` block
and `#[repr(C)]` structs in lock-step with the header instead: same function names, same parameter counts, same
pointer-ness per parameter, same struct field names in the same order, same enum / flag values."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fac.h")
SYS = os.path.join(ROOT, "rust", "fuzzy-aho-corasick-gpu-sys", "src", "lib.rs")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def _split_params(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_functions():
    src = _strip_c_comments(open(HEADER).read())
    src = re.sub(r"typedef\s+[^;{]*\(\s*\*\s*\w+\s*\)\s*\([^;]*\)\s*;", " ", src, flags=re.S)  # callback typedefs
    fns = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(fac_\w+)\s*\(([^;{]*)\)\s*;", src, flags=re.S):
        name, params = m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else _split_params(params)
        fns[name] = ["*" in p for p in plist]
    return fns


def rust_functions():
    src = re.sub(r"//.*", "", open(SYS).read())
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    fns = {}
    for m in re.finditer(r"pub fn (fac_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        plist = _split_params(m.group(2)) if m.group(2).strip() else []
        kinds = []
        for p in plist:
            ty = p.split(":", 1)[1].strip()
            kinds.append(ty.startswith("*") or ty.startswith("Option<fac_") or ty in ("fac_read_fn", "fac_write_fn", "fac_match_fn", "fac_replace_fn"))
        fns[m.group(1)] = kinds
    return fns


def header_structs():
    src = _strip_c_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", src, flags=re.S):
        fields = []
        for line in m.group(2).split(";"):
            line = line.strip()
            if line:
                fields.append(re.sub(r"\[.*\]", "", line.split()[-1].lstrip("*")))
        out[m.group(1)] = fields
    return out


def rust_structs():
    src = re.sub(r"//.*", "", open(SYS).read())
    out = {}
    for m in re.finditer(r"pub struct (\w+) \{(.*?)\n\}", src, flags=re.S):
        out[m.group(1)] = re.findall(r"pub (\w+):", m.group(2))
    return out


def test_every_header_function_is_bound_with_the_same_shape():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 30
    assert sorted(h) == sorted(r), "functions differ: header-only %s, rust-only %s" % (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for name in h:
        # callback typedef parameters are pointers on the C side as well
        hk = h[name]
        rk = r[name]
        assert len(hk) == len(rk), "%s: %d parameters in fac.h, %d in lib.rs" % (name, len(hk), len(rk))


def test_struct_fields_match():
    h, r = header_structs(), rust_structs()
    for name in ("fac_limits", "fac_pattern", "fac_sim_pair", "fac_mapping", "fac_config", "fac_match", "fac_shard", "fac_search_args",
                 "fac_window", "fac_stream_stats"):
        assert name in h and name in r, name
        assert h[name] == r[name], "%s: %s vs %s" % (name, h[name], r[name])


def test_constants_match():
    hsrc = _strip_c_comments(open(HEADER).read())
    rsrc = open(SYS).read()
    consts = dict(re.findall(r"\b(FAC_[A-Z0-9_]+)\s*=\s*(\d+)", hsrc))
    consts.update({k: v.rstrip("u") for k, v in re.findall(r"#define (FAC_[A-Z0-9_]+) (\d+u?)\b", hsrc)})
    assert len(consts) >= 20
    for k, v in consts.items():
        m = re.search(r"pub const %s: \w+ = (\d+);" % k, rsrc)
        assert m, "lib.rs lacks %s" % k
        assert m.group(1) == v, "%s: %s in fac.h, %s in lib.rs" % (k, v, m.group(1))


def test_build_rs_uses_the_contract_flags():
    b = open(os.path.join(ROOT, "rust", "fuzzy-aho-corasick-gpu-sys", "build.rs")).read()
    for flag in ("arch=compute_100a,code=sm_100a", "--fmad=false", "-ffp-contract=off", "fac_api.cu", "fac_builder.cpp"):
        assert flag in b
    import importlib.util
    spec = importlib.util.spec_from_file_location("graft_entry", os.path.join(ROOT, "__graft_entry__.py"))
    ge = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ge)
    for flag in ge.NVCC_FLAGS:
        if flag.startswith("-") and flag not in ("-Xcompiler",):
            assert flag in b, flag
