"""rust/cuda.rs cannot be compiled in this image (no cargo / rustc), so its `#[repr(C)]` structs and `extern "C"` block are
checked against include/pfv_b200.h by parsing both: same struct names, same fields in the same order with the same C
types, same function names, arities, parameter types and return types, same constants.  A field added to the header, a
swapped pair or a stale signature fails here instead of corrupting memory in a crate built elsewhere."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# canonical type names both sides are mapped to
C2CANON = {
    "int": "i32", "unsigned": "u32", "float": "f32", "char": "i8", "size_t": "usize", "void": "void",
    "uint8_t": "u8", "int8_t": "i8", "uint16_t": "u16", "int16_t": "i16", "uint32_t": "u32", "int32_t": "i32",
    "uint64_t": "u64", "int64_t": "i64",
}
RS2CANON = {"c_int": "i32", "c_char": "i8", "c_void": "void", "usize": "usize", "f32": "f32"}


def strip_comments_c(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def canon_c_type(t):
    """'const uint8_t *' -> '*const u8'; 'pfv_ctx **' -> '*mut *mut pfv_ctx'; 'const int32_t (*)[64]' -> '*const [i32;64]'"""
    t = t.strip()
    m = re.match(r"^(const\s+)?(\w+)\s*\(\s*\*\s*\)\s*\[(\d+)\]$", t)
    if m:
        return f"*{'const' if m.group(1) else 'mut'} [{C2CANON.get(m.group(2), m.group(2))};{m.group(3)}]"
    const = bool(re.match(r"^const\b", t))
    t = re.sub(r"^const\s+", "", t)
    t = re.sub(r"^struct\s+", "", t)
    stars = t.count("*")
    base = t.replace("*", "").strip()
    base = C2CANON.get(base, base)
    out = base
    for i in range(stars):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def canon_rs_type(t):
    t = t.strip()
    m = re.match(r"^\[(\w+);\s*(\d+)\]$", t)
    if m:
        return f"[{RS2CANON.get(m.group(1), m.group(1))};{m.group(2)}]"
    m = re.match(r"^\*(const|mut)\s+(.*)$", t)
    if m:
        return f"*{m.group(1)} {canon_rs_type(m.group(2))}"
    return RS2CANON.get(t, t)


def parse_header():
    s = strip_comments_c(open(os.path.join(ROOT, "include", "pfv_b200.h")).read())
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(\w+)\s*;", s, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # 'uint32_t width, height' style lists
            head = re.match(r"^((?:const\s+)?\w+)\s*(.*)$", decl, flags=re.S)       # the stars belong to the declarators
            base, rest = head.group(1), head.group(2)
            for part in rest.split(","):
                part = part.strip()
                arr = re.match(r"^(\**)\s*(\w+)\s*\[(\d+)\]$", part)
                if arr:
                    fields.append((arr.group(2), f"[{canon_c_type(base + arr.group(1))};{arr.group(3)}]"))
                else:
                    pm = re.match(r"^(\**)\s*(\w+)$", part)
                    fields.append((pm.group(2), canon_c_type(base + pm.group(1))))
        structs[m.group(2)] = fields
    funcs = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?\w+\s*\**)\s*(pfv_\w+)\s*\(([^;{]*?)\)\s*;", s, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in re.split(r",(?![^\[]*\])", args):
                a = a.strip()
                pa = re.match(r"^(const\s+)?(\w+)\s*\(\s*\*\s*\w*\s*\)\s*\[(\d+)\]$", a)       # const int32_t (*qtables)[64]
                if pa:
                    params.append(canon_c_type(f"{pa.group(1) or ''}{pa.group(2)} (*)[{pa.group(3)}]"))
                    continue
                arr = re.match(r"^(.*?)(\w+)((?:\[\d+\])+)$", a, flags=re.S)                  # int32_t out[4][64]: the first dimension decays
                if arr:
                    dims = re.findall(r"\[(\d+)\]", arr.group(3))[1:]
                    inner = canon_c_type(re.sub(r"^const\s+", "", arr.group(1)))
                    for d in reversed(dims):
                        inner = f"[{inner};{d}]"
                    params.append(("*const " if re.match(r"^\s*const\b", arr.group(1)) else "*mut ") + inner)
                    continue
                pm = re.match(r"^(.*?)(\w+)$", a, flags=re.S)
                params.append(canon_c_type(pm.group(1)))
        funcs[name] = (canon_c_type(ret), params)
    consts = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(PFV_\w+)\s+(0x[0-9a-fA-F]+|\d+)u?\b", s)}
    for m in re.finditer(r"\b(PFV_\w+)\s*=\s*(-?\d+)", s):
        consts[m.group(1)] = int(m.group(2))
    return structs, funcs, consts


def parse_rust():
    s = open(os.path.join(ROOT, "rust", "cuda.rs")).read()
    s = re.sub(r"//[^\n]*", "", s)
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\](?:\s*#\[[^\]]*\])*\s*pub struct (\w+)\s*\{(.*?)\}", s, flags=re.S):
        fields = []
        for f in re.finditer(r"pub\s+(\w+)\s*:\s*([^,\n]+?)\s*,", m.group(2)):
            fields.append((f.group(1), canon_rs_type(f.group(2))))
        structs[m.group(1)] = fields
    funcs = {}
    block = re.search(r'extern "C"\s*\{(.*?)\n\}', s, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = []
        args = m.group(2).strip()
        if args:
            for a in re.split(r",(?![^\[]*\])", args):
                a = a.strip()
                if a:
                    params.append(canon_rs_type(a.split(":", 1)[1]))
        funcs[m.group(1)] = (canon_rs_type(m.group(3)) if m.group(3) else "void", params)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (PFV_\w+)\s*:\s*\w+\s*=\s*(\d+)\s*;", s)}
    return structs, funcs, consts


def test_rust_structs_match_the_header():
    hs, _, _ = parse_header()
    rs, _, _ = parse_rust()
    assert rs, "no #[repr(C)] structs found in rust/cuda.rs"
    for name, fields in rs.items():
        assert name in hs, f"rust/cuda.rs declares {name}, include/pfv_b200.h does not"
        assert fields == hs[name], f"{name}: rust {fields} != header {hs[name]}"
    for name in ("pfv_mbhdr", "pfv_geometry", "pfv_decode_job", "pfv_encode_job", "pfv_decode_job_sparse", "pfv_encode_job_sparse"):
        assert name in rs, f"{name} is not bound in rust/cuda.rs"


def test_rust_extern_block_matches_the_header():
    _, hf, _ = parse_header()
    _, rf, _ = parse_rust()
    assert len(rf) >= 10
    for name, (ret, params) in rf.items():
        assert name in hf, f"rust/cuda.rs binds {name}, which include/pfv_b200.h does not declare"
        hret, hparams = hf[name]
        assert ret == hret, f"{name}: return type rust {ret} != header {hret}"
        assert params == hparams, f"{name}: parameters rust {params} != header {hparams}"


def test_rust_constants_match_the_header():
    _, _, hc = parse_header()
    _, _, rc = parse_rust()
    assert rc
    for name, v in rc.items():
        assert name in hc and hc[name] == v, f"{name}: rust {v} != header {hc.get(name)}"


def test_every_header_function_is_exported_by_the_library():
    """(the other direction of test_abi_cpu.py's symbol table) every function the header declares exists in libpfv_b200.so"""
    import ctypes
    _, hf, _ = parse_header()
    lib = ctypes.CDLL(os.path.join(ROOT, "pretty_fast_video_b200", "libpfv_b200.so"))
    missing = [n for n in hf if not hasattr(lib, n)]
    assert not missing, missing
