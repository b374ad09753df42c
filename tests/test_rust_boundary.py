"""The Rust side of the boundary cannot be compiled in this image (no cargo / rustc), so it is checked mechanically:
`rust/corrla-b200-sys/src/lib.rs` must declare exactly the functions, structs and constants of `include/corrla_b200.h`,
with the same argument count, order and C-compatible types, and the safe wrapper `rust/corrla-b200/src/lib.rs` must call
them with the right number of arguments.  (The same call sequences are replayed in C by tools/c_abi_replay.c on a GPU.)"""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "corrla_b200.h").read_text()
SYS = (ROOT / "rust" / "corrla-b200-sys" / "src" / "lib.rs").read_text()
WRAP = (ROOT / "rust" / "corrla-b200" / "src" / "lib.rs").read_text()

# C type (normalised) -> the Rust FFI type that has the same ABI
C2RUST = {
    "int": "c_int", "double": "f64", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize",
    "const double*": "*const f64", "double*": "*mut f64", "const float*": "*const f32", "float*": "*mut f32", "void*": "*mut c_void", "int*": "*mut c_int",
    "const char*": "*const c_char", "corrla_ctx*": "*mut corrla_ctx", "corrla_comm*": "*mut corrla_comm",
    "const corrla_comm*": "*const corrla_comm", "corrla_ctx**": "*mut *mut corrla_ctx",
    "corrla_comm**": "*mut *mut corrla_comm", "const corrla_rsvd_opts*": "*const corrla_rsvd_opts",
    "corrla_rsvd_opts*": "*mut corrla_rsvd_opts", "corrla_timings*": "*mut corrla_timings",
    "unsigned char[128]": "*mut u8", "const unsigned char[128]": "*const u8", "void": "()",
}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def norm_ctype(t):
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*\s*", "*", t)
    return t


def split_c_param(p):
    """'const double* a' -> ('const double*', 'a');  'unsigned char id[128]' -> ('unsigned char[128]', 'id')"""
    p = p.strip()
    m = re.match(r"(.*?)(\w+)\s*(\[\d+\])?$", p)
    ty, name, arr = m.group(1), m.group(2), m.group(3) or ""
    return norm_ctype(ty) + arr, name


def c_functions():
    text = strip_comments(HEADER)
    out = {}
    for m in re.finditer(r"CORRLA_API\s+([\w\s\*]+?)\s*\b(corrla_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, params = norm_ctype(m.group(1)), m.group(2), m.group(3)
        plist = [] if params.strip() in ("", "void") else [split_c_param(x) for x in params.split(",")]
        out[name] = (ret, plist)
    return out


def c_struct(name):
    text = strip_comments(HEADER)
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S)
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty = re.match(r"(.*?)(\w+(?:\s*,\s*\w+)*)$", decl)
        base = norm_ctype(ty.group(1))
        for fname in ty.group(2).split(","):
            fields.append((fname.strip(), base))
    return fields


def rust_functions():
    m = re.search(r'extern "C" \{(.*?)\n\}', SYS, flags=re.S)
    out = {}
    for f in re.finditer(r"pub fn (\w+)\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", m.group(1), flags=re.S):
        params = [x.strip() for x in re.sub(r"\s+", " ", f.group(2)).split(",") if x.strip()]
        plist = [(p.split(":", 1)[1].strip(), p.split(":", 1)[0].strip()) for p in params]
        out[f.group(1)] = ((f.group(3) or "()").strip(), plist)
    return out


def rust_struct(name):
    m = re.search(r"#\[repr\(C\)\]\s*(?:#\[derive\([^\)]*\)\]\s*)?pub struct %s \{(.*?)\}" % name, SYS, flags=re.S)
    assert m, f"{name}: no #[repr(C)] struct in the -sys crate"
    return [(f.group(1), f.group(2).strip()) for f in re.finditer(r"pub (\w+):\s*([^,\n]+),", m.group(1))]


def test_every_function_is_declared_with_matching_types():
    cf, rf = c_functions(), rust_functions()
    assert sorted(cf) == sorted(rf), (sorted(set(cf) - set(rf)), sorted(set(rf) - set(cf)))
    assert len(cf) >= 22
    for name, (ret, params) in cf.items():
        rret, rparams = rf[name]
        assert C2RUST[ret] == rret or (ret == "const char*" and rret == "*const c_char"), (name, ret, rret)
        assert len(params) == len(rparams), (name, len(params), len(rparams))
        for (cty, cname), (rty, rname) in zip(params, rparams):
            assert cty in C2RUST, (name, cname, cty)
            assert C2RUST[cty] == rty, f"{name}: parameter {cname!r} is {cty!r} in C but {rty!r} (as {rname!r}) in Rust"


def test_structs_match_field_for_field():
    for sname in ("corrla_rsvd_opts", "corrla_timings"):
        cfields, rfields = c_struct(sname), rust_struct(sname)
        assert [f for f, _ in cfields] == [f for f, _ in rfields], sname
        for (fname, cty), (_, rty) in zip(cfields, rfields):
            assert C2RUST[cty] == rty, f"{sname}.{fname}: {cty!r} in C, {rty!r} in Rust"
    # and the ctypes mirror used by the tested Python front door has the same field order
    from corrla_rs_b200 import _ffi
    assert [f for f, _ in _ffi.RsvdOpts._fields_] == [f for f, _ in c_struct("corrla_rsvd_opts")]
    assert [f for f, _ in _ffi.Timings._fields_] == [f for f, _ in c_struct("corrla_timings")]


def test_status_codes_match():
    cvals = dict(re.findall(r"(CORRLA_\w+)\s*=\s*(-?\d+)", strip_comments(HEADER)))
    rvals = dict(re.findall(r"pub const (CORRLA_\w+): c_int = (-?\d+);", SYS))
    assert cvals == rvals and len(cvals) == 8


def test_wrapper_calls_have_the_declared_arity():
    """Every `sys::corrla_*(...)` call of the safe wrapper passes as many arguments as the header declares."""
    cf = c_functions()
    calls = 0
    for m in re.finditer(r"sys::(corrla_\w+)\s*\(", WRAP):
        name, i = m.group(1), m.end()
        depth, args, cur = 1, [], ""
        while depth:
            ch = WRAP[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append(cur); cur = ""
            else:
                cur += ch
            i += 1
        if cur.strip():
            args.append(cur)
        assert name in cf, name
        assert len(args) == len(cf[name][1]), (name, len(args), len(cf[name][1]))
        calls += 1
    assert calls >= 7          # rsvd, power_iter, par_matmul, dmdc, pod + status/error strings
