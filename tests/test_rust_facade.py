"""The Rust facade (rust/felics-gpu) cannot be compiled here (no Rust toolchain in the image), so the
one thing that can be checked is checked: every function of include/felics_b200.h is declared in src/ffi.rs with the same
name, the same number of arguments and matching argument / return types, and nothing else is declared."""
import re

from conftest import ROOT

C_TO_RUST = {
    "int": "c_int", "void": "()", "size_t": "usize", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "double": "f64", "char": "c_char",
    "felics_ctx": "felics_ctx", "felics_header": "felics_header",
}


def c_functions():
    text = (ROOT / "include" / "felics_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(felics_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        out[name] = (norm_c(ret), [norm_c(a, drop_name=True) for a in args.split(",") if a.strip() and a.strip() != "void"])
    return out


def norm_c(decl, drop_name=False):
    decl = decl.strip()
    stars = decl.count("*")
    words = [w for w in re.sub(r"\*", " ", decl).split() if w not in ("const", "struct")]
    base = words[0]
    const_first = re.match(r"\s*const\b", decl) is not None
    # `const T *const *p` (pointer to const pointers) -> *const *const T ; `T *const *p` -> *const *mut T
    if stars == 2:
        inner = "*const" if const_first else "*mut"
        return f"*const {inner} {rust_base(base, 1)}" if "*const" in decl.replace(" ", "").replace("const*", "*const") or "* const" in decl else f"*mut {inner} {rust_base(base, 1)}"
    if stars == 1:
        return ("*const " if const_first else "*mut ") + rust_base(base, 1)
    return rust_base(base, 0)


def rust_base(base, stars):
    if base == "void":
        return "c_void" if stars else "()"
    return C_TO_RUST[base]


def rust_functions():
    text = (ROOT / "rust" / "felics-gpu" / "src" / "ffi.rs").read_text()
    block = text[text.index('extern "C" {'):]
    out = {}
    for name, args, ret in re.findall(r"pub fn (felics_[a-z0-9_]+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        out[name] = ((ret or "()").strip(), [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()])
    return out


def test_ffi_declares_exactly_the_header():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 25
    assert sorted(c) == sorted(r), f"only in the header: {sorted(set(c) - set(r))}; only in ffi.rs: {sorted(set(r) - set(c))}"
    for name, (cret, cargs) in c.items():
        rret, rargs = r[name]
        assert len(cargs) == len(rargs), f"{name}: {len(cargs)} arguments in C, {len(rargs)} in Rust"
        assert rret == cret, f"{name}: returns {cret} in C, {rret} in Rust"
        for i, (ca, ra) in enumerate(zip(cargs, rargs)):
            assert ca == ra, f"{name}, argument {i}: {ca} in C, {ra} in Rust"


def test_header_struct_and_codes_match():
    ffi = (ROOT / "rust" / "felics-gpu" / "src" / "ffi.rs").read_text()
    hdr = (ROOT / "include" / "felics_b200.h").read_text()
    for name, value in re.findall(r"#define (FELICS_(?:OK|ERR_[A-Z_]+|HEADER_BYTES))\s+\(?(-?\d+)\)?", hdr):
        assert re.search(rf"pub const {name}: \w+ = {value};", ffi), f"{name} = {value} missing in ffi.rs"
    fields = re.search(r"pub struct felics_header \{(.*?)\}", ffi, flags=re.S).group(1)
    assert [f.strip() for f in re.findall(r"pub (\w+: \w+)", fields)] == ["color_type: u8", "pixel_depth: u8", "width: u32", "height: u32"]


def test_crate_files_exist_and_say_they_are_not_compiled():
    crate = ROOT / "rust" / "felics-gpu"
    for f in ("Cargo.toml", "build.rs", "src/ffi.rs", "src/lib.rs", "README.md"):
        assert (crate / f).is_file(), f
    assert "Not compiled here" in (crate / "README.md").read_text()
    lib = (crate / "src" / "lib.rs").read_text()
    for item in ("pub trait CompressDecompress", "impl<T> CompressDecompress for ImageBuffer<Luma<T>, Vec<T>>", "impl<T> CompressDecompress for ImageBuffer<Rgb<T>, Vec<T>>",
                 "pub fn compress_image", "pub fn decompress_image", "pub fn read_header", "pub fn write_header", "pub enum DecompressionError"):
        assert item in lib, item
