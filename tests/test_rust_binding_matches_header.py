"""The Rust FFI crate cannot be compiled in this image, so its hand-written `extern "C"` blocks are held to the C header
here: every function it declares exists in include/fftconv_b200.h with the same number of parameters, `#[repr(C)]`
structs have the header's field count, and every entry point the crate's types call is one the GPU tests exercise
through ctypes (the signature table of fft_convolution_b200/_lib.py)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
RUST = (ROOT / "rust" / "fft_convolution_b200_sys" / "src" / "lib.rs").read_text()
HEADER = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "fftconv_b200.h").read_text(), flags=re.S)


def _rust_externs():
    out = {}
    for block in re.findall(r'extern "C" \{(.*?)\n\}', RUST, flags=re.S):
        block = re.sub(r"//[^\n]*", "", block)
        for name, args in re.findall(r"pub fn (fcb_\w+)\s*\((.*?)\)\s*(?:->[^;]+)?;", block, flags=re.S):
            args = args.strip()
            out[name] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def _header_functions():
    out = {}
    for name, args in re.findall(r"\b(fcb_\w+)\s*\(([^;{}]*?)\)\s*;", HEADER, flags=re.S):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_every_rust_extern_matches_the_header():
    rust, hdr = _rust_externs(), _header_functions()
    assert len(rust) >= 45
    for name, n in rust.items():
        assert name in hdr, f"{name} is declared in the Rust crate but not in the header"
        assert hdr[name] == n, f"{name}: {n} parameters in Rust, {hdr[name]} in the header"


def test_repr_c_structs_have_the_headers_fields():
    def rust_fields(struct):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, RUST, flags=re.S).group(1)
        return len(re.findall(r"pub \w+\s*:", re.sub(r"//[^\n]*", "", body)))

    def c_fields(struct):
        body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^{}]*)\}\s*(\w+);", HEADER))[struct]
        n = 0
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                n += decl.count(",") + 1  # `const float *add0, *add1` declares two
        return n

    assert rust_fields("FcbEngineDesc") == c_fields("fcb_engine_desc") == 6
    assert rust_fields("FcbEpilogue") == c_fields("fcb_epilogue") == 6
    assert rust_fields("FcbOptions") == c_fields("fcb_options") == 6


def test_every_entry_the_rust_types_call_is_bound_by_the_ctypes_table():
    from fft_convolution_b200 import _lib
    used = set(re.findall(r"\b(fcb_\w+)\s*\(", re.sub(r'extern "C" \{.*?\n\}', "", RUST, flags=re.S)))
    used.discard("fcb_last_error")
    assert used and used <= set(_lib.SIGNATURES), sorted(used - set(_lib.SIGNATURES))


def test_all_three_reference_types_and_the_batch_entry_are_bound():
    for t in ("CudaFFTConvolver", "CudaTwoStageFFTConvolver", "CudaCrossfadeConvolver"):
        assert re.search(r"impl Convolution for %s\b" % t, RUST), t
        assert re.search(r"impl Clone for %s\b" % t, RUST) and re.search(r"impl Drop for %s\b" % t, RUST), t
    assert "pub struct CudaFFTConvolverBatch" in RUST
