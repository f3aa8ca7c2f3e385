"""The C-ABI boundary without a GPU: libb200sdr.so loads and exports exactly what
include/b200sdr.h and include/rtlws_compat.h declare; no compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        name = m.group(1)
        if name in ("defined", "sizeof", "void", "int", "float", "double"):
            continue
        names.add(name)
    names.discard("rf_decimator_callback")
    return names


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    want = declared_functions("b200sdr.h") | declared_functions("rtlws_compat.h")
    assert len(want) >= 30
    missing = sorted(n for n in want if not hasattr(lib, n))
    assert not missing, f"declared in include/ but not exported: {missing}"
    assert set(pkg.EXPORTED_SYMBOLS) == want


def test_audio_library_exports_the_reference_audio_interface(pkg):
    """libb200audio.so = audio_main.h:6-14 (+ extensions), declared in include/rtlws_audio_compat.h."""
    lib = pkg.audio.lib()
    want = declared_functions("rtlws_audio_compat.h")
    assert {"audio_init", "audio_new_audio_available", "audio_get_audio_payload", "audio_fm_demodulator",
            "audio_close"} <= want
    missing = sorted(n for n in want if not hasattr(lib, n))
    assert not missing, f"declared in include/rtlws_audio_compat.h but not exported: {missing}"
    assert set(pkg.audio.EXPORTED_SYMBOLS) == want
    # host plumbing only: the arithmetic lives behind the C ABI of libb200sdr.so
    blob = open(pkg.audio.LIB_PATH, "rb").read()
    assert b"libb200sdr.so" in blob and b"libcudart" not in blob


def test_compat_types_have_reference_layout(pkg):
    # common_sp.h:7-20: 2-byte cmplx_u8, 8-byte cmplx_s32; resample.h:8-12: two cmplx_s32
    assert ctypes.sizeof(pkg.binding.CmplxS32) == 8
    assert ctypes.sizeof(pkg.CicDelayLine) == 16


def test_no_device_is_an_error_not_a_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    assert pkg.lib().b200_init(0) == -3                       # B200_ERR_CUDA
    assert b"cudaSetDevice" in pkg.lib().b200_last_error()
    with pytest.raises(pkg.B200Error):
        pkg.StreamRing(1, 5120)
    assert pkg.lib().spectrum_alloc(1024) is None              # the compat layer cannot allocate either


def test_product_does_not_link_or_import_the_oracle():
    # the product path may never route through oracle/: check sources and the built library
    pkg_dir = os.path.join(ROOT, "rtl-ws_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        if "build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"pyoracle|liboracle|libref_rtlws|orc_|import oracle|from oracle", text), \
                    f"{f} reaches for the oracle"
    blob = open(os.path.join(pkg_dir, "libb200sdr.so"), "rb").read()
    assert b"liboracle" not in blob and b"libref_rtlws" not in blob and b"orc_" not in blob


def test_headers_are_plain_c_and_the_c_hosts_link(pkg, tmp_path):
    """The reference is C: every header under include/ must compile as C99 on its own, and the C hosts of this repo
    (examples/virtual_dongles.c, tools/push_bench.c) must compile and link against the two libraries -- no GPU needed."""
    import subprocess
    pkg.lib()
    inc = os.path.join(ROOT, "include")
    lib_dir = os.path.join(ROOT, "rtl-ws_b200")
    for header in sorted(os.listdir(inc)):
        src = tmp_path / (header + ".c")
        src.write_text(f'#include "{header}"\nint main(void) {{ return 0; }}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + inc, "-c", str(src),
                        "-o", str(tmp_path / (header + ".o"))], check=True)
    for rel, libs in (("examples/virtual_dongles.c", ["-lb200sdr", "-lb200replay", "-lm"]),
                      ("examples/multi_gpu_gather.c", ["-lb200sdr"]),
                      ("tools/push_bench.c", ["-lb200sdr"])):
        subprocess.run(["gcc", "-O1", "-pthread", "-Wall", "-o", str(tmp_path / os.path.basename(rel)[:-2]),
                        os.path.join(ROOT, rel), "-I" + inc, "-L" + lib_dir, "-Wl,-rpath," + lib_dir] + libs, check=True)
