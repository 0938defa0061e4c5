"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/vvb200.h declares,
and refuses to compute without a CUDA device (no fallback)."""
import ctypes as C

import pytest

from vietvoice_tts_b200 import _lib
from vietvoice_tts_b200.arch import TINY, FULL, ArchConfig, VVArch


def test_library_exports_every_declared_symbol(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libvvb200.so does not export {n}"


def test_arch_struct_matches_header():
    assert C.sizeof(VVArch) == 33 * 4
    c = FULL.to_c()
    assert c.dim == 1024 and c.depth == 22 and c.nfe == 32 and abs(c.cfg_strength - 2.0) < 1e-7


def test_no_cpu_fallback(lib):
    if lib.vv_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    carch = TINY.to_c()
    rc = lib.vv_engine_create(C.byref(carch), 0, None, C.byref(h))
    assert rc == -2                                   # VV_ERR_CUDA
    assert b"no CUDA device" in lib.vv_last_error()


def test_bad_arch_rejected(lib):
    h = C.c_void_p()
    bad = ArchConfig(dim=1000, heads=16).to_c() if False else TINY.to_c()
    bad.head_dim = 32
    rc = lib.vv_engine_create(C.byref(bad), 0, None, C.byref(h))
    assert rc == -1


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/vvb200.h must compile as C99 and link against the library"""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "vvb200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %d\\n", vv_version(), vv_device_count() >= 0); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "vietvoice-tts_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{root}/include", str(src),
                        "-o", str(exe), f"-L{libdir}", "-lvvb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split()[0] == "100"


def test_makefile_builds_every_cuda_source():
    """a kernel file that is not in the Makefile's SRCS silently stays out of libvvb200.so"""
    import glob
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vietvoice-tts_b200", "csrc")
    with open(os.path.join(csrc, "Makefile")) as f:
        srcs = re.search(r"^SRCS\s*:=\s*(.*)$", f.read(), re.M).group(1).split()
    on_disk = sorted(os.path.basename(p) for p in glob.glob(os.path.join(csrc, "*.cu")))
    assert sorted(srcs) == on_disk
