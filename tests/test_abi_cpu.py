"""CPU: the C-ABI shared library builds, loads and exports every symbol declared in include/irfd_b200.h."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from speak_hack_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from speak_hack_b200 import _lib

    header = open(os.path.join(ROOT, "include", "irfd_b200.h")).read()
    declared = set(re.findall(r"\b(irfd_[a-z0-9_]+)\s*\(", header))
    declared.discard("irfd_stream_t")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_error_string(lib):
    assert lib.irfd_abi_version() == 1
    assert isinstance(lib.irfd_last_error(), bytes)


def test_invalid_arguments_return_error_codes_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so it is testable on a CPU-only box
    rc = lib.irfd_conv_gemm(None, 1, 8, 8, 64, None, 64, 3, None, None, 0, None, None, None, None, None, None, None, 0,
                            None)
    assert rc == -1 and b"null pointer" in lib.irfd_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.irfd_conv_gemm(p, 1, 8, 8, 60, p, 64, 3, p, None, 0, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"Cin" in lib.irfd_last_error()
    rc = lib.irfd_conv_gemm(p, 1, 8, 8, 64, p, 64, 5, p, None, 0, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"ksize" in lib.irfd_last_error()
    assert lib.irfd_conv_gemm_m_tiles(3, 8, 8) == 2
    assert lib.irfd_linear_bwd(p, p, p, p, 0.0, p, p, 0.0, 65, 8, 8, 1.0, 1.0, None) == -1  # batch rows > 64


def test_sass_uses_blackwell_tensor_and_tma_instructions():
    """cuobjdump evidence that the GEMM kernels are tcgen05 (UTC*MMA) + TMEM (LDTM) + TMA (UTMALDG/UTMASTG)."""
    import shutil
    import subprocess

    from speak_hack_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout


def test_sass_mma_issue_is_back_to_back():
    """Regression guard for the warp-converged elect.sync issue loops: every tcgen05 GEMM kernel issues
    its MMAs of one pipeline stage as consecutive UTCHMMA instructions.  With an `if (lane == 0)` issue region nvcc
    puts an ELECT / R2UR.BROADCAST loop (~14 instructions) in front of each one and no two are adjacent."""
    import re
    import shutil
    import subprocess

    from speak_hack_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    longest = {}
    name, run = None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name, run = m.group(1), 0
            continue
        if name is None or "/*" not in line or re.match(r"\s*/\* 0x", line):
            continue  # not an instruction line (the second half of each encoding sits on its own line)
        if "UTCHMMA" in line:
            run += 1
            longest[name] = max(longest.get(name, 0), run)
        else:
            run = 0
    gemm = {k: v for k, v in longest.items() if "gemm_kernel" in k or "halo_kernel" in k}
    assert len(gemm) >= 16, sorted(gemm)  # conv (3 tiles x 4 epilogues), conv halo, wgrad (3 tiles), wgrad halo
    for k, v in gemm.items():
        assert v >= 4, (k, v)  # four K=16 steps per 64-wide stage, back to back (eight / forty in the halo kernels)
