"""
Per-code kernel specialisation: compile the static decode kernels (H, L and the m <= 5 truth tables as
compile-time constants) for ONE code and attach them to its device object.

    code = CSSCode(h1, h2); code.specialize()           # "small-static(nvrtc:<hash>)"

Default (``compiler="nvrtc"``): ``qcss_code_specialize`` compiles the translation unit IN PROCESS with NVRTC
against the kernel headers embedded in ``libqcss.so`` and loads the cubin -- no toolkit, no subprocess; the cubin
is cached in ``quantum_css_codes_b200/jit/`` by content hash.  ``compiler="nvcc"`` is the round-1 route: the
library writes the translation unit (``qcss_code_spec_source``), nvcc builds it for sm_100a into
``jit/qcss_spec_<hash>.so`` and ``qcss_code_load_specialized`` routes the code's launches to it.  The three
descriptors built into the library (Steane, QRM-15, Golay-23) are the same mechanism run ahead of time.
"""

import ctypes
import hashlib
import os
import subprocess
import tempfile

from . import _native
from . import build as _build

JIT_DIR = os.path.join(_build.HERE, "jit")
_HEADERS = ("small_common.cuh", "decode.cuh", "core.cuh", "launch.h")


def spec_source(handle):
    lib = _native.load()
    needed = ctypes.c_int64()
    _native.check(lib.qcss_code_spec_source(handle, None, 0, ctypes.byref(needed)))
    buf = ctypes.create_string_buffer(needed.value)
    _native.check(lib.qcss_code_spec_source(handle, buf, needed.value, ctypes.byref(needed)))
    return buf.value.decode()


def _headers_digest():
    h = hashlib.sha1()
    for name in _HEADERS:
        with open(os.path.join(_build.CSRC, name), "rb") as fh:
            h.update(fh.read())
    return h


def build_spec(source):
    """Compile one specialised translation unit; returns (path of the shared object, tag)."""
    digest = _headers_digest()
    digest.update(source.encode())
    tag = digest.hexdigest()[:12]
    os.makedirs(JIT_DIR, exist_ok=True)
    so_path = os.path.join(JIT_DIR, f"qcss_spec_{tag}.so")
    if os.path.exists(so_path):
        return so_path, tag
    with tempfile.TemporaryDirectory(dir=JIT_DIR) as tmp:
        src = os.path.join(tmp, "spec.cu")
        with open(src, "w") as fh:
            fh.write(source)
        out = os.path.join(tmp, "spec.so")
        cmd = [_build.nvcc(), *_build.ARCH, "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
               "-shared", "-I", _build.CSRC, src, "-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for the specialised kernels:\n{res.stdout}\n{res.stderr}")
        os.replace(out, so_path)                       # atomic: concurrent builders race benignly
    return so_path, tag


def find_nvrtc():
    """Path of libnvrtc.so.12 for ``qcss_code_specialize``: the pip package torch depends on
    (nvidia/cuda_nvrtc/lib), else None so the library tries the loader's search path and /usr/local/cuda/lib64."""
    try:
        import nvidia.cuda_nvrtc as pkg
        base = os.path.join(os.path.dirname(pkg.__file__) if getattr(pkg, "__file__", None) else list(pkg.__path__)[0], "lib")
        for name in sorted(os.listdir(base)):
            if name.startswith("libnvrtc.so"):
                return os.path.join(base, name)
    except Exception:
        pass
    return None


def specialize(device_code, compiler="nvrtc"):
    """Build (or reuse) and attach the specialised kernels; returns the kernel family name."""
    lib = _native.load()
    if compiler == "nvrtc":
        os.makedirs(JIT_DIR, exist_ok=True)
        path = find_nvrtc()
        _native.check(lib.qcss_code_specialize(device_code.handle, path.encode() if path else None, JIT_DIR.encode()))
        return device_code.kernel_name()
    if compiler != "nvcc":
        raise ValueError("compiler must be 'nvrtc' or 'nvcc'")
    so_path, tag = build_spec(spec_source(device_code.handle))
    _native.check(lib.qcss_code_load_specialized(device_code.handle, so_path.encode(), tag.encode()))
    return device_code.kernel_name()
