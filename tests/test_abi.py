"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/lac_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "lac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lac_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_binding():
    from lac_b200 import _ffi
    assert _declared() == sorted(_ffi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from lac_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_ffi.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    assert _ffi.lib().lac_abi_version() == _ffi.ABI_VERSION


def test_struct_sizes_match_header():
    from lac_b200 import _ffi
    assert _ffi.ENC_STATE_BYTES == 32 and _ffi.DEC_STATE_BYTES == 40


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from lac_b200 import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_ffi.LacError):
        _ffi.lib()


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lac_b200 import _ffi
    rc = _ffi.lib().lac_device_info(None, None, None, None)
    assert rc == _ffi.LAC_E_CUDA


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lac_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("#", "//", "*", "/*")))
                assert "import oracle" not in code and "from oracle" not in code, f
                assert "liblac_oracle" not in code, f


def test_built_library_has_no_short_cs2r_consumers():
    """tools/sass_hazard_scan.py: the SASS pattern that once corrupted the decoder's bit window must not
    reappear in a rebuild (needs cuobjdump; skipped where the CUDA toolkit is absent)."""
    import shutil
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    so = root / "lac_b200" / "_lib" / "liblac_b200.so"
    if shutil.which("cuobjdump") is None or not so.exists():
        pytest.skip("cuobjdump or the built library is not available")
    r = subprocess.run([sys.executable, str(root / "tools" / "sass_hazard_scan.py"), str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
