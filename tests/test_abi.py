"""The C-ABI library loads and exports every symbol include/srt.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "srt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srt_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(srtlib):
    from simpleraytracing_b200 import capi
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(srtlib):
    for name in declared_symbols():
        assert hasattr(srtlib.lib, name), f"{name} declared in include/srt.h but not exported"


def test_version_and_defaults(srtlib):
    assert srtlib.lib.srt_version() == 100
    p = srtlib.default_params()
    assert p.incident == 80.0 and p.length_to_cm == 0.1 and p.repair_depth == 4 and p.kernel == 0 and p.flags == 0


def test_no_cpu_fallback_without_device(srtlib):
    """Without a usable sm_100 device srt_create must fail loudly (never fall back to the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    from simpleraytracing_b200 import Context, SrtError
    with pytest.raises(SrtError) as e:
        Context(0, srtlib)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_missing_library_raises():
    from simpleraytracing_b200 import capi
    with pytest.raises(FileNotFoundError):
        capi.load_library(ROOT / "does_not_exist.so")


def test_product_does_not_reference_oracle():
    """Nothing under simpleraytracing_b200/ or include/ may import, link or call oracle/."""
    bad = []
    for p in list((ROOT / "simpleraytracing_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cpp", ".cu", ".h", ".cuh", ".hpp"} or p.name == "Makefile":
            if p.is_file() and re.search(r"\boracle\b", p.read_text(errors="ignore")):
                bad.append(str(p.relative_to(ROOT)))
    assert not bad, bad
