"""CPU-side checks of the drop-in boundary: libffv1_b200.so loads, exports every symbol include/ffv1_b200.h declares,
reports errors the way libavcodec does, and REFUSES to compute without a GPU (no CPU fallback)."""
import ctypes, os, re, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200"))
import ffv1_b200

def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ffv1_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffv1b200_[a-z0-9_]+)\s*\(", src)))

def test_library_exports_every_declared_symbol():
    L = ffv1_b200.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libffv1_b200.so does not export " + n

def test_version_and_error_strings():
    L = ffv1_b200.lib()
    assert b"sm_100a" in L.ffv1b200_version()
    assert L.ffv1b200_strerror(-22) == b"Invalid argument"
    assert L.ffv1b200_strerror(-1094995529).startswith(b"Invalid data")

def test_no_cpu_fallback():
    """without a CUDA device the library must fail loudly instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.device_count()
    assert e.value.code == -542398533 and "no CPU fallback" in str(e.value)
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.FFV1Encoder(352, 288, "yuv420p", g=12, level=3, coder=1, slices=4)
    assert e.value.code == -542398533

def test_option_errors_do_not_need_a_gpu():
    """option resolution (encode_init, ffv1enc.c:669-1029) happens before any device work and keeps the reference's codes"""
    for kwargs, code in ((dict(slices=32), -38), (dict(level=1, slices=4), -22)):
        with pytest.raises(ffv1_b200.FFV1Error) as e:
            ffv1_b200.FFV1Encoder(1920, 1080, "yuv420p", g=16, coder=1, **kwargs)
        assert e.value.code == code
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.FFV1Encoder(1920, 1080, "rgb48le", g=16, level=3)
    assert e.value.code == -38

def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "codec.py" and False, "%s mentions the oracle" % f
