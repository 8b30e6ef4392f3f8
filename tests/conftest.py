import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "ffmpeg-ffv1-p-frames_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")

@pytest.fixture(scope="session")
def ref():
    """the unmodified reference build (oracle/_ref); skipped when it is not present"""
    from oracle import ffv1_ref
    if not ffv1_ref.available():
        try:
            ffv1_ref.build()
        except Exception:
            pass
    if not ffv1_ref.available():
        pytest.skip("oracle/_ref/libffv1ref.so not built (needs /root/reference)")
    return ffv1_ref
