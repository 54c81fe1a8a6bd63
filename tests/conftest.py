"""pytest configuration: registers the `gpu` marker, puts the package dir on sys.path, builds the
oracle (test infrastructure) and the product library if they are missing."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "motion-estimated-video-trimmer_b200"
for p in (str(PKG), str(ROOT / "tests"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    if not (ROOT / "oracle" / "libmscan_oracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True)
    if not (PKG / "libmotionscan.so").exists():
        subprocess.run([sys.executable, str(PKG / "build.py")], check=True)
    yield


def _cuda_ok() -> bool:
    try:
        import ctypes as C

        import motionscan as ms

        n = C.c_int()
        return ms.lib().mscan_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu():
    return _cuda_ok()
