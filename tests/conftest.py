import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # test infrastructure is built on demand (seconds); the product library is built by __graft_entry__.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if os.path.exists("/root/reference/libavcodec/aacdec.c"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(os.path.join(ROOT, "ffmpeg-heaac_b200", "libheaac_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ffmpeg-heaac_b200", "csrc")])


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libheaac_ref.so"))


needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref (compiled reference) not available")
