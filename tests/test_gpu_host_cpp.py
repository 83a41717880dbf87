"""The C++ host mirror of the Go API (host/csvplus.hpp) run end to end on the GPU: the reference's
TestSimpleDataSource / TestSimpleUniqueJoin / TestWriteFile / TestErrors scenarios in host/example.cpp."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_example():
    exe = os.path.join(ROOT, "host", "example")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host example ok" in r.stdout
