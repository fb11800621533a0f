"""The drop-in boundary is a C ABI: the public headers must be plain C99 and a C program
must be able to drive a scheduler through them.  examples/minimal.c is built against the CPU
oracle (same ABI) and run."""
import subprocess

import pytest

from conftest import PORT_LIB, ROOT


@pytest.mark.parametrize("header", ["ydsched.h", "ydservice.h", "ydwire.h"])
def test_headers_are_plain_c99(header):
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-x", "c", "-fsyntax-only", "-"],
                       input=f'#include "{header}"\n', text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_c_example_runs_against_the_abi(tmp_path, port_lib):
    exe = tmp_path / "minimal"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "minimal.c"), "-o",
                        str(exe), str(PORT_LIB), f"-Wl,-rpath,{PORT_LIB.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and "request 4 -> timeout" in r.stdout
