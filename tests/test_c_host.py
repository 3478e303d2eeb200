"""The C ABI from a non-Python host: include/cmgan_b200.h must compile as plain C99 and a C program must link against the in-tree library
and use the module-level host-side queries (parameter table, workspace size, error reporting) -- no GPU involved."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_host_links_and_queries(tmp_path):
    import cmgan_b200  # noqa: F401  (makes sure the library is built)
    exe = str(tmp_path / "c_host")
    libdir = os.path.join(ROOT, "cmgan_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host.c"), "-o", exe,
           "-L" + libdir, "-lcmgan_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "4", "321"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "params 351 tensors" in out and "dense_encoder.conv_1.0.weight offset 0 numel 192" in out
    assert "complex_decoder.conv.bias" in out and "rejected F=200" in out
    ws = int(out.split("tf32: ")[1].split(" bytes")[0])
    assert ws > 100 << 20          # B = 4 x 2 s needs several hundred MB of scratch
