"""The drop-in boundary driven from plain C (tests/c_abi_harness.c): include/fos_b200.h compiles as C11, every call
links against libfos_b200.so, and -- on the GPU -- config 1 runs end to end and agrees with the oracle."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "c_abi_harness"


def build_harness():
    sys.path.insert(0, str(ROOT))
    import fos_b200  # noqa: F401  (builds libfos_b200.so if needed)
    from fos_b200 import _lib
    from oracle import fos_oracle
    _lib.lib_path()
    fos_b200.build.build()
    fos_oracle.build()
    libdir = ROOT / "firstordersolvers.jl_b200"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "c_abi_harness.c"), "-o", str(EXE), f"-L{libdir}", "-lfos_b200",
           f"-L{ROOT / 'oracle'}", "-lfos_oracle", "-lm", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{ROOT / 'oracle'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_header_compiles_as_c_and_links_and_has_no_cpu_fallback():
    exe = build_harness()
    r = subprocess.run([str(exe), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok")


@pytest.mark.gpu
def test_config1_end_to_end_from_plain_c():
    exe = build_harness()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok")
