"""The drop-in boundary is a C ABI: include/b200_caesium.h must compile as strict C99, every declared symbol must link, and the
device-free calls must work from a plain C program (tests/c_abi_check.c) -- the stub a cgo / Rust-FFI maintainer would start from."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "caesium-clt_b200")


def test_header_is_c99_and_every_symbol_links(L, tmp_path):
    exe = str(tmp_path / "c_abi_check")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_check.c"),
           "-o", exe, "-L", PKG, "-lb200caesium", "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c-abi ok" in r.stdout


def test_c_check_covers_every_declared_function():
    hdr = open(os.path.join(ROOT, "include", "b200_caesium.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)) - {"b200_status", "b200_params"}
    src = open(os.path.join(ROOT, "tests", "c_abi_check.c")).read()
    missing = [f for f in sorted(declared) if "(fn)" + f not in src]
    assert not missing, missing
