"""`not gpu`: the product library (built for sm_100a) loads and exports every entry point that
include/wmbus_b200.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def declared_functions():
    names = set()
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):
        hdr = open(os.path.join(inc, h)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        names |= set(re.findall(r"\b(wmb_[a-z_0-9]+)\s*\(", hdr))
    return sorted(names)


def exported_symbols(path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return sorted(l.split()[-1] for l in out.splitlines() if l.split()[1:2] and l.split()[1] in "TDBR")


def test_header_declares_expected_surface():
    names = declared_functions()
    for n in ["wmb_create", "wmb_destroy", "wmb_push", "wmb_push_device", "wmb_poll", "wmb_decode_frames",
              "wmb_take_lines", "wmb_process", "wmb_process_device", "wmb_host_alloc", "wmb_host_free", "wmb_seek",
              "wmb_set_line_window", "wmb_boundary_state", "wmb_pending_before"]:
        assert n in names


def test_library_exports_every_declared_symbol(pkg):
    path = pkg.library_path()
    if not os.path.exists(path):
        pkg.build()
    lib = C.CDLL(path)
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} missing from {path}"
    assert lib.wmb_abi_version() == 2
    # ... and nothing else: the dynamic symbol table is exactly the declared surface (csrc/wmb_exports.map)
    extra = [s for s in exported_symbols(path) if s not in declared_functions()]
    assert extra == [], f"exported but not declared in include/*.h: {extra}"


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(pkg):
    lib = pkg.load_library()
    o = pkg.opts_from_flags(lib, "")
    ctx = C.c_void_p()
    rc = lib.wmb_create(C.byref(o), 0, C.byref(ctx))
    assert rc == -2, "wmb_create must fail with WMB_E_NODEVICE when there is no CUDA device"
    assert b"no CPU fallback" in lib.wmb_last_error()


def test_cli_usage_and_version(pkg):
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.library_path()), "rtl_wmbus_b200")
    if not os.path.exists(exe):
        pkg.build()
    r = subprocess.run([exe, "-V"], capture_output=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 0 and r.stdout.startswith(b"rtl_wmbus: ")
    r = subprocess.run([exe, "-h"], capture_output=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1 and b"-d 2 set decimation rate to 2" in r.stdout      # rtl_wmbus.c:962-964
    r = subprocess.run([exe, "-p", "X"], capture_output=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1
    r = subprocess.run([exe, "-r", "1"], capture_output=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1
