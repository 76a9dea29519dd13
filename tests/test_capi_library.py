"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/rcb200.h declares, agrees with
the numpy mirrors of its structs, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from recastnavigation_b200 import capi
    L = capi.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "librcb200.so does not export %s" % n


def test_struct_layouts_match_header():
    from recastnavigation_b200 import capi, geom
    assert geom.FIELD_DTYPE.itemsize == 40
    assert C.sizeof(capi.Params) == 24
    assert C.sizeof(capi.Counts) == 72
    assert capi.FIELD_RESULT_DTYPE.itemsize == 20
    assert C.sizeof(capi.DeviceResults) == 7 * C.sizeof(C.c_void_p)
    assert C.sizeof(capi.Volume) == 40


def test_no_cpu_fallback_without_gpu():
    from recastnavigation_b200 import capi
    L = capi.lib()
    if L.rcb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(capi.RcbError):
        capi.Batch(0)
    h = C.c_void_p()
    rc = L.rcb_batch_create(0, None, C.byref(h))
    assert rc <= -1000 and not h.value
    assert b"no CPU fallback" in L.rcb_last_error()


def test_product_code_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under recastnavigation_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "recastnavigation_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "liboracle" not in text and "voxel_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn


def test_ptx_has_no_fused_multiply_add():
    """Bit-exactness of the clipper rests on the reference's non-fused IEEE operation order.  The kernels use
    __f*_rn intrinsics and the file is built with -fmad=false: the PTX must contain no fma/mad on floats
    (div.rn.f32 is expanded later by ptxas into a correctly rounded sequence, which is fine)."""
    import shutil
    import subprocess
    import tempfile
    from recastnavigation_b200 import build
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "k.ptx")
        flags = [f for f in build.NVCC_FLAGS if f not in ("-shared", "-Xcompiler", "-fPIC", "-lineinfo")]
        subprocess.check_call([build.nvcc_path()] + flags + ["-ptx", "-I", os.path.join(ROOT, "include"), "-o", out,
                                                             os.path.join(build.CSRC, "rcb_voxel.cu")])
        ptx = open(out).read()
    assert "div.rn.f32" in ptx and "mul.rn.f32" in ptx and "add.rn.f32" in ptx
    assert not re.search(r"\b(fma|mad)\.[a-z.]*f32", ptx), "fused multiply-add in the PTX"


def test_sweep_kernels_keep_address_rebuilds_out_of_their_steps():
    """Performance guard, not parity: the chamfer sweeps are chains of dependent steps, and nvcc rebuilds a shared-memory address
    from SR_CgaCtaId wherever it did not hoist it — two such reads per step cost k_wave_sweep a third of its time.
    k_tile_sweep (the main chain's kernel) must have none inside a loop; a step of k_wave_sweep (the SASS between two of its
    barriers) may keep the one uniform read the compiler still places there, no more."""
    import shutil
    import subprocess
    from recastnavigation_b200 import build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = build.build_rcb()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    if "Function :" not in sass:
        pytest.skip("no SASS in the library")
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            funcs[name].append(line)
    checked = 0
    for name, body in funcs.items():
        if "k_wave_sweepILb1ELb1E" in name:
            bars = [i for i, l in enumerate(body) if "BAR.SYNC" in l]
            per_step = [sum("SR_CgaCtaId" in l for l in body[a:b]) for a, b in zip(bars[1:-2], bars[2:-1])]   # between step barriers
            assert per_step and max(per_step) <= 1, "%s: %s SR_CgaCtaId reads in one wavefront step" % (name, max(per_step))
            checked += 1
        elif "k_tile_sweepIhE" in name:
            addr = [int(re.search(r"/\*([0-9a-f]+)\*/", l).group(1), 16) for l in body]
            loops = []
            for i, l in enumerate(body):
                m = re.search(r"\bBRA\b.*0x([0-9a-f]+)\s*;", l)
                if m and int(m.group(1), 16) < addr[i] and int(m.group(1), 16) in addr:
                    loops.append((addr.index(int(m.group(1), 16)), i))
            assert loops
            for i, l in enumerate(body):
                if "SR_CgaCtaId" in l:
                    assert not any(a <= i <= b for a, b in loops), "%s: SR_CgaCtaId read inside a loop" % name
            checked += 1
    assert checked >= 2
