"""CPU-side checks of the boundary: the library loads, exports every symbol include/lbm2d.h declares,
the ctypes structs match the C layout, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT, make_config

pkg = importlib.import_module("01-lbm-2d_b200")
capi = importlib.import_module("01-lbm-2d_b200._capi")


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "lbm2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds():
    assert _declared_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = pkg.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.lbm_abi_version() == 2


def test_library_is_sm100a_only_and_has_no_cpu_path():
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_sass_shows_the_blackwell_paths_that_are_claimed():
    """TMA variant: UTMALDG / UTMASTG + mbarrier (SYNCS); default kernel: 64-bit L1-bypassing loads, warp shuffles, no
    local memory; PDL: ACQBULK / griddepcontrol lowered into the step kernel; strict kernel: packed fp32 adds (FADD2),
    and NO contracted multiply-add outside the division / square-root sequences."""
    sass = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur and "/*" in line:
            funcs[cur].append(line)
    def body(substr):
        names = [n for n in funcs if substr in n]
        assert names, substr
        return "\n".join(funcs[names[0]])
    tma = body("step_tma_kernelILb0ELb0E")
    assert "UTMALDG" in tma and "UTMASTG" in tma and "SYNCS" in tma
    hot = body("step_kernelILb0ELb0ELb0E")
    assert "LDG.E.64.STRONG.GPU" in hot and "SHFL" in hot and "STG.E.64" in hot
    assert "STL" not in hot and "LDL" not in hot, "the default step kernel must not touch local memory"
    # programmatic dependent launch (griddepcontrol.wait) and the early-start progress counter (acquire load that
    # invalidates L1, release = barrier + fence + 64-bit reduction)
    assert "ACQBULK" in hot and "CCTL.IVALL" in hot and "MEMBAR" in hot
    assert "RED.E.ADD.64" in hot or "ATOMG.E.ADD.64" in hot
    # strict arithmetic: two cells per packed pair; ptxas contracts FMUL2 + FADD2 into FFMA2 even with .rn
    # modifiers, which would change the rounding, so the multiplications are scalar: no FMUL2, and the only FFMA2 are
    # the 19 of the inline division / square-root sequences (2 + 3 + 3 | 2 | 3 | 2 | 2 + 2: Lane2 in lbm2d_device.cuh)
    for name in ("step_kernelILb1ELb0ELb0E", "step_kernelILb1ELb1ELb0E", "step_kernelILb1ELb0ELb1E"):
        strict = body(name)
        assert strict.count("FADD2") > 100, name
        assert strict.count("FFMA2") == 19 and "FMUL2" not in strict, (name, strict.count("FFMA2"))


def test_step_kernels_have_no_stack_frame_and_fit_their_occupancy_targets():
    """Every warp of the step kernel pays for whatever its prologue sets up (a thread executes ~500 instructions per step, and
    the kernel sits where issue rate and HBM meet): no stack frame on the plain steps -- the dense fallback for non-finite
    moments is inline on registers -- and register counts that keep 9 (strict) / 10 (fast) CTAs of 128 threads per SM."""
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", pkg.LIB_PATH], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function\s+(\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|LOCAL):(\d+)", line)}
            name = None
    def of(substr):
        names = [n for n in usage if substr in n]
        assert names, substr
        return usage[names[0]]
    for peer in ("Lb0E", "Lb1E"):   # single GPU / x-slab (peer-memory) instantiation
        strict, fast = of("step_kernelILb1ELb0ELb0E" + peer), of("step_kernelILb0ELb0ELb0E" + peer)
        assert strict["STACK"] == 0 and strict["LOCAL"] == 0 and strict["REG"] <= 56, strict      # 65536 / (9 * 128) = 56.9
        assert fast["STACK"] == 0 and fast["LOCAL"] == 0 and fast["REG"] <= 48, fast              # 65536 / (10 * 128) = 51.2, pinned at 48
        assert of("step_kernelILb1ELb1ELb0E" + peer)["STACK"] == 0                                  # strict EMIT step


def test_struct_layout_matches_c(tmp_path):
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "lbm2d.h"\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(LbmParams), offsetof(LbmParams, nu),"
        " offsetof(LbmParams, sponge_in), offsetof(LbmParams, sponge_strength), offsetof(LbmParams, bc_type),"
        " offsetof(LbmParams, bc_value), offsetof(LbmParams, arith), offsetof(LbmParams, kernel), offsetof(LbmParams, slab_x0));"
        "printf(\"%zu %zu %zu\\n\", sizeof(LbmDeviceView), offsetof(LbmDeviceView, nx_local), offsetof(LbmDeviceView, stream));return 0;}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    P, V = capi.LbmParams, capi.LbmDeviceView
    assert [int(x) for x in a.split()] == [C.sizeof(P), P.nu.offset, P.sponge_in.offset, P.sponge_strength.offset,
                                           P.bc_type.offset, P.bc_value.offset, P.arith.offset, P.kernel.offset, P.slab_x0.offset]
    assert [int(x) for x in b.split()] == [C.sizeof(V), V.nx_local.offset, V.stream.offset]


def test_missing_config_key_raises_keyerror_before_touching_the_gpu():
    cfg = make_config(32, 16)
    del cfg["domain_zones"]["sponge_out"]
    with pytest.raises(KeyError):
        pkg.LBM2D_MRT_LES(cfg)


def test_invalid_arguments_are_rejected_with_a_message():
    lib = pkg.load_library()
    h = C.c_void_p()
    assert lib.lbm_create(None, None, C.byref(h)) == 1
    assert b"null" in lib.lbm_last_error()
    p = capi.LbmParams()
    p.nx, p.ny, p.nx_global = 8, 2, 8
    assert lib.lbm_create(C.byref(p), None, C.byref(h)) == 1
    assert lib.lbm_init(None) == 1 and lib.lbm_run(None, 1) == 1


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.LbmError, match="no CUDA device|CUDA"):
        pkg.LBM2D_MRT_LES(make_config(32, 16), mask_data=np.zeros((32, 16), bool))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "01-lbm-2d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liblbm_oracle" not in text, f
