"""T1: the binary interface.  CPU-only."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = [(2, 5), (3, 1), (8, 3), (32, 5), (128, 8), (100, 30)]


@pytest.mark.parametrize("ND,NG", DIMS)
def test_ctypes_mirror_matches_reference_layout(jr, refdrv, ND, NG):
    """every field offset and struct size of the ctypes mirrors == the reference compiled with the same -DND/-DNG"""
    if not refdrv.reference_available(ND, NG):
        pytest.skip("oracle/_ref not built for these dimensions")
    ref = refdrv.Reference(ND, NG)
    lay = ref.layout()
    ctl_t, atm_t, obs_t, tbl_t = jr.abi.structs(ND, NG)
    cls = {"ctl": ctl_t, "atm": atm_t, "obs": obs_t, "tbl": tbl_t}
    assert ref.dims()["ND"] == ND and ref.dims()["NG"] == NG
    checked = 0
    for key, val in lay.items():
        if key.startswith("sizeof("):
            name = key[7:-3]
            if name == "pos":
                continue
            assert C.sizeof(cls[name]) == val, key
        else:
            s, f = key.split(".")
            assert getattr(cls[s], f).offset == val, key
        checked += 1
    assert checked > 60


@pytest.mark.parametrize("ND,NG", DIMS)
def test_dropin_layer_struct_sizes(jr, ND, NG):
    """the C mirror structs of the drop-in layer (csrc/jr_structs.h) have the same sizes as the ctypes mirrors"""
    path = os.path.join(ROOT, "jurassic-gpu_b200", "lib", f"libjurassic_b200_dropin_nd{ND}_ng{NG}.so")
    assert os.path.exists(path), "drop-in layer not built"
    jr.load_core()
    lib = C.CDLL(path)
    dims = (C.c_int * 11)()
    sizes = (C.c_longlong * 4)()
    lib.jr_b200_dims(dims, sizes)
    assert list(dims)[:2] == [ND, NG]
    assert list(dims)[2:] == [9600, 1088, 1, 400, 40, 30, 304, 1201, 5000]
    ctl_t, atm_t, obs_t, tbl_t = jr.abi.structs(ND, NG)
    assert list(sizes) == [C.sizeof(ctl_t), C.sizeof(atm_t), C.sizeof(obs_t), C.sizeof(tbl_t)]
    for sym in ("formod_GPU", "jr_b200_init", "jr_b200_formod_batch", "jr_b200_finalize", "jr_b200_core_context"):
        assert hasattr(lib, sym), sym


def test_default_dims_tbl_size(jr):
    """sizeof(tbl_t) at the reference's default dimensions"""
    assert C.sizeof(jr.abi.structs(100, 30)[3]) == 8800822408


def _declared_functions(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(jrb?_[a-z0-9_]+|formod_GPU)\s*\(", txt))


def test_core_library_exports_every_declared_symbol(jr):
    lib = jr.load_core()
    declared = _declared_functions(os.path.join(ROOT, "include", "jurassic_b200.h"))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/jurassic_b200.h but not exported"
    assert set(jr.core.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.jrb_version()


def test_dropin_exports_every_declared_symbol(jr):
    jr.load_core()
    declared = _declared_functions(os.path.join(ROOT, "include", "jurassic_b200_dropin.h"))
    assert "formod_GPU" in declared
    for path in glob.glob(os.path.join(ROOT, "jurassic-gpu_b200", "lib", "libjurassic_b200_dropin_*.so")):
        lib = C.CDLL(path)
        for name in declared:
            assert hasattr(lib, name), f"{name} missing in {os.path.basename(path)}"


def test_no_gpu_means_loud_failure(jr):
    """there is no CPU fallback: without a CUDA device creating a context must raise"""
    lib = jr.load_core()
    if lib.jrb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(jr.JrbError, match="no usable CUDA device"):
        jr.Context(0)


def test_new_entry_points_fail_loudly_without_a_gpu(jr):
    """groups, page-locking and node-shared memory have no CPU fallback either"""
    import numpy as np
    lib = jr.load_core()
    if lib.jrb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(jr.JrbError):
        jr.Group(ndev=1)
    a = np.zeros(1024)
    assert lib.jrb_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes) != 0
    assert lib.jrb_host_is_registered(a.ctypes.data_as(C.c_void_p), a.nbytes) == 0
    p = C.c_void_p()
    assert lib.jrb_shared_alloc(b"/jrb_test_no_gpu", 4096, 1, C.byref(p)) != 0 and not p.value


def test_c_caller_builds_against_the_public_headers():
    """tests/c/multi_gpu_formod.c compiles against include/*.h + the struct mirror and links the drop-in library"""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(ROOT, "tests", "c", "multi_gpu_formod"))
