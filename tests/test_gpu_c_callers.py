"""A plain C host program drives the drop-in library (tests/c/multi_gpu_formod.c): one device vs all visible devices
bit-identical, NCCL table broadcast inside the library, direct I/O on page-locked structs, concurrent formod_GPU callers,
oracle parity -- no Python in the data path.  Needs a GPU (uses all that are visible)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c", "multi_gpu_formod")


def test_c_caller_one_vs_all_devices(jr):
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "c")], check=True)
    ndev = jr.load_core().jrb_device_count()
    npk = max(8, 6 * ndev)
    r = subprocess.run([EXE, str(npk), "0"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["devices"] == ndev
    assert res["max_rel_err_rad"] <= 1e-6 and res["max_rel_err_tau"] <= 1e-6
    if ndev > 1:
        assert res["nccl_nranks"] == ndev
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"c_caller_{ndev}gpu.json"), "w") as f:
            f.write(line + "\n")
