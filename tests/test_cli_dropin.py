"""Row f4 / test T6: the reference's own command-line tool formod.c, linked (a) as shipped for the CPU and (b) with
-DhasGPU against the B200 drop-in library in place of GPUdrivers.o (oracle/Makefile target `cli`, recipe of
INTEGRATION.md).  Both read the same ctl / obs / atm / emissivity-table files and write rad.tab (6 significant digits)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU = os.path.join(ROOT, "oracle", "_ref", "formod_cpu_nd2_ng5")
B200 = os.path.join(ROOT, "oracle", "_ref", "formod_b200_nd2_ng5")


def _case(jr, tmp):
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    jr.synth.write_ascii_tables(ctl, tbl, tmp, "boxcar")
    pkg = jr.synth.example_package("limb", ctl)
    jr.synth.write_ctl(ctl, os.path.join(tmp, "limb.ctl"), "./boxcar")
    jr.synth.write_obs_tab(pkg, os.path.join(tmp, "obs.tab"))
    jr.synth.write_atm_tab(pkg, os.path.join(tmp, "atm.tab"))
    return ctl, tbl, pkg


def _run(exe, tmp, out, usegpu):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([exe, "limb.ctl", "obs.tab", "atm.tab", out, "USEGPU", str(usegpu)], cwd=tmp, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.array([l.split() for l in open(os.path.join(tmp, out)) if l.strip() and not l.startswith("#")], dtype=float)


def test_reference_cli_on_cpu_matches_oracle(jr, oracle, tmp_path):
    """the unmodified CLI (ASCII in, ASCII out) against the restatement oracle, at the 6 digits rad.tab holds"""
    if not os.path.exists(CPU):
        pytest.skip("oracle/_ref/formod_cpu_nd2_ng5 not built (needs /root/reference at build time)")
    ctl, tbl, pkg = _case(jr, str(tmp_path))
    rad = _run(CPU, str(tmp_path), "rad_cpu.tab", 0)
    oracle.formod(ctl, tbl, pkg)
    assert rad.shape == (66, 10 + 2 * ctl.nd)
    assert np.allclose(rad[:, 10:12], pkg.rad, rtol=5.1e-6) and np.allclose(rad[:, 12:14], pkg.tau, rtol=5.1e-6, atol=1e-30)
    assert np.allclose(rad[:, 7], pkg.tpz, rtol=5.1e-6)


@pytest.mark.gpu
def test_reference_cli_linked_against_dropin(jr, tmp_path):
    """`formod ... USEGPU 1` of the executable linked against libjurassic_b200_dropin == the CPU executable's rad.tab"""
    if not (os.path.exists(CPU) and os.path.exists(B200)):
        pytest.skip("oracle/_ref/formod_* not built (needs /root/reference at build time)")
    _case(jr, str(tmp_path))
    cpu = _run(CPU, str(tmp_path), "rad_cpu.tab", 0)
    gpu = _run(B200, str(tmp_path), "rad_gpu.tab", 1)
    assert cpu.shape == gpu.shape
    # both files hold 6 significant digits: allow one unit in the last printed digit
    keep = [c for c in range(cpu.shape[1]) if c != 8]
    assert np.allclose(gpu[:, keep], cpu[:, keep], rtol=2.1e-6, atol=1e-30)
    assert np.all(np.abs(gpu[:, 8] - cpu[:, 8]) < 1e-6)  # tangent longitude: round-off level values of atan2
    same = open(os.path.join(str(tmp_path), "rad_cpu.tab")).read() == open(os.path.join(str(tmp_path), "rad_gpu.tab")).read()
    print("rad.tab byte-identical:", same)


@pytest.mark.gpu
def test_reference_cli_benchmark_mode_reports_no_deviations(jr, tmp_path):
    """T6 in full: formod.c built with the reference's -DBENCHMARK_FORMOD and linked against the drop-in repeats the forward
    model USEGPU^2 times and compares every rad/tau of the repeats with the first call bit for bit (src/formod.c:71-176)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "formod_b200bench_nd2_ng5")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/formod_b200bench_nd2_ng5 not built (needs /root/reference at build time)")
    _case(jr, str(tmp_path))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([exe, "limb.ctl", "obs.tab", "atm.tab", "rad_bench.tab", "USEGPU", "3"], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "for 66 rays times 2 channels shows no deviations" in r.stdout, r.stdout[-2000:]
    assert "# ran 9 iterations for benchmark" in r.stdout
    assert "formod took" in r.stdout and "on the GPU" in r.stdout
