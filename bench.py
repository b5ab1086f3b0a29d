#!/usr/bin/env python3
"""bench.py -- ray*channel radiances per second of the EGA forward model on N B200 GPUs (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W             our CUDA path (default)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU, strong scaling
  python bench.py --impl reference ...                      the reference's own CPU path (oracle/_ref) on the host cores

Workload = "config D" of BASELINE.json / SURVEY.md 8d, the configuration the metric is quoted on: synthetic limb sounder,
920 packages of 17 profiles x 64 rays (1088 rays, the capacity of one obs_t) = 1 000 960 rays, 32 channels
(785..816 cm^-1), 5 gases, CO2+H2O continua.  It fits one GPU, so N = 1 runs all of it and N ranks share it (contiguous
package slices, "scaling": "strong").  A step = one pass of the hot path (ray tracing -> column densities ->
EGA/continua/Planck/accumulation) over all packages.

e2e   : through the reference-facing C entry jr_b200_formod_batch(ctl_t*, atm_t*[], obs_t*[]) with the reference's HOST
        structs, page-locked once (jr_b200_pin_packages): per step the device gathers the inputs from the structs over PCIe
        and stores every ray's results into obs_t while the kernels run.  With N ranks the library itself holds the NCCL
        communicator (jr_b200_dist_init: tables broadcast from rank 0), every rank owns a contiguous obs slice, and the obs_t
        array lives in node-shared page-locked memory, so rank 0 has all radiances when the step ends.
value : device path, inputs resident in HBM (jrb_run_staged on the same context), wall clock over K steps between
        synchronisations, max over ranks.
After the headline line is computed a short pass over "config E" (nadir, 128 channels, 8 gases, 4 continua) is measured the
same way and embedded as extra.config_e.
"""
import argparse
import copy
import gc
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "ray-channel radiances/sec"

# workloads of BASELINE.json; "d" is the one the metric is quoted on (the bench line), "e" is embedded as extra.config_e
WORKLOADS = {
    "d": dict(name="config D: synthetic limb sounder, 920 packages x (17 profiles x 64 rays) = 1 000 960 rays, 32 channels, 5 gases, CO2+H2O continua",
              dims=(32, 5), packages=920),
    "e": dict(name="config E: synthetic AIRS-like nadir, 460 packages x (16 profiles x 68 footprints) = 500 480 footprints, 128 channels, 8 gases, 4 continua",
              dims=(128, 8), packages=460),
}


def algorithmic_bytes_per_ray_channel(sbar, ng):
    """SURVEY.md 8d / BASELINE.md 3:  B = S(ng*176 + 16) + 16"""
    return sbar * (ng * 176 + 16) + 16


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_dict(args, workload, world):
    """identical for our arm and the reference arm (the driver compares them)"""
    W = WORKLOADS[workload]
    total = args.packages if (workload == args.config and args.packages > 0) else W["packages"]
    return {"workload": W["name"] if total == W["packages"] else W["name"] + f" -- REDUCED to {total} packages",
            "packages_total": total, "rays_total": total * 1088, "channels": W["dims"][0], "gases": W["dims"][1],
            "l2_policy": "inputs larger than L2 (line-of-sight scratch rewritten every step, 70 KB per ray; tables 0.2-1.3 GB)",
            "parallelism": f"contiguous package slices over {world} GPU(s), one process per GPU, tables broadcast once by NCCL inside the library"}


_REAL_STDOUT = None


def emit(obj):
    """print the result line on the real stdout"""
    sys.stdout.flush()
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_packages(jr, ctl, first, count, workload="d"):
    if workload == "e":
        return [jr.synth.nadir_package(ctl, seed=20240518 + first + i) for i in range(count)]
    return [jr.synth.limb_package(ctl, seed=20240517 + first + i) for i in range(count)]


def make_control(jr, workload):
    return jr.synth.control_config_e() if workload == "e" else jr.synth.control_config_d()


def fill_tbl_struct(tbl_t, tbl):
    buf = (C.c_char * C.sizeof(tbl_t))()
    t = tbl_t.from_buffer(buf)
    g, P, T, U, d = tbl.dims
    for name, idx in (("np", (slice(0, g), slice(0, d))), ("nt", (slice(0, g), slice(0, P), slice(0, d))),
                      ("nu", (slice(0, g), slice(0, P), slice(0, T), slice(0, d))), ("p", (slice(0, g), slice(0, P), slice(0, d))),
                      ("t", (slice(0, g), slice(0, P), slice(0, T), slice(0, d))),
                      ("u", (slice(0, g), slice(0, P), slice(0, T), slice(0, U), slice(0, d))),
                      ("eps", (slice(0, g), slice(0, P), slice(0, T), slice(0, U), slice(0, d)))):
        np.ctypeslib.as_array(getattr(t, name))[idx] = getattr(tbl, name)
    np.ctypeslib.as_array(t.sr)[:, :d] = tbl.sr
    np.ctypeslib.as_array(t.st)[:] = tbl.st
    return t, buf


class _StructFiller:
    """host structs of the reference (ctypes mirrors) from the flat containers"""

    def __init__(self, ctl_t, atm_t, obs_t):
        self.ctl_t, self.atm_t, self.obs_t = ctl_t, atm_t, obs_t

    def ctl(self, ctl):
        c = self.ctl_t()
        c.ng, c.nd, c.nw = ctl.ng, ctl.nd, ctl.nw
        for i, e in enumerate(ctl.emitters):
            c.emitter[i].value = e.encode()
        for i in range(ctl.nd):
            c.nu[i] = ctl.nu[i]; c.window[i] = int(ctl.window[i])
        c.hydz, c.ctm_co2, c.ctm_h2o, c.ctm_n2, c.ctm_o2 = ctl.hydz, ctl.ctm_co2, ctl.ctm_h2o, ctl.ctm_n2, ctl.ctm_o2
        c.ip, c.refrac, c.rayds, c.raydz, c.write_bbt, c.formod, c.useGPU = ctl.ip, ctl.refrac, ctl.rayds, ctl.raydz, ctl.write_bbt, ctl.formod, 1
        c.fov = b"-"
        return c

    def atm(self, pkg, a=None):
        a = a if a is not None else self.atm_t()
        n = pkg.n_atm
        a.np = n
        for name, src in (("time", pkg.atm_time), ("z", pkg.z), ("lon", pkg.lon), ("lat", pkg.lat), ("p", pkg.p), ("t", pkg.t)):
            np.ctypeslib.as_array(getattr(a, name))[:n] = src
        np.ctypeslib.as_array(a.q)[: pkg.ng, :n] = pkg.q[: pkg.ng]
        np.ctypeslib.as_array(a.k)[: pkg.nw, :n] = pkg.k[: pkg.nw]
        return a

    def obs(self, pkg, o=None):
        o = o if o is not None else self.obs_t()
        n = pkg.n_rays
        o.nr = n
        for name in ("time", "obsz", "obslon", "obslat", "vpz", "vplon", "vplat"):
            np.ctypeslib.as_array(getattr(o, name))[:n] = getattr(pkg, name)
        return o


def reference_arm(args, jr):
    """--impl reference: the reference's own CPU implementation (oracle/_ref, unmodified CPUdrivers.c path, OpenMP over
    all host cores) on a bounded sample of the same workload per step.  Falls back to the C restatement if oracle/_ref
    was not built."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    ND, NG = WORKLOADS[args.config]["dims"]
    ctl = make_control(jr, args.config)
    tbl = jr.synth.make_tables(ctl)
    npk = args.ref_packages
    pkgs = make_packages(jr, ctl, 0, npk, args.config)
    if refdrv.reference_available(ND, NG):
        ref = refdrv.Reference(ND, NG)
        ref.lib.jrref_set_threads(len(os.sched_getaffinity(0)))  # all host cores (torchrun exports OMP_NUM_THREADS=1)
        kind, cores = "reference", ref.threads()
        c = ref.make_ctl(ctl)
        tstruct, keep = fill_tbl_struct(ref.tbl_t, tbl)
        atms = [ref.make_atm(p) for p in pkgs]
        obss = [ref.make_obs(p) for p in pkgs]

        def step():
            for a, o in zip(atms, obss):
                ref.formod_tbl(c, a, o, C.addressof(tstruct))
    else:
        orc = refdrv.Oracle()
        kind, cores = "port", orc.threads()
        work = [copy.deepcopy(p) for p in pkgs]

        def step():
            for w in work:
                orc.formod(ctl, tbl, w)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    rc = sum(p.n_rays for p in pkgs) * ctl.nd
    value = rc / dt
    sample = (f"the first {npk} of the workload's packages ({rc} ray-channels) per step, formod_CPU call sequence per package, "
              f"{cores} OpenMP threads; the rate is per ray-channel and independent of the number of packages")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ray-channels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(args, args.config, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "ray-channels/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "ray-channels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })
    return 0


class Dist:
    """torch.distributed is the launcher plumbing only: barrier, max/sum over ranks, broadcast of the NCCL unique id"""

    def __init__(self, torch, dist, rank, world, local):
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def allmax(self, x):
        return self.reduce(x, "MAX")

    def allsum(self, x):
        return self.reduce(x, "SUM")

    def bcast_bytes(self, b, n):
        if not self.dist:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8))
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())


def load_traffic(workload, packages_per_launch, kernel):
    """DRAM traffic and the secondary pipe numbers of the dominant kernel come from one committed ncu --set full capture; they
    are attached only when that capture is of this very launch (same workload, packages per launch and kernel), else null"""
    p = os.path.join(ROOT, "profiles", "ega_traffic.json")
    if not os.path.exists(p):
        return None, None, None
    try:
        for e in json.load(open(p)).get("captures", []):
            if e.get("workload") == workload and e.get("packages_per_launch") == packages_per_launch and e.get("kernel_contains", "") in kernel:
                return e.get("dram_bytes_per_launch"), e.get("capture"), e.get("secondary")
    except Exception:
        pass
    return None, None, None


def run_workload(args, jr, D, workload, steps, warmup, cpu_seconds, with_gather_check):
    """measures one workload: e2e through the drop-in C entry (host structs) and the device path on the same context"""
    rank, world, local = D.rank, D.world, D.local
    ND, NG = WORKLOADS[workload]["dims"]
    cfg = config_dict(args, workload, world)
    total_pk = cfg["packages_total"]
    ctl = make_control(jr, workload)
    ctl_t, atm_t, obs_t, tbl_t = jr.abi.structs(ND, NG)
    io = _StructFiller(ctl_t, atm_t, obs_t)
    core = jr.load_core()
    lib = C.CDLL(os.path.join(ROOT, "jurassic-gpu_b200", "lib", f"libjurassic_b200_dropin_nd{ND}_ng{NG}.so"))
    lib.jr_b200_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.jr_b200_dist_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
    lib.jr_b200_dist_unique_id.argtypes = [C.c_char_p]
    lib.jr_b200_dist_gather.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int]
    lib.jr_b200_dist_gather.restype = None
    lib.jr_b200_formod_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    lib.jr_b200_formod_batch.restype = None
    lib.jr_b200_pin_packages.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
    lib.jr_b200_core_context.restype = C.c_void_p
    lib.jr_b200_core_group.restype = C.c_void_p
    lib.jr_b200_shared_alloc.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
    lib.jr_b200_shared_alloc.restype = C.c_void_p
    lib.jr_b200_shared_free.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.jr_b200_shared_free.restype = None

    # ---- tables: built on rank 0, packed once, broadcast by the library's own NCCL communicator ----
    t_tab0 = time.perf_counter()
    tbl = tstruct = keep = None
    c = io.ctl(ctl)
    c.MPIlocalrank = local
    if rank == 0:
        tbl = jr.synth.make_tables(ctl)
        tstruct, keep = fill_tbl_struct(tbl_t, tbl)
    if world == 1:
        lib.jr_b200_init(C.addressof(c), C.addressof(tstruct), local)
    else:
        uid = C.create_string_buffer(128)
        if rank == 0:
            assert lib.jr_b200_dist_unique_id(uid) == 0, "NCCL unique id"
        uid = C.create_string_buffer(D.bcast_bytes(uid.raw, 128), 128)
        lib.jr_b200_dist_init(C.addressof(c), C.addressof(tstruct) if rank == 0 else None, rank, world, uid, local, 0)
    t_tables = time.perf_counter() - t_tab0
    gst = jr.abi.GroupStats()
    core.jrb_group_get_stats(C.c_void_p(lib.jr_b200_core_group()), C.byref(gst))
    ctx = jr.Context(handle=lib.jr_b200_core_context())  # the drop-in's own context (device 0 of its group, lane 0)

    # ---- this rank's contiguous slice of packages; obs_t of ALL packages in node-shared page-locked memory when N > 1 ----
    first, count = jr.shard.shard_range(total_pk, rank, world)
    counts = [jr.shard.shard_range(total_pk, r, world)[1] for r in range(world)]
    firsts = [jr.shard.shard_range(total_pk, r, world)[0] for r in range(world)]
    pkgs = make_packages(jr, ctl, first, count, workload)
    atms = [io.atm(p) for p in pkgs]
    shm_name, shm_ptr, shm_bytes = None, None, 0
    if world > 1:
        shm_name = f"/jrb_bench_{os.environ.get('MASTER_PORT', '0')}_{workload}".encode()
        shm_bytes = C.sizeof(obs_t) * total_pk
        if rank == 0:
            shm_ptr = lib.jr_b200_shared_alloc(shm_name, shm_bytes, 1)
        D.barrier()
        if rank != 0:
            shm_ptr = lib.jr_b200_shared_alloc(shm_name, shm_bytes, 0)
        assert shm_ptr, "node-shared page-locked memory could not be set up"
        all_obs = [obs_t.from_address(shm_ptr + k * C.sizeof(obs_t)) for k in range(total_pk)]
        obss = all_obs[first:first + count]
        for p, o in zip(pkgs, obss):
            io.obs(p, o)
    else:
        obss = [io.obs(p) for p in pkgs]
        all_obs = obss
    ap_ = (C.c_void_p * count)(*[C.addressof(x) for x in atms])
    op_ = (C.c_void_p * count)(*[C.addressof(x) for x in obss])
    assert lib.jr_b200_pin_packages(ap_, op_ if world == 1 else None, count) == 0, "page-locking the host structs failed"
    D.barrier()

    # ---- end to end through the drop-in call, host structs in and out ----
    def e2e_step():
        lib.jr_b200_formod_batch(C.addressof(c), ap_, op_, count)
        if D.dist:  # rank 0 may read the node-shared obs_t array once every rank is through
            D.dist.barrier()

    for _ in range(max(warmup, 1)):
        e2e_step()
    sampler = ClockSampler(local)
    gc.collect()
    gc.disable()  # no collector pauses inside the timed regions
    D.barrier()
    sampler.start()
    t0 = time.perf_counter()
    per_step = []
    for _ in range(steps):
        ts = time.perf_counter()
        e2e_step()
        per_step.append((time.perf_counter() - ts) * 1e3)
    D.barrier()
    dt_e = D.allmax(time.perf_counter() - t0)
    st = ctx.stats()
    assert st["io_direct"] == 1, "the e2e path did not run in direct (page-locked) mode"
    my_rc, my_rays, my_los = st["n_ray_channels"], st["n_rays"], st["n_los_points"]
    tot_rc, tot_rays, tot_los = D.allsum(my_rc), D.allsum(my_rays), D.allsum(my_los)
    e2e = {"value": tot_rc / (dt_e / steps), "unit": "ray-channels/s", "ms_per_step": dt_e / steps * 1e3,
           "h2d_bytes_per_step": int(D.allsum(st["h2d_bytes"])), "d2h_bytes_per_step": int(D.allsum(st["d2h_bytes"])),
           "ms_per_step_min": float(min(per_step)), "ms_per_step_max": float(max(per_step)),
           "host_ms": {"stage": st["host_ms_stage"], "mask_scan": st["host_ms_pack"], "device": st["ms_total_device"], "scatter": st["host_ms_scatter"]},
           "io": "direct: inputs gathered by the device from the page-locked atm_t/obs_t, results stored by the kernels into obs_t rows (no copy phase)",
           "api": f"jr_b200_formod_batch(ctl_t*, atm_t*[], obs_t*[], n) with host structs (ND={ND}, NG={NG})" +
                  ("; obs_t[] of all ranks in node-shared page-locked memory (jr_b200_shared_alloc), NCCL communicator of the library: %d ranks" % gst.nccl_nranks if world > 1 else "")}
    first_rad = np.ctypeslib.as_array(obss[0].rad)[: pkgs[0].n_rays, : ctl.nd].copy()

    # ---- device path on the same context: inputs resident in HBM ----
    av = (jr.abi.AtmView * count)(*[jr.abi.atm_view_of(a, ND, NG) for a in atms])
    ov = (jr.abi.ObsView * count)(*[jr.abi.obs_view_of(o, ND) for o in obss])
    ctx.stage_views(count, av, ov)
    for _ in range(max(warmup, 1)):
        ctx.run_staged()
    ctx.stage_views(count, av, ov)  # resets the accumulated kernel times
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.run_staged()  # synchronous: returns after the step's kernels have finished
    D.barrier()
    dt = D.allmax(time.perf_counter() - t0)
    clocks = sampler.stop()
    gc.enable()
    st = ctx.stats()
    launches = int(D.allsum(st["cum_launches"]))
    ms_per_step = dt / steps * 1e3
    value = tot_rc / (ms_per_step / 1e3)
    if not np.array_equal(np.ctypeslib.as_array(obss[0].rad)[: pkgs[0].n_rays, : ctl.nd], first_rad, equal_nan=True):
        raise SystemExit("bench: device-path results differ from the drop-in call's")

    # ---- roofline of the dominant kernel (EGA), measured live with CUDA events on its stream ----
    sbar = tot_los / max(tot_rays, 1)
    bytes_rc = algorithmic_bytes_per_ray_channel(sbar, ctl.ng)
    peak, peak_src = hbm_peak()
    n_launch = max(int(st["cum_ega_launches"]), 1)
    ega = st["cum_ms_ega"] / n_launch                       # average duration of one launch
    rc_per_launch = my_rc * st["cum_runs"] / n_launch       # ray-channels one launch processes
    achieved = rc_per_launch * bytes_rc / (ega / 1e3) / 1e9
    if not st["ega_kernel_variant"]:
        kernel = "ega_generic_kernel"
    elif st["ega_tiled"]:
        kernel = (f"ega_tiled_kernel<mask={st['ega_ctm_mask']}> ({st['ega_ngb']} gases, 32 channels per warp, tiles of 6 segments, "
                  f"{'lock-step' if st['ega_phase_lock'] else 'free-running'} CTAs)")
    else:
        kernel = (f"ega_fast_kernel<mask={st['ega_ctm_mask']}> ({st['ega_ngb']} gases, {st['ega_channels_per_warp']} channels per warp, "
                  f"{'lock-step' if st['ega_phase_lock'] else 'free-running'} CTAs)")
    traffic, traffic_src, secondary = load_traffic(workload, count // max(st["n_chunks"], 1), kernel)
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "secondary": secondary,
                "peak_source": peak_src, "algorithmic_bytes_per_ray_channel": bytes_rc, "mean_los_points": sbar,
                "ray_channels_per_launch": rc_per_launch, "launches_per_step": st["n_chunks"],
                "kernel_ms": ega, "raytrace_ms_per_step": st["cum_ms_raytrace"] / max(st["cum_runs"], 1),
                "kernel_share_of_step": st["cum_ms_ega"] / max(st["cum_runs"], 1) / ms_per_step}

    # ---- parity of the measured data + CPU baseline: the reference's own CPU path on the host cores (rank 0) ----
    cpu, parity = None, None
    if rank == 0 and cpu_seconds > 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refdrv
        if refdrv.reference_available(ND, NG):
            ref = refdrv.Reference(ND, NG)
            ref.lib.jrref_set_threads(len(os.sched_getaffinity(0)))
            cc = ref.make_ctl(ctl)
            kind, cores = "reference", ref.threads()

            def run1(p):
                o = ref.make_obs(p)
                ref.formod_tbl(cc, ref.make_atm(p), o, C.addressof(tstruct))
                return (np.ctypeslib.as_array(o.rad)[: p.n_rays, : ctl.nd].copy(), np.ctypeslib.as_array(o.tau)[: p.n_rays, : ctl.nd].copy())
        else:
            orc = refdrv.Oracle()
            kind, cores = "port", orc.threads()

            def run1(p):
                q = copy.deepcopy(p)
                orc.formod(ctl, tbl, q)
                return q.rad, q.tau
        e_rad = e_tau = 0.0
        n_cmp = 0

        def compare(k_global, pkg, r_ref, t_ref):
            nonlocal e_rad, e_tau, n_cmp
            o = all_obs[k_global]
            rad = np.ctypeslib.as_array(o.rad)[: pkg.n_rays, : ctl.nd]
            tau = np.ctypeslib.as_array(o.tau)[: pkg.n_rays, : ctl.nd]
            e_rad = max(e_rad, float(np.max(np.abs(rad - r_ref) / (np.abs(r_ref) + 1e-12 * np.max(np.abs(r_ref))))))
            e_tau = max(e_tau, float(np.max(np.abs(tau - t_ref) / (np.abs(t_ref) + 1e-12))))
            n_cmp += 1

        if world == 1:
            run1(pkgs[0])  # warm-up
            done, el = 0, 0.0
            while done < min(len(pkgs), 64) and el < cpu_seconds:
                t0 = time.perf_counter()
                r_ref, t_ref = run1(pkgs[done])
                el += time.perf_counter() - t0  # (the comparison below is not part of the CPU timing)
                compare(done, pkgs[done], r_ref, t_ref)
                done += 1
            cpu = {"value": done * 1088 * ctl.nd / el, "unit": "ray-channels/s", "cores": cores, "kind": kind,
                   "sample": f"the first {done} packages of the workload ({done*1088*ctl.nd} ray-channels) in {el:.1f} s, formod_CPU call sequence per package, serial ray tracing as in the reference"}
            where = f"the first {done} packages"
        else:  # SURVEY.md section 7, T8 on the measured multi-rank data: package 0 of EVERY rank's slice against the reference
            for r in range(world):
                p = make_packages(jr, ctl, firsts[r], 1, workload)[0]
                r_ref, t_ref = run1(p)
                compare(firsts[r], p, r_ref, t_ref)
            where = f"package 0 of each of the {world} ranks' slices, read on rank 0 from the node-shared obs_t array"
        parity = {"packages": n_cmp, "ray_channels": int(n_cmp * 1088 * ctl.nd), "which": where, "against": kind, "max_rel_err_rad": e_rad,
                  "max_rel_err_tau": e_tau, "tolerance": 1e-6, "ok": bool(e_rad <= 1e-6 and e_tau <= 1e-6)}

    # ---- N > 1: the library's NCCL gather (ncclSend/ncclRecv of the compact device results to rank 0, then D2H + scatter into a
    # private obs_t array on rank 0) must deliver the very bits the node-shared array holds; its cost is reported beside e2e ----
    gather = None
    if world > 1 and with_gather_check:
        cnt = (C.c_int * world)(*counts)
        priv, pp = None, None
        if rank == 0:
            priv = [obs_t() for _ in range(total_pk)]
            for k in range(total_pk):
                priv[k].nr = all_obs[k].nr
            pp = (C.c_void_p * total_pk)(*[C.addressof(x) for x in priv])
        lib.jr_b200_formod_batch(C.addressof(c), ap_, op_, count)
        tg = []
        for _ in range(2):
            D.barrier()
            t0 = time.perf_counter()
            lib.jr_b200_dist_gather(pp, cnt, world, 0)
            D.barrier()
            tg.append(D.allmax(time.perf_counter() - t0) * 1e3)
        core.jrb_group_get_stats(C.c_void_p(lib.jr_b200_core_group()), C.byref(gst))
        if rank == 0:
            same = True
            for r in range(1, world):
                for k in range(firsts[r], firsts[r] + counts[r]):
                    same = same and np.array_equal(np.ctypeslib.as_array(priv[k].rad), np.ctypeslib.as_array(all_obs[k].rad), equal_nan=True) \
                        and np.array_equal(np.ctypeslib.as_array(priv[k].tau), np.ctypeslib.as_array(all_obs[k].tau)) \
                        and np.array_equal(np.ctypeslib.as_array(priv[k].tpz), np.ctypeslib.as_array(all_obs[k].tpz))
            gather = {"ms": float(min(tg)), "bytes": int(gst.gather_bytes), "ms_scatter_on_root": float(gst.ms_gather_scatter),
                      "bit_identical_to_shared_array": bool(same),
                      "what": "jr_b200_dist_gather: grouped ncclSend/ncclRecv of rad/tau/tangent points to rank 0, page-locked D2H, scatter into obs_t[]"}
            if parity is not None:
                parity["nccl_gather_bit_identical"] = bool(same)
                parity["ok"] = bool(parity["ok"] and same)
            del priv

    # ---- the reference's own symbol: formod_GPU(ctl, atm, obs) on ONE package (what every unmodified JURASSIC caller does,
    # src/formod.c:65,100): latency mode of the library (one-gas passes + cooperative tracer), bit-identical to the batch ----
    single = None
    if rank == 0 and workload == "d":
        lib.formod_GPU.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.formod_GPU.restype = None
        o1 = obs_t()
        C.memmove(C.addressof(o1), C.addressof(obss[0]), C.sizeof(obs_t))
        a1p, o1p = (C.c_void_p * 1)(C.addressof(atms[0])), (C.c_void_p * 1)(C.addressof(o1))
        lib.jr_b200_pin_packages(a1p, o1p, 1)
        for _ in range(3):
            lib.formod_GPU(C.addressof(c), C.addressof(atms[0]), C.addressof(o1))
        t0 = time.perf_counter()
        for _ in range(20):
            lib.formod_GPU(C.addressof(c), C.addressof(atms[0]), C.addressof(o1))
        wall = (time.perf_counter() - t0) / 20 * 1e3
        s1 = ctx.stats()
        same = bool(np.array_equal(np.ctypeslib.as_array(o1.rad), np.ctypeslib.as_array(obss[0].rad), equal_nan=True) and
                    np.array_equal(np.ctypeslib.as_array(o1.tau), np.ctypeslib.as_array(obss[0].tau)))
        single = {"what": "formod_GPU(ctl_t*, atm_t*, obs_t*) on one package: 1088 rays x %d channels x %d gases" % (ctl.nd, ctl.ng),
                  "ms_wall_per_call": wall, "ms_device": s1["ms_total_device"], "ms_raytrace": s1["ms_raytrace"], "ms_ega": s1["ms_ega"],
                  "gas_blocks": s1["ega_gas_blocks"], "ray_channels_per_s": 1088 * ctl.nd / (wall / 1e3),
                  "bit_identical_to_batch": same}
        if not same:
            raise SystemExit("bench: formod_GPU on one package differs from the same package inside the batch")

    out = {"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "single_package": single,
           "gpu_launches": launches, "clocks": clocks, "config": cfg, "gather": gather,
           "tables": {"seconds_incl_generation": t_tables, "blob_bytes": int(st["table_blob_bytes"]), "nccl_nranks": int(gst.nccl_nranks)},
           "rays_per_gpu": int(my_rays), "los_chunks_per_step": int(st["n_chunks"])}
    if parity is not None and not parity["ok"]:
        emit({"error": "parity failed", "workload": workload, "parity": parity})
        raise SystemExit(f"bench: device results differ from the CPU reference: {parity}")

    lib.jr_b200_finalize()
    core.jrb_host_unregister_all()
    if shm_ptr:
        D.barrier()
        lib.jr_b200_shared_free(shm_name, shm_ptr, shm_bytes, 1 if rank == 0 else 0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="d", choices=["d", "e"], help="workload of the bench line (d = the one the metric is quoted on)")
    ap.add_argument("--packages", type=int, default=0, help="total packages of 1088 rays (default: the full workload, 920 for d); profiling only")
    ap.add_argument("--ref-packages", type=int, default=2, help="packages per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config-e", action="store_true", help="skip the embedded config E pass")
    ap.add_argument("--e-packages", type=int, default=0, help="total packages of the embedded config E pass (default: all 460)")
    ap.add_argument("--no-gather-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    # libraries (NCCL, the reference's printf's) write to stdout; keep fd 1 clean for the single JSON line
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    jr = importlib.import_module("jurassic-gpu_b200")
    if args.impl == "reference":
        return reference_arm(args, jr)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = Dist(torch, dist, rank, world, local)

    cpu_s = 0.0 if args.no_cpu_baseline else args.cpu_baseline_seconds
    res = run_workload(args, jr, D, args.config, args.steps, args.warmup, cpu_s, not args.no_gather_check)
    extra = {}
    if args.config == "d" and not args.no_config_e:
        a2 = copy.copy(args)
        a2.packages = args.e_packages
        a2.config = "e"
        e = run_workload(a2, jr, D, "e", 3, 1, min(cpu_s, 6.0), False)
        extra["config_e"] = {"metric": METRIC, "value": e["value"], "ms_per_step": e["ms_per_step"], "e2e": e["e2e"], "roofline": e["roofline"],
                             "parity": e["parity"], "cpu_baseline": e["cpu_baseline"], "clocks": e["clocks"], "config": e["config"],
                             "steps": 3, "warmup": 1, "gpu_launches": e["gpu_launches"]}
    if rank == 0:
        out = {"metric": METRIC, "value": res["value"], "unit": "ray-channels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
               "roofline": res["roofline"], "cpu_baseline": res["cpu_baseline"], "parity": res["parity"],
               "extra": dict(extra, single_package=res["single_package"], gather=res["gather"], tables=res["tables"], rays_per_gpu=res["rays_per_gpu"],
                             los_chunks_per_step=res["los_chunks_per_step"])}
        if out["cpu_baseline"] is not None and res["parity"] is not None:
            out["cpu_baseline"]["parity"] = res["parity"]
        emit(out)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
