"""Host-side logic that needs no GPU: table packer properties, control handling, and the N > 1 path (sharding, table
broadcast, result gather) on two gloo ranks with the CPU oracle standing in for the device."""
import copy
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packer_properties(jr):
    ctl = jr.synth.control_limb_example()
    tbl = jr.synth.make_tables(ctl)
    info = jr.core.tables_pack_info(tbl, ctl.ng, ctl.nd)
    assert info["all_shared"] == 1 and info["monotone"] == 1 and info["gas_axes_same"] == 1
    per_gas = jr.core.tables_pack_info(jr.synth.make_tables(ctl, gas_axis_shift=True), ctl.ng, ctl.nd)
    assert per_gas["all_shared"] == 1 and per_gas["gas_axes_same"] == 0
    assert info["n_entries"] == int(tbl.nu[tbl.nu >= 2].sum())
    assert info["nbytes"] < 20e6
    # channel-dependent axes -> generic kernel
    jit = jr.synth.make_tables(ctl, axis_jitter=True)
    assert jr.core.tables_pack_info(jit, ctl.ng, ctl.nd)["all_shared"] == 0
    # a non-monotone column is detected (the hinted search is only equivalent to the bisection for monotone data)
    bad = copy.deepcopy(tbl)
    bad.eps[1, 5, 3, 10, 0] = bad.eps[1, 5, 3, 8, 0]
    assert jr.core.tables_pack_info(bad, ctl.ng, ctl.nd)["monotone"] == 0
    # missing tables do not break the shared-axes property
    part = jr.synth.make_tables(ctl, skip_pairs=[(0, 0), (3, 1)])
    assert jr.core.tables_pack_info(part, ctl.ng, ctl.nd)["all_shared"] == 1


def test_packed_blob_is_deterministic_and_position_independent(jr):
    ctl = jr.synth.control_nadir_example()
    tbl = jr.synth.make_tables(ctl)
    a = jr.core.tables_pack_host(tbl, ctl.ng, ctl.nd)
    b = jr.core.tables_pack_host(copy.deepcopy(tbl), ctl.ng, ctl.nd)
    assert a.size == jr.core.tables_pack_info(tbl, ctl.ng, ctl.nd)["nbytes"]
    assert np.array_equal(a, b)
    assert bytes(a[:8]) == b"RJBTBL01"
    # brackets: entry k of a column = (u_k, eps_k, u_k+1, eps_k+1)
    hdr = np.frombuffer(a[:200].tobytes(), dtype=np.uint64)
    off_col, off_brk = int(hdr[10]), int(hdr[11])  # TblHeader, csrc/jrb_device.cuh
    col = np.frombuffer(a[off_col:off_col + 8 * 36 * 12 * 3].tobytes(), dtype=np.uint32).reshape(36, 12, 3, 2)
    first, nu = col[7, 4, 1]
    assert nu == tbl.nu[0, 7, 4, 1]
    brk = np.frombuffer(a[off_brk + 16 * int(first): off_brk + 16 * int(first + nu)].tobytes(), dtype=np.float32).reshape(-1, 4)
    assert np.array_equal(brk[:, 0], tbl.u[0, 7, 4, :nu, 1]) and np.array_equal(brk[:, 1], tbl.eps[0, 7, 4, :nu, 1])
    assert np.array_equal(brk[:-1, 2], tbl.u[0, 7, 4, 1:nu, 1]) and np.array_equal(brk[:-1, 3], tbl.eps[0, 7, 4, 1:nu, 1])


def test_control_auto_continuum_switches(jr):
    """read_ctl switches a continuum off when no channel is in its range (src/jurassic.c:954-968)"""
    assert jr.synth.control_config_d().ctm_mask == 0b1100
    assert jr.synth.control_config_e().ctm_mask == 0b1111
    c = jr.Control(["O3"], [1000.0])
    assert (c.ctm_co2, c.ctm_h2o, c.ctm_n2, c.ctm_o2) == (1, 1, 0, 0) and c.ctm_mask == 0  # no CO2/H2O emitter
    assert jr.Control(["co2", "h2o"], [2200.0]).ctm_mask == 0b1110  # find_emitter is case-insensitive


def test_shard_range_partitions_exactly(jr):
    for total in (0, 1, 7, 115, 920):
        for world in (1, 2, 3, 8):
            spans = [jr.shard.shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert jr.shard.shard_range(920, 3, 8) == (345, 115)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rank_main(rank, world, port, tmp):
    import importlib
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    jr = importlib.import_module("jurassic-gpu_b200")
    import refdrv
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0, 950.0])
    total = 5
    # tables: packed on rank 0, broadcast, checked against a local packing on the receiver
    tbl = jr.synth.make_tables(ctl)
    blob = torch.from_numpy(jr.core.tables_pack_host(tbl, ctl.ng, ctl.nd)) if rank == 0 else None
    blob = jr.shard.broadcast_blob(dist, blob, 0, "cpu")
    assert hashlib.sha1(blob.numpy().tobytes()).hexdigest() == hashlib.sha1(jr.core.tables_pack_host(tbl, ctl.ng, ctl.nd).tobytes()).hexdigest()
    # rays: contiguous slice per rank
    first, count = jr.shard.shard_range(total, rank, world)
    orc = refdrv.Oracle()
    rows = []
    for g in range(first, first + count):
        p = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=4, dz=12.0, seed=900 + g)
        orc.formod(ctl, tbl, p)
        rows.append(np.concatenate([p.rad, p.tau], axis=1))
    mine = torch.from_numpy(np.concatenate(rows, axis=0))
    counts = [jr.shard.shard_range(total, r, world)[1] * 4 for r in range(world)]
    allrows = jr.shard.gather_rows(dist, mine, counts, 0, "cpu")
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), allrows.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_broadcast_and_gather(jr, oracle, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    ctl = jr.Control(["CO2", "H2O"], [792.0, 832.0, 950.0])
    tbl = jr.synth.make_tables(ctl)
    rows = []
    for g in range(5):
        p = jr.synth.limb_package(ctl, n_profiles=1, rays_per_profile=4, dz=12.0, seed=900 + g)
        oracle.formod(ctl, tbl, p)
        rows.append(np.concatenate([p.rad, p.tau], axis=1))
    want = np.concatenate(rows, axis=0)
    assert got.shape == want.shape == (20, 6)
    assert np.array_equal(got, want)  # 1-rank and 2-rank results are bit-identical and in package order
