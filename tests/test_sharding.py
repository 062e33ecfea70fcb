"""The N>1 path on CPU: two gloo ranks run bench.py's sharding logic (PDIFS schedule, GOP-aligned frame ranges,
histogram all-reduce, chunk-size exchange, container assembly on rank 0). The per-rank chunk encoder is the
oracle here (no GPU in this test); the assembled file must equal the single-process encode byte for byte."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from agmv_testlib import LZSS, OPT, QUALITY, oracle, oracle_encode, ptr, synth_frames  # noqa: E402

W, H, N_PER_RANK = 48, 32, 32
u32p, u64p, i32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)


def _rank(rank, world, port, light, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = oracle()
    lib.orc_encode_frames.restype = C.c_long
    lib.orc_encode_frames.argtypes = [u32p, C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_uint32, u32p, u32p, C.c_int, u8p, C.c_size_t]
    opt, quality = (OPT["III"], QUALITY["LOW"]) if light else (OPT["I"], QUALITY["LOW"])
    n_total = N_PER_RANK * world
    frames = np.ascontiguousarray(synth_frames(W, H, N_PER_RANK, seed=7, first=1 + rank * N_PER_RANK))
    # pass 1: local histogram, all-reduce (the palette is global)
    mc = lib.orc_max_clr(quality)
    hist = np.zeros(mc + 1, np.uint64)
    lib.orc_histogram_add(ptr(hist, u64p), ptr(frames, u32p), frames.size, quality)
    t = torch.from_numpy(hist.view(np.int64))
    dist.all_reduce(t)
    p0, p1 = np.zeros(256, np.uint32), np.zeros(256, np.uint32)
    lib.orc_build_palette(ptr(hist, u64p), quality, opt, ptr(p0, u32p), ptr(p1, u32p))
    # pass 2: this rank's GOP-aligned range of the global schedule
    sa_all, sb_all = bench.pdifs_schedule(n_total, light)
    e0, e1 = bench.shard_ranges(len(sa_all), world, 12 if light else 4)[rank]
    sa = np.ascontiguousarray(sa_all[e0:e1] - rank * N_PER_RANK)
    sb = np.ascontiguousarray(np.where(sb_all[e0:e1] >= 0, sb_all[e0:e1] - rank * N_PER_RANK, -1).astype(np.int32))
    assert sa.min() >= 0 and max(sa.max(), sb.max()) < N_PER_RANK
    out = np.zeros(W * H * 4 * max(1, e1 - e0) + 64, np.uint8)
    ln = lib.orc_encode_frames(ptr(frames, u32p), W, H, ptr(sa, i32p), ptr(sb, i32p), e1 - e0, e0, ptr(p0, u32p), ptr(p1, u32p), 1,
                               ptr(out, u8p), out.size)
    assert ln >= 0
    # chunk-image sizes of every rank -> offsets; payload gather on rank 0
    sizes = torch.zeros(world, dtype=torch.int64)
    sizes[rank] = ln
    dist.all_reduce(sizes)
    pad = int(sizes.max())
    mine = torch.zeros(pad, dtype=torch.uint8)
    mine[:ln] = torch.from_numpy(out[:ln])
    gathered = [torch.zeros(pad, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        hdr = bytearray(38 + 1536)
        hdr[0:4] = b"AGMV"
        hdr[8:12] = W.to_bytes(4, "little")
        hdr[12:16] = H.to_bytes(4, "little")
        hdr[16], hdr[17] = 1, 1
        hdr[36] = 16
        for i in range(256):
            for k, p in ((0, p0), (1, p1)):
                c = int(p[i])
                hdr[38 + k * 768 + 3 * i: 38 + k * 768 + 3 * i + 3] = bytes([(c >> 16) & 255, (c >> 8) & 255, c & 255])
        create_n, fps = n_total - 1, 24
        images = [gathered[r][: int(sizes[r])].numpy().tobytes() for r in range(world)]
        data = bench.assemble_container(bytes(hdr), images, len(sa_all), bench.fps_field(n_total, create_n, fps, light))
        q.put(data)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("light", [True, False])
def test_two_rank_sharded_encode_is_byte_identical(light):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500 + (1 if light else 0)
    procs = [ctx.Process(target=_rank, args=(r, world, port, light, q)) for r in range(world)]
    for p in procs:
        p.start()
    data = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_total = N_PER_RANK * world
    whole = synth_frames(W, H, n_total, seed=7)
    opt = OPT["III"] if light else OPT["I"]
    ref = oracle_encode(whole, n_total - 1, 24, opt, QUALITY["LOW"], LZSS)
    assert data == ref


def test_schedule_counts_match_baseline_configs():
    assert len(bench.pdifs_schedule(212, False)[0]) == 104      # C1: 212 source frames, HEAVY -> 104
    assert len(bench.pdifs_schedule(2000, True)[0]) == 1497     # C3: 2000 source frames, LIGHT -> 1497
    assert len(bench.pdifs_schedule(8000, True)[0]) == 5997     # C4
    assert bench.fps_field(2000, 1999, 24, True) == 18          # C3 header fps
    assert bench.fps_field(212, 212, 24, False) == 12           # C1 header fps
    for world in (1, 2, 4, 8):
        r = bench.shard_ranges(1497 * world + (3 if world > 1 else 0), world, 12)
        assert r[0][0] == 0 and all(a % 12 == 0 for a, _ in r) and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
