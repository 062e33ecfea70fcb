"""Run under torchrun on N GPUs: frame-range sharded encode over NCCL must be byte-identical to the single-process
encode of the same sequence (oracle and 1-GPU). Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import libagmv_b200  # noqa: E402
from agmv_testlib import LZSS, OPT, QUALITY, oracle_encode, synth_frames  # noqa: E402

W, H, N_PER_RANK = 320, 240, 64


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = libagmv_b200.Context(local, stream.cuda_stream)
    ok = True
    for optname, light in (("III", True), ("I", False)):
        n_total = N_PER_RANK * world
        frames = torch.from_numpy(synth_frames(W, H, N_PER_RANK, seed=5, first=1 + rank * N_PER_RANK).view(np.int32)).to(dev)
        ctx.enc_begin(W, H, OPT[optname], QUALITY["MID"], LZSS)
        ctx.enc_histogram(frames.data_ptr(), N_PER_RANK, True)
        p, nb = ctx.enc_histogram_ptr()
        ht = torch.as_tensor(bench._DevArray(p, nb, "<i8"), device=dev)
        dist.all_reduce(ht)
        torch.cuda.synchronize()
        ctx.enc_build_palette()
        sa_all, sb_all = bench.pdifs_schedule(n_total, light)
        e0, e1 = bench.shard_ranges(len(sa_all), world, 12 if light else 4)[rank]
        sa = sa_all[e0:e1] - rank * N_PER_RANK
        sb = np.where(sb_all[e0:e1] >= 0, sb_all[e0:e1] - rank * N_PER_RANK, -1).astype(np.int32)
        nbytes = ctx.enc_frames(frames.data_ptr(), N_PER_RANK, True, sa, sb, e0)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = nbytes
        dist.all_reduce(sizes)
        pad = int(sizes.max().item())
        ip, ib = ctx.enc_image_ptr()
        mine = torch.zeros(pad, dtype=torch.uint8, device=dev)
        mine[:ib] = torch.as_tensor(bench._DevArray(ip, ib, "|u1"), device=dev)
        gathered = [torch.zeros(pad, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            hdr = ctx.enc_header(n_total - 1, 24).tobytes()
            images = [gathered[r][: int(sizes[r].item())].cpu().numpy().tobytes() for r in range(world)]
            data = bench.assemble_container(hdr, images, len(sa_all), bench.fps_field(n_total, n_total - 1, 24, light))
            whole = synth_frames(W, H, n_total, seed=5)
            ref = oracle_encode(whole, n_total - 1, 24, OPT[optname], QUALITY["MID"], LZSS)
            single, _ = ctx.encode_sequence(whole, n_total - 1, 24, OPT[optname], QUALITY["MID"], LZSS)
            same = data == ref and single.tobytes() == ref
            print(f"OPT_{optname}: {world}-GPU sharded == oracle == 1-GPU: {same} ({len(data)} bytes)", flush=True)
            ok = ok and same
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
