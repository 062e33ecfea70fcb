#!/usr/bin/env python
"""bench.py - headline benchmark of the libagmv frame hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input: encode the
BASELINE.json config-3 sequence (1920x1080, AGMV_OPT_III, AGMV_HIGH_QUALITY, LZSS; 2000
source frames per GPU) and decode the stream it produced. The metric is the one
BASELINE.json names - encode & decode frames/sec at 1080p - reported as round-trip
source frames per second (the encode-only and decode-only rates are in `detail`).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C-ABI)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path on the host cores

N > 1 (one process per GPU under torchrun): weak scaling. The job is one sequence of
2000*N source frames sharded by GOP-aligned frame range; the colour histogram is
all-reduced over NCCL (the palette is global), every rank encodes its range, chunk sizes
are all-gathered and the payload is gathered on rank 0 for container assembly. Decode is
stream-parallel: each rank decodes its own independent stream. value = frames all ranks
processed / max-over-ranks device time.

Timing: W >= 3 untimed warm-up steps, then exactly K steps between barrier +
torch.cuda.synchronize() on both sides, timed with CUDA events on the launching stream
(the library launches everything on torch's current stream). Inputs (16.6 GB of frames) are
far larger than the 126 MB L2, so nothing needs flushing between iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 1920, 1080
OPT_III, HIGH, LZSS_C = 3, 1, 1
METRIC = "encode & decode frames/sec at 1080p"
UNIT = "source frames/s (encode+decode round trip)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing: report that instead of a number
            self.proc = None
            self.err = str(e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Samples taken before this call are dropped (nvidia-smi is started early so that its start-up cost stays
        outside the timed region)."""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t >= getattr(self, "t0", 0.0)]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# schedule helpers (host logic; tested on CPU in tests/test_sharding.py)
# --------------------------------------------------------------------------------------
def pdifs_schedule(n_src, light):
    """(src_a, src_b) per encoded frame for source frames 1..n_src; 0-based indices (src/agmv_encode.c:2727-2770, 3610-3612)."""
    sa, sb, i, end = [], [], 1, n_src
    while i <= end:
        if light:
            sa += [i - 1, i, i + 2]
            sb += [-1, i + 1, -1]
            i += 4
        else:
            sa.append(i - 1)
            sb.append(i)
            i += 2
        if i + 4 >= end:
            break
    return np.array(sa, np.int32), np.array(sb, np.int32)


def shard_ranges(n_enc, world, gop_group):
    """Contiguous encoded-frame ranges whose boundaries are multiples of `gop_group` encoded frames
    (12 for LIGHT profiles = 3 GOPs = 4 PDIFS groups, 4 for HEAVY)."""
    units = (n_enc + gop_group - 1) // gop_group
    per = (units + world - 1) // world
    out = []
    for r in range(world):
        a = min(r * per * gop_group, n_enc)
        b = min((r + 1) * per * gop_group, n_enc)
        out.append((a, b))
    return out


def assemble_container(header, images, n_enc, fps_field):
    """Container assembly on rank 0: header | chunk images in rank order; back-patch frame count and fps."""
    out = bytearray(header)
    for img in images:
        out += img
    out[4:8] = int(n_enc).to_bytes(4, "little")
    out[18:22] = int(fps_field).to_bytes(4, "little")
    return bytes(out)


def fps_field(n_src, create_n, fps, light):
    adjusted = n_src - 1
    adjusted = int(adjusted * 0.75) if light else adjusted // 2
    rate = np.float32(adjusted) / np.float32(create_n + 1)
    v = float(np.float32(fps) * rate)
    return int(np.floor(v + 0.5))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
_ALL_CPUS = None


def bind_to_gpu_numa_node(torch, local):
    """Run this rank's host threads (and so its pinned allocations) on the CPU cores next to its GPU: with several GPUs on
    one node the e2e leg is host-memory bound, and buffers on the far socket cost a second hop per transfer."""
    global _ALL_CPUS
    try:
        import pynvml
        _ALL_CPUS = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1} & _ALL_CPUS
        if cpus:
            os.sched_setaffinity(0, cpus)
            log(f"rank on GPU {local}: bound to {len(cpus)} of {len(_ALL_CPUS)} cores")
    except Exception as e:  # affinity is an optimisation only
        log(f"no NUMA binding: {e}")


def run_b200(args):
    import torch
    import torch.distributed as dist
    import libagmv_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # a real (non-default) stream: the library launches on it and torch's events time it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = libagmv_b200.Context(local, stream.cuda_stream)

    n_local = args.frames                      # source frames per GPU
    COMP = 2 if args.compression == "lz77" else LZSS_C
    if COMP == 2 and world > 1:
        # AGMV_LZ77 carries its bitstream buffer from frame to frame (src/agmv_encode.c:218-224): frame-range shards of one
        # sequence would each start from an empty buffer and the assembled file would not be the single-process stream
        raise SystemExit("--compression lz77 is not sharded across GPUs (DESIGN.md section 8, N2): run it with --gpus 1")
    n_total = n_local * world
    P = W * H
    # inputs resident in HBM before the timed region
    frames = torch.empty((n_local, H, W), dtype=torch.int32, device=dev)
    ctx.synth_frames(frames.data_ptr(), W, H, 1 + rank * n_local, n_local, 1234)
    torch.cuda.synchronize()

    light = True
    sa_all, sb_all = pdifs_schedule(n_total, light)
    n_enc_total = len(sa_all)
    ranges = shard_ranges(n_enc_total, world, 12)
    e0, e1 = ranges[rank]
    sa = sa_all[e0:e1] - rank * n_local
    sb = np.where(sb_all[e0:e1] >= 0, sb_all[e0:e1] - rank * n_local, -1).astype(np.int32)
    assert sa.min() >= 0 and max(sa.max(), sb.max()) < n_local, "shard needs frames outside its range"
    n_enc = e1 - e0
    create_n, fps = n_total - 1, 24

    dec_out = torch.empty((n_enc, H, W), dtype=torch.int32, device=dev)
    host_stream = {}

    def encode_step(fetch=False):
        ctx.enc_begin(W, H, OPT_III, HIGH, COMP)
        ctx.enc_histogram(frames.data_ptr(), n_local, True)
        if world > 1:
            p, nb = ctx.enc_histogram_ptr()
            # wrap the library's device histogram (u64 bins) as a tensor without copying
            ht = torch.as_tensor(_DevArray(p, nb, "<i8"), device=dev)
            dist.all_reduce(ht)
        ctx.enc_build_palette()
        nbytes = ctx.enc_frames(frames.data_ptr(), n_local, True, sa, sb, e0)
        if world > 1:
            sizes = torch.zeros(world, dtype=torch.int64, device=dev)
            sizes[rank] = nbytes
            dist.all_reduce(sizes)  # chunk-image sizes of every rank -> file offsets
            host_stream["sizes"] = sizes
        return nbytes

    def decode_step(stream_bytes):
        sid, w, h, n = ctx.dec_open(stream_bytes)
        ctx.dec_frames(sid, n, w, h, device_ptr=dec_out.data_ptr())
        ctx.dec_close(sid)
        return n

    def local_stream(nbytes):
        img, _, _ = ctx.enc_fetch(nbytes, n_enc)
        hdr = ctx.enc_header(create_n, fps)
        return assemble_container(hdr.tobytes(), [img.tobytes()], n_enc, fps_field(n_total, create_n, fps, light))

    # one untimed pass builds the stream the decode half works on (and warms every workspace)
    nbytes = encode_step()
    stream_bytes = local_stream(nbytes)
    stream_pin = torch.empty(len(stream_bytes), dtype=torch.uint8, pin_memory=True)   # the compressed stream lives in pinned host memory
    stream_pin.copy_(torch.frombuffer(bytearray(stream_bytes), dtype=torch.uint8))
    stream_np = stream_pin.numpy()
    n_dec = decode_step(stream_np)
    assert n_dec == n_enc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def full_step():
        encode_step()
        decode_step(stream_np)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        full_step()
    sampler.mark()
    l0 = ctx.launches
    total_ms = timed(full_step, args.steps)          # the headline: K steps, no profiling events in the stream
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    # same K steps again with a CUDA event pair around every launch: per-kernel-class device time, live
    ctx.profile(True)
    prof_ms = timed(full_step, args.steps)
    prof = ctx.profile_read()
    ctx.profile(False)

    # encode-only / decode-only rates (same rules, separate regions)
    enc_ms = timed(encode_step, args.steps)
    dec_ms = timed(lambda: decode_step(stream_np), args.steps)
    us, cs = ctx.enc_sizes(n_enc)                    # per-frame sizes of the timed workload (read before the e2e leg re-encodes a shorter clip)
    dec_out.untyped_storage().resize_(0)             # the decoded frames of the timed workload are not needed any more (12.4 GB)
    torch.cuda.empty_cache()

    # ---- e2e: the same step through the public API with HOST buffers -------------------
    # Every sequence goes host -> encode -> host stream -> decode -> host frames through agmvb_encode_sequence /
    # agmvb_dec_open / agmvb_dec_frames. A sequence cannot overlap its own transfers with its own compute (the palette
    # needs every frame before the first one can be encoded), so throughput comes from running `--e2e-streams`
    # independent sequences at a time, one host thread + context (own CUDA stream) each: one sequence's upload
    # overlaps another's kernels and download, as a service encoding several clips would run it.
    e2e = None
    if not args.no_e2e:
        import threading
        e2e_frames = min(n_local, args.e2e_frames)
        n_streams = max(1, args.e2e_streams)
        # pinned host memory: frames in + frames out + stream per sequence in flight; stay within a quarter of this rank's share
        try:
            import psutil
            budget = psutil.virtual_memory().available / max(1, world) / 4
        except Exception:
            budget = 16e9
        while n_streams * e2e_frames * P * 4 * 1.9 > budget and (n_streams > 1 or e2e_frames > 64):
            if n_streams > 1:
                n_streams -= 1
            else:
                e2e_frames //= 2
        sa_e, sb_e = pdifs_schedule(e2e_frames, light)
        iters = args.e2e_iters if args.e2e_iters > 0 else max(16, args.steps)   # the pipeline of sequences needs a few rounds to fill and drain; run to run the leg varies by ~5 %
        bpp = 3 if args.e2e_pixels == "bgr24" else 4

        class Worker:
            def __init__(self, k, e2e_frames=e2e_frames):
                self.nf = e2e_frames
                n_dec = len(pdifs_schedule(e2e_frames, light)[0])
                self.ctx = ctx if k == 0 else libagmv_b200.Context(local)
                # host pixels in the reference's own frame format: the packed B,G,R rows of 24-bit BMP files (AGMVB_PIX_BGR24)
                self.ctx.set_host_format(1 if bpp == 3 else 0)
                self.host_frames = torch.empty((e2e_frames, H, W, bpp) if bpp == 3 else (e2e_frames, H, W),
                                               dtype=torch.uint8 if bpp == 3 else torch.int32, pin_memory=True)
                if bpp == 3:
                    src = frames[:e2e_frames].view(torch.uint8).view(e2e_frames, H, W, 4)[..., :3]
                    self.host_frames.copy_(src)
                else:
                    self.host_frames.copy_(frames[:e2e_frames])
                self.out_host = torch.empty(4096 + e2e_frames * (P // 2), dtype=torch.uint8, pin_memory=True)
                self.dec_host = torch.empty((n_dec, H, W, bpp) if bpp == 3 else (n_dec, H, W),
                                            dtype=torch.uint8 if bpp == 3 else torch.int32, pin_memory=True)
                self.h2d = self.d2h = 0
                self.err = None

            def step(self):
                c = self.ctx
                e2e_frames = self.nf
                hf = self.host_frames.numpy()
                data, ne = c.encode_sequence(hf if bpp == 3 else hf.view(np.uint32), e2e_frames - 1, 24, OPT_III, HIGH, COMP,
                                             out=self.out_host.numpy())
                sid, w, h, n = c.dec_open(data)
                c.dec_frames(sid, n, w, h, host_ptr=self.dec_host.data_ptr())
                c.dec_close(sid)
                self.h2d = e2e_frames * P * bpp + len(data)
                self.d2h = len(data) + n * P * bpp

            def run(self, k):
                try:
                    torch.cuda.set_device(local)
                    for _ in range(k):
                        self.step()
                except Exception as e:  # surfaced after join
                    self.err = e

        # build and warm the workers; if a rank runs out of (device or pinned) memory every rank halves the number of sequences
        # in flight together and tries again
        workers = None
        while True:
            ok = 1
            try:
                workers = [Worker(k) for k in range(n_streams)]
                for wk in workers:
                    wk.step()                               # warm every context's workspaces
            except Exception as e:
                log(f"e2e with {n_streams} sequences in flight failed on rank {rank}: {e}")
                ok = 0
            if world > 1:
                t = torch.tensor([ok], dtype=torch.int32, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ok = int(t.item())
            if ok:
                break
            for wk in workers or []:
                if wk.ctx is not ctx:
                    wk.ctx.close()
            workers = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            if n_streams == 1:
                break
            n_streams = max(1, n_streams // 2)
        if workers:
            barrier()
            t0 = time.perf_counter()
            threads = [threading.Thread(target=wk.run, args=(iters,)) for wk in workers]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            barrier()
            dt = time.perf_counter() - t0
            for wk in workers:
                if wk.err is not None:
                    raise wk.err
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            seqs = n_streams * iters
            e2e = {"value": e2e_frames * world * seqs / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(wk.h2d for wk in workers)),
                   "d2h_bytes_per_step": int(sum(wk.d2h for wk in workers)), "frames_per_step": e2e_frames * n_streams,
                   "concurrent_sequences": n_streams, "host_pixels": args.e2e_pixels,
                   "note": f"{n_streams} independent {e2e_frames}-frame sequences at a time per GPU (one host thread + context each): "
                           "agmvb_encode_sequence + agmvb_dec_open/agmvb_dec_frames on pinned host buffers "
                           + ("(packed 24-bit BMP pixel rows in and out, AGMVB_PIX_BGR24)" if bpp == 3 else "(u32 pixels)") + ", host wall clock"}
        # second figure: ONE sequence of the whole configuration (all of this GPU's source frames) host to host, nothing else in
        # flight - upload, encode, fetch, open, decode and download follow each other, as a single AGMV_EncodeAGMV +
        # AGMV_DecodeAGMV caller sees them
        single_need = n_local * P * bpp * 1.75 + n_local * (P // 2) + 4096      # pinned bytes: frames in, decoded frames out, stream
        try:
            import psutil
            single_ok = single_need < psutil.virtual_memory().total / max(1, world) / 2.5
        except Exception:
            single_ok = world == 1
        if e2e is not None and not args.no_e2e_single and not single_ok:
            e2e["single_sequence"] = {"skipped": f"{single_need / 1e9:.1f} GB of pinned host memory per rank for one {n_local}-frame sequence does not fit {world} ranks on this host"}
        if e2e is not None and not args.no_e2e_single and single_ok:
            workers = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            try:
                wk = Worker(0, n_local)
                wk.step()
                barrier()
                t0 = time.perf_counter()
                for _ in range(2):
                    wk.step()
                barrier()
                dt1 = (time.perf_counter() - t0) / 2
                if world > 1:
                    t = torch.tensor([dt1], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt1 = float(t.item())
                e2e["single_sequence"] = {"value": n_local * world / dt1, "unit": UNIT, "frames": n_local,
                                          "note": "one sequence per GPU, nothing overlapped: upload + encode + fetch + open + decode + download in turn"}
                del wk
            except Exception as e:
                log(f"single-sequence e2e leg failed on rank {rank}: {e}")
        ctx.set_host_format(0)
        workers = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # per-step statistics the algorithmic byte counts are made of
    usize_total, csize_total = float(us.sum()), float(cs.sum())
    n_interp = int((sb >= 0).sum())
    stats = {"P": P, "n_src": n_local, "n_enc": n_enc, "n_interp": n_interp, "usize": usize_total, "csize": csize_total,
             "lz_levels": 12, "lz_groups": max(1, prof.get("lz_link", (1, 0))[0] // max(1, args.steps))}
    # dominant kernel class by device time inside the timed region
    roof = None
    if prof:
        name, (cnt, ms) = max(prof.items(), key=lambda kv: kv[1][1])
        alg_step = ALG_BYTES[name](stats) if name in ALG_BYTES else None
        per_launch_ms = ms / cnt
        roof = {"kernel": name, "bound": "hbm", "launches_in_region": cnt, "avg_launch_ms": per_launch_ms, "peak": peak, "unit": "GB/s",
                "peak_source": peak_src, "traffic": None}
        if alg_step is not None:
            bytes_per_launch = alg_step * args.steps / cnt
            roof["achieved"] = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["algorithmic_bytes_per_launch"] = bytes_per_launch
            roof["algorithmic_bytes_rule"] = ALG_RULE.get(name)
            ratio = TRAFFIC_NCU.get("traffic_over_algorithmic", {}).get(name)
            if ratio is not None:  # measured DRAM bytes per algorithmic byte (ncu --set full, profiles/), scaled to this run's launch size
                roof["traffic"] = ratio * bytes_per_launch
                roof["traffic_source"] = "profiles/traffic.json: dram__bytes_read+write per launch / algorithmic bytes per launch at capture, x this run's algorithmic bytes per launch"
        # the whole path against its compulsory traffic: encode 8*W*H per source frame + chunks out,
        # decode chunks in + 4*W*H out per frame
        enc_bytes = 8.0 * P * n_local + 24.0 * n_enc + csize_total
        dec_bytes = 24.0 * n_enc + csize_total + 4.0 * P * n_enc
        roof["path"] = {"encode_GBps": enc_bytes * args.steps / (enc_ms * 1e-3) / 1e9, "encode_frac": enc_bytes * args.steps / (enc_ms * 1e-3) / 1e9 / peak,
                        "decode_GBps": dec_bytes * args.steps / (dec_ms * 1e-3) / 1e9, "decode_frac": dec_bytes * args.steps / (dec_ms * 1e-3) / 1e9 / peak}

    value = n_total * args.steps / (total_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"BASELINE config 3: {n_local} source frames/GPU 1920x1080 24fps, AGMV_OPT_III, AGMV_HIGH_QUALITY, {args.compression.upper()}; "
                               f"encode then decode of the {n_enc}-frame stream", "source_frames_per_gpu": n_local,
                   "encoded_frames_per_gpu": n_enc, "stream_bytes_per_gpu": len(stream_bytes),
                   "l2_policy": "inputs (16.6 GB/GPU) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"frame-range shard x{world}" if world > 1 else "single GPU"},
        "detail": {"encode_source_fps": n_total * args.steps / (enc_ms * 1e-3), "encode_encoded_fps": n_enc * world * args.steps / (enc_ms * 1e-3),
                   "decode_fps": n_enc * world * args.steps / (dec_ms * 1e-3),
                   "kernel_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                   "profiled_ms_per_step": prof_ms / args.steps,
                   "alg_bytes_per_launch": {k: ALG_BYTES[k](stats) * args.steps / v[0] for k, v in prof.items() if k in ALG_BYTES}},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, bounded_seconds=25)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# BASELINE configs 4 and 5 (--config c4 | c5): the sharded 4K encode and the batched multi-stream decode
# --------------------------------------------------------------------------------------
def _dist_setup(args):
    import torch
    import torch.distributed as dist
    import libagmv_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = libagmv_b200.Context(local, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    return torch, dist, rank, world, local, dev, stream, ctx, barrier, timed


def run_c4(args):
    """BASELINE config 4: ONE 3840x2160 sequence of 8000 source frames (OPT_III / HIGH / LZSS), sharded by GOP-aligned frame
    range over the ranks. The frames cannot sit in HBM (265 GB), so every rank generates its source frames in-kernel, chunk by
    chunk, in both passes (histogram, then encode) - the generator is part of the timed step. Step = histogram of the rank's
    share of the source frames, NCCL all-reduce of the bins, palette, encode of the rank's range, size exchange, gather of the
    chunk images on rank 0 and container assembly in host memory. The assembled file is checked against a single-GPU encode."""
    import hashlib
    torch, dist, rank, world, local, dev, stream, ctx, barrier, timed = _dist_setup(args)
    W4, H4 = 3840, 2160
    P = W4 * H4
    n_total = args.frames if args.frames != 2000 else 8000
    light = True
    sa_all, sb_all = pdifs_schedule(n_total, light)
    n_enc_total = len(sa_all)
    create_n, fps = n_total - 1, 24
    CH = 192                                            # source frames generated at a time (6.4 GB)
    buf = torch.empty((CH + 16, H4, W4), dtype=torch.int32, device=dev)
    img_cap = int(n_enc_total * P * 0.15) + (64 << 20)          # chunk images of the whole sequence (rank 0 also encodes it alone, as the check)
    img_dev = torch.empty(img_cap, dtype=torch.uint8, device=dev)
    state = {}

    def encode_range(e0, e1, h0, h1, world_for_reduce):
        ctx.enc_begin(W4, H4, OPT_III, HIGH, LZSS_C)
        for c in range(h0, h1, CH):                      # pass 1: every source frame of the share, once
            n = min(CH, h1 - c)
            ctx.synth_frames(buf.data_ptr(), W4, H4, 1 + c, n, 1234)
            ctx.enc_histogram(buf.data_ptr(), n, True)
        if world_for_reduce > 1:
            p, nb = ctx.enc_histogram_ptr()
            dist.all_reduce(torch.as_tensor(_DevArray(p, nb, "<i8"), device=dev))
        ctx.enc_build_palette()
        off = 0
        for a in range(e0, e1, CH // 16 * 12):           # pass 2: whole groups of 12 encoded = 16 source frames
            b = min(e1, a + CH // 16 * 12)
            lo = int(sa_all[a])
            hi = int(max(sa_all[b - 1], sb_all[a:b].max())) + 1
            ctx.synth_frames(buf.data_ptr(), W4, H4, 1 + lo, hi - lo, 1234)
            nbytes = ctx.enc_frames(buf.data_ptr(), hi - lo, True, sa_all[a:b] - lo, np.where(sb_all[a:b] >= 0, sb_all[a:b] - lo, -1), a)
            ptr, _ = ctx.enc_image_ptr()
            assert off + nbytes <= img_cap
            img_dev[off:off + nbytes].copy_(torch.as_tensor(_DevArray(ptr, nbytes, "|u1"), device=dev), non_blocking=True)
            off += nbytes
        return off

    host_out = torch.empty(2048 + img_cap, dtype=torch.uint8, pin_memory=True) if rank == 0 else None

    def step():
        e0, e1 = shard_ranges(n_enc_total, world, 12)[rank]
        h0, h1 = n_total * rank // world, n_total * (rank + 1) // world
        nbytes = encode_range(e0, e1, h0, h1, world)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = nbytes
        if world > 1:
            dist.all_reduce(sizes)                       # chunk-image sizes of every rank -> file offsets
        sz = sizes.cpu().tolist()
        total = int(sum(sz))
        if rank == 0:
            hdr = ctx.enc_header(create_n, fps)
            hl = len(hdr)
            host_out[:hl].copy_(torch.from_numpy(hdr))
            gathered = torch.empty(total, dtype=torch.uint8, device=dev) if world > 1 else img_dev
            if world > 1:
                gathered[:sz[0]].copy_(img_dev[:sz[0]])
                o = sz[0]
                for r in range(1, world):
                    dist.recv(gathered[o:o + sz[r]], src=r)
                    o += sz[r]
            host_out[hl:hl + total].copy_(gathered[:total])           # one D2H of the file body
            torch.cuda.current_stream().synchronize()
            out = host_out.numpy()
            out[4:8] = np.frombuffer(int(n_enc_total).to_bytes(4, "little"), np.uint8)
            out[18:22] = np.frombuffer(int(fps_field(n_total, create_n, fps, light)).to_bytes(4, "little"), np.uint8)
            state["len"] = hl + total
        else:
            dist.send(img_dev[:nbytes], dst=0)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(1, args.warmup)):
        step()
    sampler.mark()
    l0 = ctx.launches
    total_ms = timed(step, args.steps)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ok = None
    if rank == 0:
        digest = hashlib.sha256(host_out.numpy()[:state["len"]].tobytes()).hexdigest()
        if world > 1:                                    # the same sequence on this GPU alone: the file must not depend on the GPU count
            n1 = encode_range(0, n_enc_total, 0, n_total, 1)
            hdr = ctx.enc_header(create_n, fps)
            one = bytearray(hdr.tobytes()) + bytearray(img_dev[:n1].cpu().numpy().tobytes())
            one[4:8] = int(n_enc_total).to_bytes(4, "little")
            one[18:22] = int(fps_field(n_total, create_n, fps, light)).to_bytes(4, "little")
            ok = hashlib.sha256(bytes(one)).hexdigest() == digest
            assert ok, "the sharded stream differs from the single-GPU stream"
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak = 6554.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = (8.0 * P * n_total + state["len"]) * args.steps / (total_ms * 1e-3) / 1e9
    line = {"metric": "encode frames/sec at 3840x2160 (frame-range sharded sequence)", "value": n_total * args.steps / (total_ms * 1e-3),
            "unit": "source frames/s (encode, container assembled on rank 0)", "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE config 4: one {n_total}-frame 3840x2160 sequence, AGMV_OPT_III, AGMV_HIGH_QUALITY, LZSS, "
                                   f"frames generated in-kernel in both passes, {n_enc_total} encoded frames", "parallelism": f"frame-range shard x{world}",
                       "stream_bytes": state["len"], "stream_sha256": digest, "equals_single_gpu_stream": ok,
                       "l2_policy": "every pass streams 6.4 GB chunks of generated frames: far beyond the 126 MB L2"},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": None,
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "achieved": gbs, "frac": gbs / peak / world,
                         "note": "whole encode path against its compulsory bytes (8*W*H per source frame + the stream), per GPU"}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE config 5: 512 independent 1080p streams (stream s = a config-3-style encode of 64 source frames, seed 1000+s,
    45 encoded frames), decoded stream-parallel: every rank owns 512 / N streams and advances all of them together
    (agmvb_dec_batch: one launch per pipeline stage for every stream). Step = open (header parse, chunk index, upload from
    pinned host memory) + decode of every frame into the per-stream frame ring + per-frame checksums back on the host."""
    torch, dist, rank, world, local, dev, stream, ctx, barrier, timed = _dist_setup(args)
    S = 512
    mine = list(range(rank, S, world))
    src = torch.empty((64, H, W), dtype=torch.int32, device=dev)
    streams = []
    for s in mine:                                       # untimed: make this rank's streams
        ctx.synth_frames(src.data_ptr(), W, H, 1, 64, 1000 + s)
        data, ne = ctx.encode_sequence(None, 63, 24, OPT_III, HIGH, LZSS_C, device_ptr=src.data_ptr(), shape=(64, H, W))
        pin = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
        pin.copy_(torch.from_numpy(np.array(data, copy=True)))
        streams.append(pin.numpy())
    n_fr = int(ne)
    del src
    torch.cuda.empty_cache()
    state = {}

    def step():
        sids = [ctx.dec_open(d)[0] for d in streams]
        state["ck"] = ctx.dec_batch(sids, n_fr, None, checksums=True)
        for sid in sids:
            ctx.dec_close(sid)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(1, args.warmup)):
        step()
    sampler.mark()
    l0 = ctx.launches
    total_ms = timed(step, args.steps)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    # checks: the batched decoder against the single-stream decoder on a sample, and one digest over every frame of every stream
    # (a sum, so it does not depend on how the streams are dealt to the ranks)
    for k in range(0, len(mine), max(1, len(mine) // 4)):
        single = ctx.decode_all(streams[k])
        wts = np.uint64(2654435761) + np.uint64(2) * np.arange(W * H, dtype=np.uint64)
        exp = np.array([int((f.reshape(-1).astype(np.uint64) * wts).sum(dtype=np.uint64)) for f in single], dtype=np.uint64)
        assert np.array_equal(exp, state["ck"][k]), f"batched decode of stream {mine[k]} differs from the single-stream decoder"
    dig = torch.tensor([int(state["ck"].sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF))], dtype=torch.int64, device=dev)
    nbytes = torch.tensor([sum(len(d) for d in streams)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(dig)
        dist.all_reduce(nbytes)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak = 6554.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    frames = S * n_fr
    gbs = (4.0 * W * H * frames + float(nbytes.item())) * args.steps / (total_ms * 1e-3) / 1e9
    v = frames * args.steps / (total_ms * 1e-3)
    line = {"metric": "decode frames/sec at 1080p (512 independent streams, stream-parallel)", "value": v, "unit": "frames/s (open + upload + decode)",
            "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE config 5: {S} streams x {n_fr} frames 1920x1080 (each a 64-source-frame OPT_III / HIGH / LZSS encode, "
                                   f"seeds 1000+s), {S // world} streams per GPU advanced together", "parallelism": f"stream-parallel x{world}",
                       "compressed_bytes": int(nbytes.item()), "checksum_digest": int(dig.item()) & 0x7FFFFFFFFFFFFFFF,
                       "l2_policy": "each step decodes 23040 frames (191 GB of pixels) through an 8-frame ring per stream: far beyond L2"},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": int(nbytes.item()), "d2h_bytes_per_step": frames * 8,
                    "note": "the timed step already starts from compressed streams in pinned host memory and ends with per-frame checksums on the host"},
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "achieved": gbs, "frac": gbs / peak / world,
                         "note": "whole decode path against its compulsory bytes (stream in + 4*W*H out per frame), per GPU"}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


# Algorithmic (compulsory) bytes PER STEP for each kernel class as a function of the step's statistics
# (SURVEY.md 8d per-kernel figures). K4 (LZSS) is usize in + csize out for the whole stage;
# its 15 refinement passes each have to touch every bitstream byte once, so each pass launch is charged usize.
ALG_BYTES = {
    "hist": lambda s: 4.0 * s["P"] * s["n_src"],
    "quantize": lambda s: 4.0 * s["P"] * (s["n_enc"] + s["n_interp"]) + 2.0 * s["P"] * s["n_enc"],
    "classify": lambda s: 2.0 * s["P"] * s["n_enc"] * 1.75 + s["n_enc"] * s["P"] / 16,
    "emit": lambda s: 2.0 * s["P"] * s["n_enc"] + s["usize"],
    "lz_link": lambda s: s["usize"],
    "lz_link3": lambda s: s["usize"],
    "lz_level": lambda s: s["usize"] * 12,
    "lz_pack": lambda s: s["usize"] + s["csize"],
    "lz77": lambda s: s["usize"] + s["csize"],
    "expand": lambda s: s["csize"] + s["usize"],
    "index": lambda s: s["usize"],
    "reconstruct": lambda s: s["usize"] + 4.0 * s["P"] * s["n_enc"],
}
ALG_RULE = {
    "lz_level": "K4 reads usize and writes csize once for the whole stage; each of its 12 chain-level passes has to see every bitstream byte "
                "once, so a level launch is charged one pass over the bitstream (usize bytes); it actually streams 9 B per position "
                "(4-byte link word in and out, one key byte) and gathers 4 B per chain hop",
    "lz_link": "K4 hash links: one pass over the bitstream (usize bytes) per launch",
    "lz_link3": "K4 level-3 links: one pass over the bitstream (usize bytes) per launch",
    "expand": "D2: csize in + usize out",
    "lz77": "K4 (LZ77 flavour): usize in + csize out",
    "reconstruct": "D3: usize in + 4*W*H out per frame",
    "quantize": "K1+K2: 4 B per source pixel read (8 when interpolating) + 2 B entry written",
    "hist": "K0a: 4 B per source pixel",
}
# dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full captures (profiles/), bench config
TRAFFIC_NCU = {}
try:
    TRAFFIC_NCU = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass


# --------------------------------------------------------------------------------------
# CPU legs: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------
def cpu_clips(n_src, procs, use_ref, quality="HIGH", size=(1920, 1080)):
    """Each of `procs` processes encodes its own n_src-frame clip (OPT_III / `quality` / LZSS) and decodes it without export.
    The workers time the encode and the decode themselves (tests/cpu_worker.py); the processes run side by side, so the
    aggregate rate is frames / the slowest worker's time. Returns a dict, or None if a worker failed."""
    worker = os.path.join(ROOT, "tests", "cpu_worker.py")
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, worker, str(n_src), str(1000 + k), "ref" if use_ref else "port", str(size[0]), str(size[1]), quality],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for k in range(procs)]
    outs = [p.communicate()[0] for p in ps]
    wall = time.perf_counter() - t0
    if not all(p.returncode == 0 for p in ps):
        return None
    rows = []
    for o in outs:
        t = o.split()
        rows.append({t[i]: float(t[i + 1]) for i in range(0, len(t) - 1, 2)})
    enc = max(r["encode_s"] for r in rows)
    dec = max(r["decode_s"] for r in rows)
    both = max(r["encode_s"] + r["decode_s"] for r in rows)
    n_enc = sum(r["frames"] for r in rows)
    return {"procs": procs, "source_frames": n_src * procs, "encoded_frames": int(n_enc), "encode_s": enc, "decode_s": dec, "roundtrip_s": both,
            "palette_s": max(r["palette_s"] for r in rows), "wall_s": wall,
            "roundtrip_source_fps": n_src * procs / max(both, 1e-6), "encode_source_fps": n_src * procs / max(enc, 1e-6), "decode_fps": n_enc / max(dec, 1e-6)}


def reference_unmodified(procs):
    """The UNMODIFIED reference (oracle/_ref, gcc -O2 from /root/reference) on a prefix it can finish: 8 source frames of the
    config-3 profile at LOW quality (HIGH needs ~357 s of AGMV_BubbleSort before the first frame), one
    process on one core and `procs` processes side by side; the sort's fixed cost is measured on a 64x64 clip, where nothing
    else takes time. Decode = the per-frame public API without BMP export."""
    from agmv_testlib import have_ref
    if not have_ref():
        return {"unavailable": "oracle/_ref not built (needs the reference sources at build time)"}
    out = {"clip": "8 source frames 1920x1080, AGMV_OPT_III, AGMV_LOW_QUALITY, LZSS (3 encoded frames)"}
    fixed = cpu_clips(8, 1, True, "LOW", (64, 64))
    one = cpu_clips(8, 1, True, "LOW")
    if fixed:
        out["fixed_palette_s_low_quality"] = fixed["encode_s"]
    if one:
        out["one_core"] = {k: one[k] for k in ("encode_s", "decode_s", "roundtrip_source_fps", "encode_source_fps", "decode_fps", "encoded_frames")}
        if fixed:
            out["one_core"]["encode_s_per_encoded_frame_after_fixed"] = (one["encode_s"] - fixed["encode_s"]) / max(1, one["encoded_frames"])
    if procs > 1:
        many = cpu_clips(8, procs, True, "LOW")
        if many:
            out["all_cores"] = {k: many[k] for k in ("procs", "encode_s", "decode_s", "roundtrip_source_fps", "encode_source_fps", "decode_fps")}
    return out


def cpu_baseline(args, bounded_seconds=25, with_reference=True):
    if _ALL_CPUS:
        os.sched_setaffinity(0, _ALL_CPUS)   # the CPU legs use every host core, not just the ones next to GPU 0
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.cpu_procs or cores))
    n_src = 8
    r = cpu_clips(n_src, procs, use_ref=False)
    one = cpu_clips(n_src, 1, use_ref=False)
    line = {"value": r["roundtrip_source_fps"] if r else None, "unit": UNIT, "cores": procs, "kind": "port",
            "per_core": (r["roundtrip_source_fps"] / procs) if r else None,
            "encode_source_fps": r["encode_source_fps"] if r else None, "decode_fps": r["decode_fps"] if r else None,
            "palette_s": r["palette_s"] if r else None,
            "one_core": {k: one[k] for k in ("roundtrip_source_fps", "encode_source_fps", "decode_fps", "encode_s", "decode_s", "palette_s")} if one else None,
            "sample": f"{procs} processes x one {n_src}-source-frame 1080p clip (OPT_III, HIGH, LZSS), oracle port encode + no-export decode, "
                      f"timed inside the workers (slowest worker: {r['roundtrip_s']:.1f} s)" if r else "worker failed",
            "note": "oracle/agmv_oracle.c (qsort instead of the reference's O(n^2) bubble sort, memoised quantiser): faster than the "
                    "unmodified reference, so this flatters the CPU; the unmodified reference is timed beside it on a LOW-quality prefix"}
    if with_reference:
        line["reference_unmodified"] = reference_unmodified(procs)
    return line


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.cpu_procs or cores))
    n_src = 8
    vals, dts, last = [], [], None
    for _ in range(args.warmup + args.steps):
        r = cpu_clips(n_src, procs, use_ref=False)
        vals.append(r["roundtrip_source_fps"])
        dts.append(r["roundtrip_s"])
        last = r
    vals, dts = vals[args.warmup:], dts[args.warmup:]
    v = float(np.mean(vals))
    sample = (f"each step: {procs} processes x one {n_src}-source-frame 1080p clip (OPT_III, HIGH, LZSS), encode + no-export decode with the "
              f"oracle port of the reference algorithm, timed inside the workers")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(dts)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "BASELINE config 3 profile (1920x1080, AGMV_OPT_III, AGMV_HIGH_QUALITY, LZSS), bounded sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                             "encode_source_fps": last["encode_source_fps"], "decode_fps": last["decode_fps"], "palette_s": last["palette_s"],
                             "reference_unmodified": reference_unmodified(procs)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 (default, the headline): 1080p encode + decode; c4: one 8000-frame 4K sequence sharded by frame range; c5: 512 streams decoded stream-parallel")
    ap.add_argument("--frames", type=int, default=2000, help="source frames per GPU (BASELINE config 3: 2000)")
    ap.add_argument("--compression", default="lzss", choices=["lzss", "lz77"], help="entropy coder (BASELINE config 3: lzss; lz77 = SURVEY 8f N2)")
    ap.add_argument("--e2e-frames", type=int, default=512)
    ap.add_argument("--e2e-streams", type=int, default=6, help="independent sequences in flight per GPU in the e2e leg")
    ap.add_argument("--e2e-pixels", default="bgr24", choices=["bgr24", "u32"], help="host pixel format of the e2e leg")
    ap.add_argument("--e2e-iters", type=int, default=0, help="sequences per worker in the e2e leg (default max(16, steps))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-single", action="store_true", help="skip the one-sequence-of-the-whole-configuration figure of the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-procs", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_c4(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
