#!/usr/bin/env python
"""Time the phases of one host-to-host sequence (upload+encode, open, decode+download) for both host pixel formats."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import libagmv_b200
W, H, N = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 512
P = W * H
dev = torch.device("cuda", 0)
ctx = libagmv_b200.Context(0)
frames = torch.empty((N, H, W), dtype=torch.int32, device=dev)
ctx.synth_frames(frames.data_ptr(), W, H, 1, N, 1234)
torch.cuda.synchronize()
for bpp in (4, 3):
    ctx.set_host_format(1 if bpp == 3 else 0)
    if bpp == 3:
        hf = torch.empty((N, H, W, 3), dtype=torch.uint8, pin_memory=True)
        hf.copy_(frames.view(torch.uint8).view(N, H, W, 4)[..., :3])
    else:
        hf = torch.empty((N, H, W), dtype=torch.int32, pin_memory=True)
        hf.copy_(frames)
    out = torch.empty(4096 + N * (P // 2), dtype=torch.uint8, pin_memory=True)
    n_enc = (N // 4) * 3
    dh = torch.empty((n_enc + 8, H, W, bpp) if bpp == 3 else (n_enc + 8, H, W), dtype=torch.uint8 if bpp == 3 else torch.int32, pin_memory=True)
    # raw copy rates
    t0 = time.perf_counter(); d = torch.empty_like(hf, device=dev); d.copy_(hf, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"bpp {bpp}: raw H2D {hf.numel() * hf.element_size() / 1e9 / (t1 - t0):.1f} GB/s (first), ", end="")
    t0 = time.perf_counter(); d.copy_(hf, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{hf.numel() * hf.element_size() / 1e9 / (t1 - t0):.1f} GB/s; ", end="")
    t0 = time.perf_counter(); hf.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"raw D2H {hf.numel() * hf.element_size() / 1e9 / (t1 - t0):.1f} GB/s")
    del d
    for it in range(3):
        a = hf.numpy() if bpp == 3 else hf.numpy().view(np.uint32)
        t0 = time.perf_counter()
        data, ne = ctx.encode_sequence(a, N - 1, 24, 3, 1, 1, out=out.numpy())
        t1 = time.perf_counter()
        sid, w, h, n = ctx.dec_open(data)
        t2 = time.perf_counter()
        ctx.dec_frames(sid, n, w, h, host_ptr=dh.data_ptr())
        t3 = time.perf_counter()
        ctx.dec_close(sid)
        t4 = time.perf_counter()
        print(f"  iter {it}: encode_sequence {1e3 * (t1 - t0):.1f} ms, dec_open {1e3 * (t2 - t1):.1f}, dec_frames {1e3 * (t3 - t2):.1f}, close {1e3 * (t4 - t3):.1f}; total {1e3 * (t4 - t0):.1f} ms = {N / (t4 - t0):.0f} fps")
ctx.set_host_format(0)
