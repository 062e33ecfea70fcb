#!/usr/bin/env python
"""How much of the e2e gap is PCIe/host and how much is the GPU itself? K contexts run 512-frame sequences concurrently with
device-resident frames in and out (no bulk PCIe traffic), and - for comparison - with host buffers."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import libagmv_b200
W, H, N = 1920, 1080, int(os.environ.get("PROBE_FRAMES", "512"))
P = W * H
dev = torch.device("cuda", 0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = 4
class Wk:
    def __init__(self):
        self.ctx = libagmv_b200.Context(0)
        self.frames = torch.empty((N, H, W), dtype=torch.int32, device=dev)
        self.ctx.synth_frames(self.frames.data_ptr(), W, H, 1, N, 1234)
        self.out = torch.empty(4096 + N * (P // 2), dtype=torch.uint8, pin_memory=True)
        self.dec = torch.empty((N, H, W), dtype=torch.int32, device=dev)
    def step(self):
        data, ne = self.ctx.encode_sequence(None, N - 1, 24, 3, 1, 1, device_ptr=self.frames.data_ptr(), shape=(N, H, W), out=self.out.numpy())
        sid, w, h, n = self.ctx.dec_open(data)
        self.ctx.dec_frames(sid, n, w, h, device_ptr=self.dec.data_ptr())
        self.ctx.dec_close(sid)
    def run(self):
        for _ in range(iters):
            self.step()
for k in (1, K):
    ws = [Wk() for _ in range(k)]
    for w in ws:
        w.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=w.run) for w in ws]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{k} contexts, device-resident frames: {k * iters * N / dt:.0f} source fps ({1e3 * dt / (k * iters):.1f} ms per sequence)")
    del ws
    torch.cuda.empty_cache()
