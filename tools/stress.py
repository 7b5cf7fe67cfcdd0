"""Hang / flakiness stress: many back-to-back replays of the prefill stack, the small-batch stack and a MoE block.
Rare synchronisation bugs (mbarrier parity aliasing) passed every single-launch parity test in round 2 and only showed as
one hang or fault per few thousand CTAs: this is the guard.  usage: python tools/stress.py [replays]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stack = bench.build_stack(dev, 0, 1, 8)
for M, reps in ((2048, n), (300, n), (64, 2 * n), (16, 2 * n), (1, 2 * n)):
    x = (torch.randn(M, 4096, device=dev) * 0.5).to(torch.float16)
    g, out = bench.capture(stack, x, 1)
    g.replay()
    torch.cuda.synchronize()
    ref = out.clone()
    t0 = time.time()
    for _ in range(reps):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref), f"M={M}: replay result changed"
    print(f"ok M={M}: {reps} replays of 8 layers in {time.time() - t0:.2f} s, deterministic", flush=True)
    del g
mx = bench.build_mixtral(dev, 0, 1, 2, shard=(0, 4))
for M in (1, 8, 40):
    x = (torch.randn(M, 4096, device=dev) * 0.5).to(torch.float16)
    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        bench.run_mixtral(mx, x, 1)
    torch.cuda.current_stream().wait_stream(s_)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = bench.run_mixtral(mx, x, 1)
    g.replay()
    torch.cuda.synchronize()
    ref = out.clone()
    for _ in range(n):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref), f"moe M={M}: replay result changed"
    print(f"ok moe M={M}: {n} replays, deterministic", flush=True)
print("stress ok")
