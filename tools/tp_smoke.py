"""2-rank NCCL smoke with progress prints: init -> all_reduce -> TP QuantLinear pair -> graph capture -> replay."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(75, exit=True)
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
def P(*a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
P("init_process_group")
dist.init_process_group("nccl", device_id=dev)
P("first all_reduce")
t = torch.ones(8, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); P("ok", t[0].item())
from gptqmodel_b200 import B200QuantLinear, tp
from helpers import random_layer
up = random_layer(1024, 2048, seed=1, device="cuda"); down = random_layer(2048, 1024, seed=2, device="cuda")
mk = lambda L: B200QuantLinear.from_checkpoint_tensors(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], 4, 128, device=dev)
mu, md = mk(tp.shard_columns(up, rank, world)), mk(tp.shard_rows(down, rank, world))
x = (torch.randn(1, 1024, device=dev) * 0.5).to(torch.float16)
def step():
    h = md(mu(x)); dist.all_reduce(h); return h
P("eager step"); y = step(); torch.cuda.synchronize(); P("eager ok", float(y.float().abs().mean()))
P("P2PAllReduce init")
ar = tp.P2PAllReduce(dev, max_elems=8192)
P("peers", [hex(int(p)) for p in ar.hdl.buffer_ptrs])
for it in range(5):
    v = ((torch.arange(4096, device=dev) % 17).float() * 0.01 * (rank + 1) + it).to(torch.float16)
    ref = v.clone(); dist.all_reduce(ref)
    got = ar(v.clone()); torch.cuda.synchronize()
    P("p2p allreduce", it, "max diff", float((got.float() - ref.float()).abs().max()))
nccl_step = step
def step():
    h = md(mu(x)); ar(h); return h
y2 = step(); torch.cuda.synchronize(); P("p2p step vs nccl step", float((y2.float() - y.float()).abs().max()))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); P("warm ok; capturing")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(4): y = step()
P("captured; replay")
for _ in range(3): g.replay()
torch.cuda.synchronize(); P("replay ok", float(y.float().abs().mean()))
import time as _t
for name, fn in (("p2p", step), ("nccl", nccl_step)):
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg):
        for _ in range(50): fn()
    gg.replay(); torch.cuda.synchronize(); t0 = _t.perf_counter()
    for _ in range(20): gg.replay()
    torch.cuda.synchronize(); P(name, "us per (2 linears + allreduce):", (_t.perf_counter() - t0) / 20 / 50 * 1e6)
    del gg
del g; torch.cuda.synchronize(); dist.barrier(); P("done"); os._exit(0)
