"""Micro-benchmark: what does saturated PCIe traffic (an unrelated pinned copy loop) do to launch / event latency?
A: chain of tiny kernels on one stream;  B: the same chain as a CUDA graph;  C: two streams handing an event back and
forth (kernel, record, wait, kernel ...);  D: C captured in a graph.  Each alone, under host->device copies, under
device->host copies."""
import torch

dev = torch.device("cuda", 0)
N = 200
x = torch.zeros(8 << 20, device=dev)      # ~10 us per add_: the host stays ahead of the device
y = torch.zeros(8 << 20, device=dev)
nbytes = 20 * 1024 * 1024
hbuf = torch.empty(nbytes // 4).pin_memory()
dbuf = torch.empty(nbytes // 4, device=dev)
bg = torch.cuda.Stream()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def background(kind, n=30):
    with torch.cuda.stream(bg):
        for _ in range(n):
            if kind == "h2d":
                dbuf.copy_(hbuf, non_blocking=True)
            elif kind == "d2h":
                hbuf.copy_(dbuf, non_blocking=True)


def chain():
    for _ in range(N):
        x.add_(1.0)


def pingpong():
    for _ in range(N // 2):
        with torch.cuda.stream(s1):
            x.add_(1.0)
            e = torch.cuda.Event(); e.record(s1)
        s2.wait_event(e)
        with torch.cuda.stream(s2):
            y.add_(1.0)
            e2 = torch.cuda.Event(); e2.record(s2)
        s1.wait_event(e2)


def graph_of(fn, fork=False):
    g = torch.cuda.CUDAGraph()
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        if fork:
            def body():
                cur = torch.cuda.current_stream()
                for _ in range(N // 2):
                    x.add_(1.0)
                    e = torch.cuda.Event(); e.record(cur)
                    s2.wait_event(e)
                    with torch.cuda.stream(s2):
                        y.add_(1.0)
                        e2 = torch.cuda.Event(); e2.record(s2)
                    cur.wait_event(e2)
            with torch.cuda.graph(g, stream=cs):
                body()
        else:
            with torch.cuda.graph(g, stream=cs):
                fn()
    return g


gA = graph_of(chain)
gD = graph_of(None, fork=True)


def timed(label, fn, stream, kind):
    torch.cuda.synchronize()
    if kind:
        background(kind)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        fn()
        e1.record(stream)
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / N
    torch.cuda.synchronize()
    print("%-44s background %-5s %.2f us per kernel" % (label, kind, us), flush=True)


for rep in range(2):
    for kind in (None, "h2d", "d2h"):
        timed("A chain of %d tiny kernels, one stream" % N, chain, s1, kind)
        timed("B the same chain as a CUDA graph", gA.replay, s1, kind)
        timed("C two streams, event hand-off per kernel", pingpong, s1, kind)
        timed("D C as a CUDA graph", gD.replay, s1, kind)
