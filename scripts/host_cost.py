#!/usr/bin/env python
"""Tuning aid: host-side cost of one e2e step (upload, engine replay, loss tail, result queue), wall clock, GPU kept idle
between sections by synchronising -- shows whether the pipelined e2e loop is host- or device-bound."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vibertgrid_pytorch_b200 import synth, losses
from vibertgrid_pytorch_b200.prefetch import DevicePrefetcher, HostResultQueue
cfg = synth.CONFIGS["cfg2"]
dev = torch.device("cuda", 0)
net, kw = bench.build_net(cfg, dev)
eng = net._get_engine()
host = [bench.pin(synth.make_batch(cfg, i)) for i in range(4)]
res = [bench.to_device(b, dev, False) for b in host]
for i in range(4): net(*res[i % 4])
torch.cuda.synchronize()
def T(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
print("upload 34 tensors (to(device, non_blocking)): host %.3f ms, incl. drain %.3f ms" % T(lambda: bench.to_device(host[0], dev, True)))
print("net(*resident) full forward call:            host %.3f ms, incl. drain %.3f ms" % T(lambda: net(*res[0])))
print("engine.run only (input copies + graph replay): host %.3f ms, incl. drain %.3f ms" % T(lambda: eng.run(*res[0], want_seg=True)))
out = eng.run(*res[0], want_seg=True)
print("loss tail (aux + main):                       host %.3f ms, incl. drain %.3f ms" % T(lambda: (losses.aux_loss(net, out), losses.main_loss(net, out))))
ent = list(eng._graphs.values())[-1]
print("graph.replay() alone:                         host %.3f ms, incl. drain %.3f ms" % T(lambda: ent["graph"].replay()))
q = HostResultQueue()
r = net(*res[0])
def push_pop():
    q.push(r[4], r[0]); q.pop()
print("result queue push+pop:                        host %.3f ms, incl. drain %.3f ms" % T(push_pop))
it = iter(DevicePrefetcher((host[i % 4] for i in range(100)), dev))
print("prefetcher next():                            host %.3f ms, incl. drain %.3f ms" % T(lambda: next(it), 40))
