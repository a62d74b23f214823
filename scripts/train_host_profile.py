"""Host-side profile (cProfile) of the training step at a BASELINE config: where the Python / ctypes / autograd time of the
launch-bound backward goes.  The backward is forced onto the calling thread so cProfile sees it."""
import cProfile, io, os, pstats, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import synth
from vibertgrid_pytorch_b200.net import ViBERTgridNet

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
os.chdir(tempfile.mkdtemp())
synth.write_bert_dir(cfg, os.getcwd())
net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval")).cuda()
synth.fill_state_dict_(net, 0)
net.train()
img, seg, cls, coors, corpus, mask = synth.make_batch(cfg, 0)
c = lambda ts: tuple(t.cuda() for t in ts)
dev = (c(img), c(seg), c(cls), c(coors), corpus.cuda(), mask.cuda())


def step():
    loss = net(*dev)
    net.zero_grad()
    loss.backward()
    return loss


with torch.autograd.set_multithreading_enabled(False):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        step()
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"host issue time per step (fwd + bwd, single thread): {(t1 - t0) / 3 * 1e3:.1f} ms; with device drain {(t2 - t0) / 3 * 1e3:.1f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    torch.cuda.synchronize()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).strip_dirs().sort_stats(key).print_stats(45)
    print(f"==== by {key} (3 steps)")
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
