"""Which part of the training step breaks CUDA-graph capture?  forward only / forward + backward, error modes, backward thread."""
import os, sys, tempfile, dataclasses, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import synth
from vibertgrid_pytorch_b200.net import ViBERTgridNet
from vibertgrid_pytorch_b200.plan import plan_batch

cfg = dataclasses.replace(synth.CONFIGS["mid"], ragged=False)
os.chdir(tempfile.mkdtemp()); synth.write_bert_dir(cfg, os.getcwd())
net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval")).cuda(); synth.fill_state_dict_(net, 0); net.train()
batch = synth.make_batch(cfg, 0)
c = lambda ts: tuple(t.cuda() for t in ts)
img, seg, cls, coors, corpus, mask = batch
dev = (c(img), c(seg), c(cls), c(coors), corpus.cuda(), mask.cuda())
from vibertgrid_pytorch_b200.train_engine import TrainEngine
eng = TrainEngine(net); eng.use_graphs = False
loss = eng.loss(*dev); loss.backward(); torch.cuda.synchronize(); print("eager ok", float(loss))
params = [p for p in net.parameters() if p.requires_grad]
plan = plan_batch([tuple(im.shape[-2:]) for im in dev[0]], [int(s.shape[0]) for s in dev[1]], [int(x.shape[0]) for x in dev[3]], int(dev[4].shape[1]),
                  [float(net.image_min_size[0])] * len(dev[0]), float(net.image_max_size))
st = dict(image=[im.contiguous() for im in dev[0]], coors=torch.cat([x.reshape(-1, 4) for x in dev[3]], 0).to(torch.int64).contiguous(),
          seg_ids=torch.cat([s.reshape(-1) for s in dev[1]], 0).to(torch.int32).contiguous(),
          cls=torch.cat([x.reshape(-1) for x in dev[2]], 0).to(torch.int32).contiguous(), corpus=dev[4].contiguous(), mask=dev[5].to(torch.int32).contiguous(),
          tab=torch.from_numpy(plan.table).cuda())
eng._step_seed = torch.zeros(1, dtype=torch.int64, device="cuda")
# instrument every ops.* call: report the first one after which the capturing stream is invalidated
from vibertgrid_pytorch_b200 import ops as _ops, autograd as _A
import types
_state = {"bad": None, "on": False}
def _wrap(name, fn):
    def w(*a, **k):
        r = fn(*a, **k)
        if _state["on"] and _state["bad"] is None:
            try:
                torch.cuda.is_current_stream_capturing()
            except Exception as e:
                _state["bad"] = name
                print(f"  >>> capture invalidated after ops.{name}: {str(e)[:80]}", flush=True)
        return r
    return w
for _n in dir(_ops):
    _f = getattr(_ops, _n)
    if isinstance(_f, types.FunctionType) and not _n.startswith("_"):
        setattr(_ops, _n, _wrap(_n, _f))
for _n in dir(_A):
    _c = getattr(_A, _n)
    if isinstance(_c, type) and issubclass(_c, torch.autograd.Function) and _c is not torch.autograd.Function:
        _orig = _c.backward
        def _mk(nm, ob):
            def bw(ctx, *g):
                r = ob(ctx, *g)
                if _state["on"] and _state["bad"] is None:
                    try:
                        torch.cuda.is_current_stream_capturing()
                    except Exception as e:
                        _state["bad"] = nm
                        print(f"  >>> capture invalidated after {nm}.backward: {str(e)[:80]}", flush=True)
                return r
            return staticmethod(bw)
        _c.backward = _mk(_n, _orig)
named = list(net.named_parameters())
groups = [("head", "field_type_classification_head"), ("late_fusion fuse", "late_fusion_net.fuse_embedding_net"), ("roi emb", "late_fusion_net.ROI_embedding_net"),
          ("seg head", "semantic_segmentation_head"), ("backbone.fuse", "backbone.fuse"), ("backbone.merge", "backbone.merge"), ("backbone.skip", "backbone.skip"),
          ("conv_6", "backbone.conv_6"), ("conv_5", "backbone.conv_5"), ("conv_4", "backbone.conv_4"), ("conv_3 early", "backbone.conv_3_x.early"), ("conv_3", "backbone.conv_3"),
          ("conv_2", "backbone.conv_2"), ("conv_1", "backbone.conv_1"), ("bert last layer", "bert_model.encoder.layer.1."), ("bert layer 0", "bert_model.encoder.layer.0."),
          ("bert embeddings", "bert_model.embeddings")]
for gname, pref in groups:
    sub = [p_ for n_, p_ in named if n_.startswith(pref) and p_.requires_grad]
    if not sub:
        print(f"[{gname}] no params"); continue
    try:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            with torch.enable_grad():
                l = eng._forward(plan, st)
                gr = torch.autograd.grad(l, sub, allow_unused=True)
        g.replay(); torch.cuda.synchronize()
        print(f"[{gname}] capture of forward + grad ok ({sum(x is not None for x in gr)}/{len(sub)} grads)", flush=True)
    except Exception as e:
        print(f"[{gname}] FAILED {type(e).__name__}: {str(e)[:120]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass
sys.exit(0)
for name, bwd, mode, mt in (("forward only", False, "global", True), ("fwd+bwd global", True, "global", True),
                            ("fwd+bwd thread_local", True, "thread_local", True), ("fwd+bwd global, single-thread autograd", True, "global", False)):
    try:
        torch.cuda.synchronize()
        _state["bad"] = None; _state["on"] = True
        g = torch.cuda.CUDAGraph()
        with torch.autograd.set_multithreading_enabled(mt):
            with torch.cuda.graph(g, capture_error_mode=mode):
                with torch.enable_grad():
                    l = eng._forward(plan, st)
                    if bwd:
                        gr = torch.autograd.grad(l, params, allow_unused=True)
        g.replay(); torch.cuda.synchronize()
        print(f"{name}: capture + replay ok, loss {float(l):.5f}")
    except Exception as e:
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}")
        try: torch.cuda.synchronize()
        except Exception as e2: print("  sync after failure:", str(e2)[:100])
