"""Capture single autograd stages (forward + backward) in isolation to find the one that invalidates a CUDA-graph capture."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import autograd as A, ops

dev = "cuda"
def trial(name, fn, *inputs):
    req = [t for t in inputs if t.requires_grad]
    try:
        y = fn(*inputs); gr = torch.autograd.grad(y.float().sum(), req, allow_unused=True); torch.cuda.synchronize()   # warm-up, eager
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            with torch.enable_grad():
                y = fn(*inputs)
                yy = y.float().sum()
            f_ok = True
        g.replay(); torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            with torch.enable_grad():
                y = fn(*inputs)
                gr = torch.autograd.grad(y.float().sum(), req, allow_unused=True)
        g2.replay(); torch.cuda.synchronize()
        print(f"[{name}] forward capture ok, forward+backward capture ok", flush=True)
    except Exception as e:
        print(f"[{name}] FAILED: {str(e)[:90]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

P = lambda *s: torch.randn(*s, device=dev, requires_grad=True)
x = P(256, 1024)
trial("torch add/sum", lambda a: a * 2 + 1, x)
trial("F.cross_entropy", lambda a: torch.nn.functional.cross_entropy(a, torch.randint(0, 5, (256,), device=dev)), P(256, 5))
trial("LinearSmall N=5", lambda a, w, b: A.linear(a, w, b), x, P(5, 1024), P(5))
trial("LinearPS 1024->512", lambda a, w, b: A.linear(a, w, b), x, P(512, 1024), P(512))
trial("GeluF", lambda a: A.GeluF.apply(a), x)
trial("LayerNormPS", lambda a, w, b: A.LayerNormPS.apply(a, w, b, 1e-12), P(256, 768), P(768), P(768))
trial("DropoutF", lambda a: A.DropoutF.apply(a, 0.1, 123, None), x)
cu = torch.tensor([0, 100, 256], dtype=torch.int32, device=dev)
trial("AttentionF", lambda q: A.AttentionF.apply(q, cu, 2, 156, 2, 0.1, 5, None), P(256, 384))
xc = P(2, 32, 32, 64)
trial("ConvPS 3x3", lambda a, w: A.ConvPS.apply(a, w, None, 1, 1), xc, P(64, 64, 3, 3))
trial("ConvPS 3x3 s2", lambda a, w: A.ConvPS.apply(a, w, None, 2, 1), xc, P(128, 64, 3, 3))
trial("BatchNormTrainF", lambda a, w, b: A.BatchNormTrainF.apply(a, w, b, None, True, 1e-5, [], None), xc, P(64), P(64))
trial("MaxPoolF", lambda a: A.MaxPoolF.apply(a), xc)
trial("Up2F", lambda a: A.Up2F.apply(a), xc)
trial("AvgPoolF", lambda a: A.AvgPoolF.apply(a), xc)
