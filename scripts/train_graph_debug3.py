"""Which recipe lets a backward be captured AFTER an eager backward already ran on the default stream?  One variant per process."""
import sys
import torch
dev = "cuda"
variant = sys.argv[1]
x = torch.randn(256, 1024, device=dev, requires_grad=True)
y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x]); torch.cuda.synchronize()      # the user's eager step, default stream
g = torch.cuda.CUDAGraph()
try:
    if variant == "plain":
        with torch.cuda.graph(g):
            y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    elif variant == "single_thread":
        with torch.autograd.set_multithreading_enabled(False):
            with torch.cuda.graph(g):
                y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    elif variant == "side_warmup":
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    elif variant == "thread_local":
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    elif variant == "relaxed":
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    elif variant == "alias":
        with torch.cuda.graph(g):
            xa = x.detach().requires_grad_()
            y2 = (xa * 2 + 1).sum(); gr = torch.autograd.grad(y2, [xa])
    elif variant == "alias_module":
        lin = torch.nn.Linear(1024, 8).cuda()
        out = lin(x).sum(); out.backward(); torch.cuda.synchronize()          # eager step through the module, default stream
        names = [n for n, _ in lin.named_parameters()]
        with torch.cuda.graph(g):
            al = {n: p_.detach().requires_grad_() for n, p_ in lin.named_parameters()}
            with torch.nn.utils.stateless._reparametrize_module(lin, al, tie_weights=True):
                with torch.enable_grad():
                    y2 = lin(x.detach()).sum(); gr = torch.autograd.grad(y2, list(al.values()))
        with torch.no_grad():
            lin.weight.add_(1.0)                                               # an optimizer step between replays
        g.replay(); torch.cuda.synchronize()
        want = x.detach().sum(0)
        print("   replay sees the parameter storage:", bool(torch.allclose(gr[0][0], want, rtol=1e-4)), "module restored:", lin.weight is dict(lin.named_parameters())["weight"])
    elif variant == "single_thread_relaxed":
        with torch.autograd.set_multithreading_enabled(False):
            with torch.cuda.graph(g, capture_error_mode="relaxed"):
                y = (x * 2 + 1).sum(); gr = torch.autograd.grad(y, [x])
    g.replay(); torch.cuda.synchronize()
    print(f"[{variant}] ok, grad {float(gr[0].mean())}")
except Exception as e:
    print(f"[{variant}] FAILED: {str(e)[:100]}")
