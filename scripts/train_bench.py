"""Training-step timing at a BASELINE config: forward + backward + the reference's two optimizer steps
(train_SROIE.py:217-235: SGD for the CNN / heads, AdamW for parameters whose name contains "bert_model")."""
import os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import synth
from vibertgrid_pytorch_b200.net import ViBERTgridNet

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = synth.CONFIGS[cfg_name]
os.chdir(tempfile.mkdtemp())
synth.write_bert_dir(cfg, os.getcwd())
net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval")).cuda()
synth.fill_state_dict_(net, 0)
net.train()
bert = [p for n, p in net.named_parameters() if "bert_model" in n]
cnn = [p for n, p in net.named_parameters() if "bert_model" not in n]
opt_cnn = torch.optim.SGD(cnn, lr=1e-3, momentum=0.9, weight_decay=5e-4)
opt_bert = torch.optim.AdamW(bert, lr=1e-5, weight_decay=1e-2)
batch = synth.make_batch(cfg, 0)
img, seg, cls, coors, corpus, mask = batch
c = lambda ts: tuple(t.cuda() for t in ts)
dev = (c(img), c(seg), c(cls), c(coors), corpus.cuda(), mask.cuda())
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(3 + steps):
    if it == 3:
        torch.cuda.synchronize(); t_f = t_b = t_o = 0.0
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record()
    loss = net(*dev)
    e1.record()
    opt_cnn.zero_grad(); opt_bert.zero_grad()
    loss.backward()
    e2.record()
    opt_cnn.step(); opt_bert.step()
    e3.record()
    torch.cuda.synchronize()
    if it >= 3:
        t_f += e0.elapsed_time(e1); t_b += e1.elapsed_time(e2); t_o += e2.elapsed_time(e3)
    print(f"step {it} loss {float(loss):.4f}", flush=True)
n = steps
print(f"{cfg_name}: forward {t_f / n:.2f} ms, backward {t_b / n:.2f} ms, optimizers {t_o / n:.2f} ms, "
      f"step {(t_f + t_b + t_o) / n:.2f} ms = {cfg.batch / ((t_f + t_b + t_o) / n / 1e3):.1f} images/s; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if os.environ.get("VBG_TRAIN_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        loss = net(*dev)
        torch.cuda.synchronize()
        t0 = time.time()
        opt_cnn.zero_grad(); opt_bert.zero_grad()
        loss.backward()
        t1 = time.time()
        opt_cnn.step(); opt_bert.step()
        torch.cuda.synchronize()
    print(f"host time of backward() call: {(t1 - t0) * 1e3:.1f} ms")
    rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    if not rows:
        rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"profiled one step: {tot:.2f} ms of device time over {sum(r[2] for r in rows)} kernels")
    for k, t, n in rows[:45]:
        print(f"{t:8.3f} ms {100 * t / tot:5.1f}% x{n:<5d} {k[:110]}")
    # where the step's wall time goes: kernels as intervals on the device timeline (a graph replay: no host in the way)
    ks = [e for e in prof.events() if e.device_type.name == "CUDA" and e.time_range.end > e.time_range.start]
    if ks:
        iv = sorted((e.time_range.start, e.time_range.end) for e in ks)
        span = iv[-1][1] - iv[0][0] if iv else 0.0
        span = max(b for _, b in iv) - iv[0][0]
        busy, cur_a, cur_b = 0.0, iv[0][0], iv[0][1]
        for a, b in iv[1:]:
            if a > cur_b:
                busy += cur_b - cur_a
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        busy += cur_b - cur_a
        total = sum(b - a for a, b in iv)
        short = sum(1 for a, b in iv if b - a < 5.0)
        print(f"device timeline of the profiled step: span {span / 1e3:.2f} ms, some kernel running {busy / 1e3:.2f} ms "
              f"(idle gaps {(span - busy) / 1e3:.2f} ms), sum of kernel durations {total / 1e3:.2f} ms (overlap {(total - busy) / 1e3:.2f} ms), "
              f"{len(iv)} kernels of which {short} shorter than 5 us")
