"""Input pipelining for the drop-in: upload batch i+1 on a side stream while the forward of batch i runs.

The reference's loops move every tensor of a batch with a blocking per-tensor ``.to(device)`` right before the forward
(``pipeline/train_val_utils.py:257-262``): ~34 small copies per batch (25 MB at the BASELINE shape) serialised with the
compute.  ``DevicePrefetcher`` wraps any iterable of host batches in the reference collate layout
``(image, seg_indices, segment_classes, coors, corpus, mask)`` -- tuples of tensors or tensors, ideally pinned -- and yields the
same structure on the device; the copies of the next batch overlap the kernels of the current one.

    for image, seg, cls, coors, corpus, mask in DevicePrefetcher(loader, device):
        loss, pred_mask, pred_ss, gt, pred = net(image, seg, cls, coors, corpus, mask)
"""
from __future__ import annotations

import torch


def _map(batch, fn):
    return [tuple(fn(t) for t in x) if isinstance(x, (tuple, list)) else fn(x) for x in batch]


class DevicePrefetcher:
    def __init__(self, batches, device):
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher uploads to a CUDA device")

    def _upload(self, batch, stream):
        with torch.cuda.stream(stream):
            dev = _map(batch, lambda t: t.to(self.device, non_blocking=True))
        ev = torch.cuda.Event()
        ev.record(stream)
        return dev, ev

    def _hand_over(self, item):
        dev, ev = item
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        _map(dev, lambda t: t.record_stream(cur))       # allocated on the copy stream, consumed on the compute stream
        return dev

    def __iter__(self):
        stream = torch.cuda.Stream(self.device)
        it = iter(self.batches)
        try:
            nxt = self._upload(next(it), stream)
        except StopIteration:
            return
        for b in it:
            cur = self._hand_over(nxt)
            nxt = self._upload(b, stream)               # in flight while the caller computes on `cur`
            yield cur
        yield self._hand_over(nxt)


class HostResultQueue:
    """Device->host read of per-step results without stalling the launch of the next step.

    ``push(*tensors)`` enqueues an asynchronous copy of the step's results into pinned host buffers on the current stream;
    ``pop()`` blocks only until THAT copy has landed and returns the host tensors.  Reading step i's results after step i+1
    has been launched keeps the GPU busy while the host prepares the next step (the reference's ``loss.item()`` right after
    every forward, ``pipeline/train_val_utils.py:269``, drains the device every step).

        q = HostResultQueue()
        for batch in DevicePrefetcher(loader, device):
            loss, pm, ps, gt, pred = net(*batch)
            q.push(pred, loss)
            if len(q) > 1:
                pred_h, loss_h = q.pop()        # results of the previous step
        while len(q): pred_h, loss_h = q.pop()
    """

    def __init__(self):
        self._items = []
        self._free = {}

    def __len__(self):
        return len(self._items)

    def push(self, *tensors):
        host = []
        for t in tensors:
            t = t.detach()
            key = (tuple(t.shape), t.dtype)
            pool = self._free.setdefault(key, [])
            h = pool.pop() if pool else torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            host.append(h)
        ev = torch.cuda.Event()
        ev.record()
        self._items.append((host, ev))

    def pop(self):
        host, ev = self._items.pop(0)
        ev.synchronize()
        out = [h.clone() for h in host]
        for h in host:
            self._free.setdefault((tuple(h.shape), h.dtype), []).append(h)
        return out
