"""Batch-sharded data parallelism for the decoder training step.

The reference is single-process (no ``torch.distributed`` anywhere); BASELINE.json adds "work is
partitioned across one 8xB200 box by batch sharding; training steps use NCCL gradient allreduce over
NVLink; decode is shard-local with no collective".  One process per GPU, identical replicas:

* ``shard_batch``     rank r takes samples [r*B/g, (r+1)*B/g) of every batch-leading tensor;
* ``GradAllReducer``  buckets the parameters in reverse registration order (the order backward
  produces gradients), and as soon as the last gradient of a bucket has been accumulated it packs the
  bucket and launches an asynchronous all-reduce -- communication overlaps the rest of backward.
  ``finish()`` waits, averages and scatters the result back into ``.grad``.

Works with any ``torch.distributed`` backend (``nccl`` on the GPU box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch(tensors, rank, world_size):
    """Slice dim 0 of every tensor (None passes through) into this rank's contiguous shard."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        B = t.shape[0]
        if B % world_size:
            raise ValueError(f"batch {B} is not divisible by world size {world_size}")
        per = B // world_size
        out.append(t[rank * per:(rank + 1) * per])
    return out


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank ``src``'s parameters and buffers."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


class GradAllReducer:
    def __init__(self, module, bucket_bytes=32 << 20, group=None, average=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        params = [p for p in module.parameters() if p.requires_grad]
        self.buckets = []          # list of lists of params
        cur, cur_bytes = [], 0
        for p in reversed(params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[p] = bi
        self._flat = [torch.empty(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device)
                      for b in self.buckets]
        self._pending = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self._paused = False
        self.reset()

    def pause(self):
        """Gradient accumulation: ignore gradient hooks until ``resume()`` (micro-batches before the
        last one must not trigger the all-reduce)."""
        self._paused = True

    def resume(self):
        self._paused = False

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _launch(self, bi):
        flat = self._flat[bi]
        off = 0
        for p in self.buckets[bi]:
            n = p.numel()
            g = p.grad
            if g is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(g.reshape(-1))
            off += n
        if self.world > 1:
            self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group,
                                             async_op=True)

    def _on_grad(self, p):
        if self._paused:
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """Wait for every bucket, write the (averaged) gradients back, re-arm for the next step."""
        for bi, b in enumerate(self.buckets):
            if self._pending[bi] > 0:      # some parameter received no gradient this step
                self._launch(bi)
            if self._work[bi] is not None:
                self._work[bi].wait()
            flat = self._flat[bi]
            if self.average and self.world > 1:
                flat.div_(self.world)
            off = 0
            for p in b:
                n = p.numel()
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
