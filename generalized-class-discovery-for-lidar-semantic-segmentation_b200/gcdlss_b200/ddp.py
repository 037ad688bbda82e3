"""Data-parallel gradient exchange: the one collective of the hot path.

Scans shard across ranks with no forward communication (BN statistics are per GPU, as in the
reference, which has no SyncBN).  After the wgrad kernels of a bucket have run, its flat fp32
gradient buffer is all-reduced over NCCL (NVLink 5 / NVSwitch) while backward continues on the
remaining layers.  Parameters' ``.grad`` are views into the flat buffers, so nothing is copied.

Equivalent of what ``pl.Trainer(gpus=-1)`` sets up implicitly in the reference (ref main.py:286-293).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, params, bucket_bytes: int = 32 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if self.world > 1:
            # like DistributedDataParallel: every rank starts from rank 0's parameters (ranks that constructed the model
            # under different RNG state would otherwise diverge silently).  Buffers: see broadcast_buffers().
            with torch.no_grad():
                for p in self.params:
                    dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
        # NCCL averages in the collective itself (ReduceOp.AVG): no separate divide pass over the 151 MB of gradients
        self._avg = self.world > 1 and dist.get_backend(process_group) == "nccl"
        self.trace = None            # diagnostics (tools/ddp_check.py): a list collects (event, bucket, ...) tuples
        self.measure = False         # bench.py: CUDA events around the wait in finish() -> exposed communication time
        self.exposed_events = []
        self.buckets = []            # (flat tensor, [params])
        self._owner = {}
        self._pending = []
        self._handles = []
        # reverse registration order ~ the order gradients become ready in backward
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad_ready)
        # Gradient sink (CUDA, more than one rank): the trunk's backward pass accumulates straight into the buckets and reports
        # per network stage which parameters are final, with an event recorded behind the stage's kernels
        # (functional.TrunkFunction.backward); a bucket whose last gradient arrives that way is all-reduced behind that event
        # on a gate stream, i.e. while the rest of the backward pass is still running, instead of behind everything queued so far.
        import os
        self.direct = os.environ.get("GCDLSS_DDP_DIRECT", "1") not in ("0", "")      # A/B switch for the gradient sink
        self.use_gate = os.environ.get("GCDLSS_DDP_GATE", "1") not in ("0", "")      # 0: buckets of the sink go out behind everything queued so far
        self._gate = {}              # bucket -> (sequence number, event) of the latest stage that touched it
        self._sunk = set()           # parameters whose gradient arrived through the sink in the current step
        self._stage_events = {}
        self._seq = 0
        self._gate_stream = None
        if self.world > 1 and self.params and self.params[0].is_cuda:
            from . import ops
            self._gate_stream = torch.cuda.Stream(device=self.params[0].device)
            ops.set_grad_sink(self)
        self.reset()

    def _close(self, plist):
        # every slice starts on a 16-byte boundary: the weight-gradient kernels that accumulate straight into the buckets
        # (gradient sink) use 16-byte vector reductions; the padding is all-reduced along (zeros)
        offs, n = [], 0
        for p in plist:
            offs.append(n)
            n += (p.numel() + 3) & ~3
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        for p, off in zip(plist, offs):
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._owner[p] = len(self.buckets)
        self.buckets.append((flat, list(plist)))

    def reset(self):
        """Zero the flat gradients and re-arm the per-bucket counters (call before each backward)."""
        for flat, _ in self.buckets:
            flat.zero_()
        self._pending = [len(pl) for _, pl in self.buckets]
        self._handles = []
        self._gate = {}
        self._seq = 0
        self._sunk = set()

    def _on_grad_ready(self, p):
        if p in self._sunk:          # autograd runs the (empty) accumulation of a parameter whose gradient went through the sink
            return                   # and fires its post-accumulate hook all the same: that parameter was counted by mark_ready()
        b = self._owner[p]
        if self.trace is not None:
            self.trace.append(("hook", b, tuple(p.shape), torch.cuda.current_stream().cuda_stream))
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            self._handles.append(self._launch(b))

    def _launch(self, b, gate=None):
        if self.trace is not None:
            self.trace.append(("launch", b, gate is not None, list(self._pending), torch.cuda.current_stream().cuda_stream))
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if gate is None:
            return dist.all_reduce(self.buckets[b][0], op=op, group=self.group, async_op=True)
        # the collective waits for whatever the *current* stream has queued: a stream that has queued nothing but the wait for
        # the stage's event lets it start as soon as the bucket's gradients are final
        self._gate_stream.wait_event(gate)
        with torch.cuda.stream(self._gate_stream):
            return dist.all_reduce(self.buckets[b][0], op=op, group=self.group, async_op=True)

    # ---- gradient sink interface (functional.TrunkFunction.backward)
    def grad_ptr(self, p):
        """Device address of the bucket slice that is ``p.grad`` (None if the parameter is not ours or its .grad was replaced).
        ``direct = False`` switches the sink off: a step that sends more than one backward pass through the same parameters
        (Stage 2: the student runs on two batches) must let autograd sum the passes before a bucket is reduced."""
        if not self.direct:
            return None
        b = self._owner.get(p)
        if b is None or p.grad is None:
            return None
        flat = self.buckets[b][0]
        ptr = p.grad.data_ptr()
        if not (flat.data_ptr() <= ptr < flat.data_ptr() + flat.numel() * 4) or not p.grad.is_contiguous() or ptr % 16:
            return None
        return ptr

    def stage_event(self, i: int):
        ev = self._stage_events.get(i)
        if ev is None:
            ev = self._stage_events[i] = torch.cuda.Event()
            ev.record()              # torch creates the CUDA event lazily, at the first record: the library needs the handle
        return ev

    def mark_ready(self, params, event):
        """The gradients of ``params`` are complete once ``event`` has happened (they were accumulated into the buckets by the
        kernels themselves; autograd sees no gradient for them and fires no hook)."""
        self._seq += 1
        touched = set()
        for p in params:
            b = self._owner[p]
            self._pending[b] -= 1
            self._gate[b] = (self._seq, event)
            self._sunk.add(p)
            touched.add(b)
        if self.world > 1:
            for b in sorted(touched):
                if self._pending[b] == 0:
                    self._handles.append(self._launch(b, gate=self._gate[b][1] if self.use_gate else None))

    def broadcast_buffers(self, module):
        """Rank 0's buffers (BN running statistics) to every rank; call once after construction."""
        if self.world > 1:
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            for b in module.buffers():
                dist.broadcast(b.data, src=src, group=self.group)

    def check_views(self):
        """``.grad`` of every parameter must still alias its flat bucket (``optimizer.zero_grad(set_to_none=True)`` detaches
        them silently; use ``reset()`` instead)."""
        for flat, plist in self.buckets:
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            for p in plist:
                if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("GradBucketReducer: a parameter's .grad no longer aliases its all-reduce bucket; "
                                       "zero gradients with reducer.reset(), not optimizer.zero_grad(set_to_none=True)")

    def finish(self):
        """Wait for the outstanding all-reduces and turn sums into means."""
        self.check_views()
        if self.world > 1:
            for b, left in enumerate(self._pending):       # parameters that received no gradient this step
                if left > 0:
                    self._handles.append(self._launch(b))
            if self.measure:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for h in self._handles:
                h.wait()
            if not self._avg:
                for flat, _ in self.buckets:
                    flat.div_(self.world)
            if self.measure:
                e1.record()
                self.exposed_events.append((e0, e1))
        self._handles = []

    def grad_bytes(self) -> int:
        return sum(f.numel() * 4 for f, _ in self.buckets)


def shard_scans(n_scans_global: int, rank: int, world: int):
    """Indices of the scans rank ``rank`` processes (contiguous blocks, like a DistributedSampler without shuffle).
    The ``n_scans_global % world`` trailing scans go one each to the first ranks, so none is dropped."""
    per, extra = divmod(n_scans_global, world)
    start = rank * per + min(rank, extra)
    return list(range(start, start + per + (1 if rank < extra else 0)))
