"""
Data parallelism for the training step: one process per GPU (torchrun), identical replicas,
per-replica BatchNorm statistics (the semantics of the reference's torch.nn.DataParallel,
src/fusionnet_model.py:395-401) and ONE exchange per step -- a sum all-reduce of the parameter
gradients over NCCL / NVLink (gloo on CPU for the tests).  The ten never-used
``blocks*.1.projection`` weights have no gradient and are excluded
statically from the buckets, so no rank ever waits on them.
"""
import torch
import torch.distributed as dist


def used_parameters(model):
    """Parameters that receive a gradient: everything except the projection of ResNet blocks
    whose shortcut is the identity (reference src/net_utils.py:317-320)."""
    import net_utils
    unused = set()
    for root in (model.encoder, model.decoder):
        for m in root.modules():
            if isinstance(m, net_utils.ResNetBlock) and m.stride == 1 and m.in_channels == m.out_channels:
                unused.update(id(p) for p in m.projection.parameters())
    return [p for p in model.parameters() if id(p) not in unused]


class DistributedGradSync(object):
    """Bucketed gradient all-reduce (average).  Buckets follow REVERSE execution order
    (decoder first), so bucket i can be reduced while the backward of earlier layers runs."""

    def __init__(self, model, process_group=None, bucket_bytes=32 << 20):
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        params = used_parameters(model)[::-1]
        self.buckets, cur, size = [], [], 0
        for p in params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), device=b[0].device, dtype=torch.float32)
                     for b in self.buckets]
        self.broadcast_state(model)

    def broadcast_state(self, model):
        """Rank 0's parameters and buffers win (DataParallel keeps replica 0's running stats)."""
        for root in (model.encoder, model.decoder):
            for t in list(root.parameters()) + list(root.buffers()):
                dist.broadcast(t.data, src=0, group=self.group)

    def __call__(self, param_grads=None):
        flat = getattr(self, 'flat_grad', None)
        if flat is not None:            # FusedAdam flat gradient buffer: one in-place all-reduce, no copies
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
            return
        works = []
        for bucket, flat in zip(self.buckets, self.flat):
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for work, bucket, flat in zip(works, self.buckets, self.flat):
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p) / self.world
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n

    # -- overlapped form (FusionNetModel.train_step_graphed with two graphs): slices of the flat gradient buffer are reduced
    #    asynchronously on NCCL's stream while the rest of the backward runs; finish() waits and averages
    def reduce_slice(self, lo, hi):
        flat = self.flat_grad[lo:hi]
        if flat.numel() == 0:
            return
        if not hasattr(self, '_works'):
            self._works = []
        self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in getattr(self, '_works', []):
            w.wait()
        self._works = []
        self.flat_grad.div_(self.world)

    def payload_bytes(self):
        return sum(f.numel() * 4 for f in self.flat)


def attach_if_distributed(model):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        model.grad_hook = DistributedGradSync(model)
    return model


def use_flat_gradients(model, optimizer):
    """Let the gradient sync all-reduce rcfd.optim.FusedAdam's flat gradient buffer in place."""
    if model.grad_hook is not None and hasattr(optimizer, 'flat_grad'):
        model.grad_hook.flat_grad = optimizer.flat_grad
    return model
