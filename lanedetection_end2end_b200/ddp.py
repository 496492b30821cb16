"""Data-parallel plumbing (new capability -- the reference is single-GPU):
one process per GPU, full model replica, per-replica BatchNorm statistics, and ONE
all-reduce (sum, then 1/world) over a single flat fp32 gradient buffer per step
(2 063 344 elements = 8.25 MB for the 2-lane model).  Parameters whose gradient is None
(``net.encoder.output_conv``, unused in training -- BP/Networks/ERFNet.py:84,92-93)
contribute zeros so every rank reduces the same layout.
"""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def pack(self):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def unpack(self):
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                p.grad.copy_(v)

    def __call__(self):
        """Average gradients over the group; returns the flat buffer (for tests)."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.pack()
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / world)
        self.unpack()
        return self.flat


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)
