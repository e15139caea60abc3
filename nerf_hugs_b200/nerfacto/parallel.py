"""Multi-GPU training of the torch twins: one process per GPU, rays sharded over the ranks, parameters replicated.

The reference wraps the model in `nn.DataParallel` (nerfacto/train.py:101-102: one process scatters the batch and gathers
the outputs through GPU 0).  Here every rank runs the unchanged loop body on its shard of the batch and calls
`allreduce_gradients(model)` between `loss.backward()` and `optimizer.step()`: the mean of the per-rank gradients, as
DataParallel's backward would produce for equally sized shards.  Small tensors travel in one flat bucket (one NCCL launch
instead of ~50 latency-bound ones); tensors above `big` elements (the hash tables) are reduced in place.
"""
import torch
import torch.distributed as dist


def allreduce_gradients(model, big: int = 1 << 20):
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return
  world = dist.get_world_size()
  small = []
  for p in model.parameters():
    if p.grad is None or p.numel() == 0:
      continue
    if p.numel() >= big:
      dist.all_reduce(p.grad)
      p.grad.div_(world)
    else:
      small.append(p.grad)
  if small:
    flat = torch.cat([g.reshape(-1) for g in small])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for g in small:
      g.copy_(flat[o:o + g.numel()].view_as(g)); o += g.numel()
