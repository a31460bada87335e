"""Data-parallel plumbing: the path shards by image (one full replica per GPU); the only
exchange step is the average of the parameter gradients, done as ONE all-reduce over the engine's flat
gradient buffer (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
Reference: A1/main.py:206-208 (DDP wrap), A2/models/anchor_detr.py:323-325 (num_boxes all-reduce)."""
import torch.distributed as dist


def average_flat_grads(flat, group, scale_fn):
    """SUM all-reduce then scale by 1/world; scale_fn(tensor, s) scales in place (cdetr_scale on the GPU)."""
    world = dist.get_world_size(group)
    dist.all_reduce(flat, group=group)
    scale_fn(flat, 1.0 / world)
    return flat


def shard_seed(seed, rank):
    """Each rank draws its own synthetic shard (weak scaling: B images per GPU)."""
    return seed * 1000 + rank
