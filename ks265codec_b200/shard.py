"""GOP sharding across ranks + gather of the encoded NAL units.

A shard is one closed GOP (`-iper` pictures starting with an IDR): shards are independent, so rank r of world N
encodes shards r, r+N, r+2N, ... with no data-path collective; the only exchange is the final gather of the
variable-length Annex-B byte strings to rank 0, in shard order (sizes all_gather + padded payload gather).
Works on any torch.distributed backend: NCCL (tensors on the rank's GPU, NVLink/NVSwitch) or gloo (CPU tests).
"""
import numpy as np


def assign_shards(n_shards, rank, world):
    """shard indices owned by `rank`"""
    return list(range(rank, n_shards, world))


def shard_frames(n_frames, iper):
    """[(first_frame, n_frames_in_shard)] for a sequence of n_frames with intra period iper"""
    return [(s, min(iper, n_frames - s)) for s in range(0, n_frames, iper)]


def gather_bitstreams(local, n_shards, group=None, device=None):
    """local: {shard_index: bytes-like} encoded by this rank.  Returns the concatenated stream (bytes) on rank 0,
    None elsewhere.  Collective: every rank must call it."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return b"".join(bytes(local[s]) for s in sorted(local))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    sizes = torch.zeros(n_shards, dtype=torch.int64, device=dev)
    for s, b in local.items():
        sizes[s] = len(b)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)          # every shard has exactly one owner
    sizes_h = sizes.cpu().numpy()
    per_rank = [int(sum(sizes_h[s] for s in assign_shards(n_shards, r, world))) for r in range(world)]
    cap = max(per_rank) if per_rank else 0
    mine = np.zeros(cap, np.uint8)
    off = 0
    for s in assign_shards(n_shards, rank, world):
        b = np.frombuffer(bytes(local[s]), np.uint8) if s in local else np.zeros(0, np.uint8)
        mine[off:off + b.size] = b
        off += b.size
    t = torch.from_numpy(mine).to(dev)
    if rank == 0:
        bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.gather(t, gather_list=bufs, dst=0, group=group)
        host = [b.cpu().numpy() for b in bufs]
        offs = [0] * world
        out = []
        for s in range(n_shards):
            r = s % world
            out.append(host[r][offs[r]:offs[r] + int(sizes_h[s])].tobytes())
            offs[r] += int(sizes_h[s])
        return b"".join(out)
    dist.gather(t, gather_list=None, dst=0, group=group)
    return None
