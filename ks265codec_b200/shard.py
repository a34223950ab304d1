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


def _as_u8(b):
    """zero-copy uint8 view of a bytes-like / numpy array"""
    return b if isinstance(b, np.ndarray) and b.dtype == np.uint8 else np.frombuffer(b, np.uint8)


def gather_bitstreams(local, n_shards, group=None, device=None, as_array=False):
    """local: {shard_index: bytes-like or uint8 array} encoded by this rank.  Returns the concatenated stream on rank 0 (bytes, or a
    uint8 array with as_array=True, which saves one pass over the whole stream), None elsewhere.  Collective: every rank must call it.
    Cost per rank: one pass to pack its shards into a (pinned) staging tensor, one H2D, one NCCL gather over NVLink; rank 0 adds one D2H
    of world x cap bytes and one interleaving pass into shard order."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        parts = [_as_u8(local[s]) for s in sorted(local)]
        whole = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        return whole if as_array else whole.tobytes()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    on_gpu = dev.type == "cuda"
    sizes = torch.zeros(n_shards, dtype=torch.int64)
    for s, b in local.items():
        sizes[s] = _as_u8(b).size
    sizes = sizes.to(dev)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)          # every shard has exactly one owner
    sizes_h = sizes.cpu().numpy()
    per_rank = [int(sum(sizes_h[s] for s in assign_shards(n_shards, r, world))) for r in range(world)]
    cap = max(max(per_rank), 1)
    mine = torch.empty(cap, dtype=torch.uint8, pin_memory=on_gpu)
    mv = mine.numpy()
    off = 0
    for s in assign_shards(n_shards, rank, world):
        if s in local:
            b = _as_u8(local[s])
            mv[off:off + b.size] = b
            off += b.size
    t = mine.to(dev, non_blocking=on_gpu)
    if rank != 0:
        dist.gather(t, gather_list=None, dst=0, group=group)
        return None
    bufs = torch.empty((world, cap), dtype=torch.uint8, device=dev)
    dist.gather(t, gather_list=list(bufs.unbind(0)), dst=0, group=group)
    if on_gpu:
        host_t = torch.empty((world, cap), dtype=torch.uint8, pin_memory=True)
        host_t.copy_(bufs, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = host_t.numpy()
    else:
        host = bufs.numpy()
    out = np.empty(int(sizes_h.sum()), np.uint8)
    offs = [0] * world
    pos = 0
    for s in range(n_shards):                                           # shard order = stream order
        r, n = s % world, int(sizes_h[s])
        out[pos:pos + n] = host[r, offs[r]:offs[r] + n]
        offs[r] += n; pos += n
    return out if as_array else out.tobytes()
