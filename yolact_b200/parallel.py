"""Batch sharding across the GPUs of one box: one process per GPU, replicated weights, images split
contiguously on dim 0, NO collective in the forward pass (no cross-image op exists anywhere in
Yolact.forward / Detect / postprocess).  The only exchange is one all_gather of fixed-size padded
detection records at the end of a global batch -- the analogue of the reference's
CustomDataParallel.gather (list concat, eval.py:630-634).  Rank-major order == original batch order.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous [start, stop) of the global batch owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_records(box, coef, cls, score, count):
    """[b,M,4] f32, [b,M,k] f32, [b,M] i64, [b,M] f32, [b] i32 -> one float32 tensor [b, 1 + M*(6+k)].
    Class ids (< 2^24) and counts are exactly representable in fp32."""
    b, M = score.shape
    return torch.cat([count.view(b, 1).float(), cls.float(), score, box.reshape(b, M * 4), coef.reshape(b, -1)], 1)


def unpack_records(rec, M, k):
    b = rec.shape[0]
    count = rec[:, 0].to(torch.int32)
    o = 1
    cls = rec[:, o:o + M].to(torch.int64)
    o += M
    score = rec[:, o:o + M]
    o += M
    box = rec[:, o:o + 4 * M].reshape(b, M, 4)
    o += 4 * M
    coef = rec[:, o:o + k * M].reshape(b, M, k)
    return box, coef, cls, score, count


def gather_detections(box, coef, cls, score, count, per_rank_batch, group=None):
    """All ranks contribute `per_rank_batch` images (pad the last shard); returns the global tensors
    in original batch order on every rank.  ~16 KB per image: negligible on NVLink."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    M, k = score.shape[1], coef.shape[2]
    rec = pack_records(box, coef, cls, score, count)
    if rec.shape[0] < per_rank_batch:
        pad = torch.zeros(per_rank_batch - rec.shape[0], rec.shape[1], dtype=rec.dtype, device=rec.device)
        rec = torch.cat([rec, pad], 0)
    if world == 1:
        return unpack_records(rec, M, k)
    out = torch.empty(world * per_rank_batch, rec.shape[1], dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return unpack_records(out, M, k)
