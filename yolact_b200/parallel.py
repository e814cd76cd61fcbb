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


def pack_records(box, coef, cls, score, count, out=None):
    """[b,M,4] f32, [b,M,k] f32, [b,M] i64, [b,M] f32, [b] i32 -> one float32 tensor [b, 1 + M*(6+k)].
    Class ids (< 2^24) and counts are exactly representable in fp32.  CUDA tensors: ONE kernel (yb_pack_detections);
    CPU tensors (the gloo tests of the host logic): the same layout with torch ops."""
    b, M = score.shape
    k = coef.shape[2]
    if not box.is_cuda:
        return torch.cat([count.view(b, 1).float(), cls.float(), score, box.reshape(b, M * 4), coef.reshape(b, -1)], 1)
    from . import _lib
    from .output_utils import _ops_handle
    lib = _lib.load()
    rec = out if out is not None else torch.empty(b, 1 + M * (6 + k), dtype=torch.float32, device=box.device)
    box, coef, score = box.contiguous().float(), coef.contiguous().float(), score.contiguous().float()
    cls, count = cls.contiguous().long(), count.contiguous().int()
    _lib.check(lib.yb_pack_detections(_ops_handle(box.device, k), _lib.ptr(box), _lib.ptr(coef), _lib.ptr(cls), _lib.ptr(score),
                                      _lib.ptr(count), b, M, k, _lib.ptr(rec), _lib.current_stream(box.device)),
               "yb_pack_detections")
    return rec


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
    b = int(score.shape[0])
    if b < per_rank_batch:   # short last shard: zero rows (count 0)
        rec = torch.zeros(per_rank_batch, 1 + M * (6 + k), dtype=torch.float32, device=box.device)
        if b > 0:
            rec[:b] = pack_records(box, coef, cls, score, count)
    else:
        rec = pack_records(box, coef, cls, score, count)
    if world == 1:
        return unpack_records(rec, M, k)
    out = torch.empty(world * per_rank_batch, rec.shape[1], dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return unpack_records(out, M, k)
