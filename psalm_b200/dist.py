"""Multi-GPU plumbing of the hot path: images are independent units, so ranks hold full
weight replicas, take disjoint images and meet only twice — a MAX reduction of the device-timed step
time and ONE all_gather of fixed-size per-image prediction records (what the reference *would* do with
detectron2.comm.all_gather, psalm/eval/segmentation_evaluation/referring_evaluation.py:162-164, but never
enables).  Backend: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Round-robin image sharding: rank r takes items i == r (mod world)."""
    return list(range(rank, n_items, world))


def max_over_ranks(values, device):
    """values: list of floats (e.g. elapsed ms) -> element-wise max over ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def compact_record(result, num_queries=100):
    """Fixed-size per-image record [num_queries, 3] fp32: (score, class, mask area) of the instance
    predictions, zero padded — the payload of the final gather."""
    inst = result["instances"]
    rec = torch.zeros(num_queries, 3, dtype=torch.float32, device=inst.scores.device)
    n = min(num_queries, inst.scores.shape[0])
    rec[:n, 0] = inst.scores[:n]
    if inst.has("pred_classes"):
        rec[:n, 1] = inst.pred_classes[:n].float()
    rec[:n, 2] = inst.pred_masks[:n].flatten(1).sum(1)
    return rec


def gather_records(local):
    """local [n_local, Q, 3] -> [world * n_local, Q, 3] in rank order (every rank holds the result)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, 0)
