"""N > 1 host path on CPU: world_size-2 gloo processes shard images, gather records, max-reduce times."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from psalm_b200 import dist as PD
from psalm_b200.structures import Instances


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = PD.shard_indices(5, rank, world)
    recs = []
    for i in range(3):   # every rank contributes the same count (weak scaling: fixed per-rank batch)
        inst = Instances((4, 4))
        inst.scores = torch.full((7,), float(10 * rank + i))
        inst.pred_classes = torch.arange(7)
        inst.pred_masks = torch.ones(7, 4, 4)
        recs.append(PD.compact_record({"instances": inst}, num_queries=10))
    allr = PD.gather_records(torch.stack(recs))
    mx = PD.max_over_ranks([float(rank + 1), 5.0 - rank], "cpu")
    # per-step prediction gather (class / score + panoptic id maps) and metric-accumulator reduction
    res = []
    for i in range(2):
        inst = Instances((4, 6))
        inst.scores = torch.full((3,), float(rank) + 0.5)
        inst.pred_classes = torch.tensor([1, 2, 3]) + 10 * rank
        res.append({"instances": inst, "panoptic_seg": (torch.full((4, 6), 7 * rank + i, dtype=torch.int32), [])})
    meta, maps = PD.pack_predictions(res, num_queries=5)
    gmeta, gmaps = PD.gather_predictions(meta, maps)
    g = torch.Generator().manual_seed(rank)
    for r_ in res:
        r_["instances"].pred_masks = (torch.rand(3, 4, 6, generator=g) > 0.5).float()
    bits = PD.pack_instance_masks(res, num_queries=5)                   # [2, 5, 4, 1] uint8
    gbits = PD.gather_tensor(bits)
    back = PD.unpack_mask_bits(gbits[2 * rank:2 * rank + 2], 6)
    assert tuple(gbits.shape) == (4, 5, 4, 1) and all(
        torch.equal(back[i, :3], res[i]["instances"].pred_masks > 0) and not back[i, 3:].any() for i in range(2))
    acc = PD.reduce_sum([torch.tensor([1.0, float(rank)]), torch.ones(2, 2, dtype=torch.int64)], "cpu")
    extra = (tuple(gmeta.shape), gmeta[:, 0, 1].tolist(), gmeta[:, 4, 1].tolist(), gmaps[:, 0, 0].tolist(),
             acc[0].tolist(), acc[1].sum().item())
    out_q.put((rank, mine, allr[:, 0, 0].tolist(), allr.shape, mx, extra))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_reduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for r in res:
        assert r[2] == [0.0, 1.0, 2.0, 10.0, 11.0, 12.0]      # rank order preserved
        assert tuple(r[3]) == (6, 10, 3)
        assert r[4] == [2.0, 5.0]                               # max over ranks
        shape, cls0, cls4, map00, acc0, acc1 = r[5]
        assert shape == (4, 5, 2) and cls0 == [1.0, 1.0, 11.0, 11.0] and cls4 == [-1.0] * 4   # empty slots = -1
        assert map00 == [0, 1, 7, 8] and acc0 == [2.0, 1.0] and acc1 == 8


def test_single_process_passthrough():
    x = torch.zeros(2, 10, 3)
    assert PD.gather_records(x) is x
    assert PD.max_over_ranks([3.0], "cpu") == [3.0]


def test_mask_bit_packing_matches_numpy_packbits():
    import numpy as np
    g = torch.Generator().manual_seed(0)
    for W in (8, 13, 64, 1021):
        m = torch.rand(3, 5, W, generator=g) > 0.4
        bits = PD.pack_mask_bits(m.float())
        assert bits.dtype == torch.uint8 and tuple(bits.shape) == (3, 5, (W + 7) // 8)
        assert np.array_equal(bits.numpy(), np.packbits(m.numpy(), axis=-1))
        assert torch.equal(PD.unpack_mask_bits(bits, W), m)
