"""Diagnostic: where does the bf16 path drift at the bench configuration (1024^2, Phi-1.5)?  B = 1 vs B = 4, eager vs
graph, free-running and teacher-forced stage errors against one oracle pass."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import parity  # noqa: E402
from psalm_b200.layout import PsalmConfig  # noqa: E402
from psalm_b200.psalm import PSALM  # noqa: E402


def main():
    cfg = PsalmConfig()
    dt = {"bf16": torch.bfloat16, "f16": torch.float16}[os.environ.get("DT", "bf16")]
    sd = bench.bench_weights(cfg, dt)
    B = 4
    inp = bench.bench_inputs(B, 1)
    _, ores, it = bench.oracle_eval(sd, inp, 0, 32, relaxed=True)
    m = PSALM(sd, cfg, dt, "cuda", "panoptic", use_cuda_graph=False)
    m.is_thing_list = inp["is_thing_list"]
    for nb in (1, 4):
        plan = m.make_plan(inp["input_ids"][:nb], inp["attention_mask"][:nb], (1024, 1024), inp["class_name_ids"][:nb],
                           inp["cls_indices"][:nb], inp["class_name_embedding_indices"][:nb]).to("cuda")
        tr = {}
        out = m.forward_core(inp["images"][:nb].cuda(), plan, trace=tr)
        one = {k: (v[:1] if torch.is_tensor(v) else v) for k, v in out.items()}
        tr1 = {k: ([t[:1] for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v[:1] if torch.is_tensor(v) else v))
               for k, v in tr.items()}
        print("B=%d eager final : %s" % (nb, parity.fmt(parity.final_metrics(one, it))))
        print("B=%d eager free  : %s" % (nb, parity.fmt(parity.free_stage_errors(tr1, it))))
        if nb == 1:
            print("B=1 forced       : %s" % parity.fmt(parity.forced_stage_errors(m, it, inp["images"][:1])))
            base = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}
        else:
            d = (out["pred_masks"][:1].float() - base["pred_masks"].float())
            print("B=4 vs B=1 pred_masks l2rel %.3e" % float(d.norm() / base["pred_masks"].float().norm()))
            for k in ("hidden", "mask_features", "seg_query"):
                pass
    mg = PSALM(sd, cfg, dt, "cuda", "panoptic", use_cuda_graph=True)
    mg.is_thing_list = inp["is_thing_list"]
    plan = mg.make_plan(inp["input_ids"], inp["attention_mask"], (1024, 1024), inp["class_name_ids"], inp["cls_indices"],
                        inp["class_name_embedding_indices"]).to("cuda")
    g = mg.forward_core_graphed(inp["images"].cuda(), plan)
    one = {k: (v[:1] if torch.is_tensor(v) else v) for k, v in g.items() if k != "post"}
    print("B=4 graph final  : %s" % parity.fmt(parity.final_metrics(one, it)))


if __name__ == "__main__":
    main()
