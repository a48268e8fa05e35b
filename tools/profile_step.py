"""Profile one eval_seg step at the bench workload.
  python tools/profile_step.py torch   -> kernel-time table (torch.profiler) in gpurun_out/step_profile.txt
  ncu --profile-from-start off ... python tools/profile_step.py ncu   -> profiled region = one step"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psalm_b200 import synth  # noqa: E402
from psalm_b200.layout import PsalmConfig  # noqa: E402
from psalm_b200.psalm import PSALM  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "torch"
    B = int(os.environ.get("B", "1"))
    cfg = PsalmConfig()
    sd = synth.synth_state_dict(cfg, seed=0, device="cuda")
    m = PSALM(sd, cfg, torch.bfloat16, "cuda", "panoptic")
    del sd
    inp = synth.synth_inputs(batch=B, height=1024, width=1024, task="panoptic", n_classes=134, seed=1)
    m.is_thing_list = inp["is_thing_list"]
    images = inp["images"].cuda()
    plan = m.make_plan(inp["input_ids"], inp["attention_mask"], (1024, 1024), inp["class_name_ids"], inp["cls_indices"],
                       inp["class_name_embedding_indices"]).to("cuda")

    def step():
        out = m.forward_core(images, plan)
        return m.post_process(out, (1024, 1024), inp["seg_info"])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if mode == "ncu":
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    from torch.profiler import ProfilerActivity, profile, record_function
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for name, fn in (("all", step),):
            with record_function(name):
                fn()
        torch.cuda.synchronize()
    tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70)
    open(os.path.join(ROOT, "gpurun_out", "step_profile.txt"), "w").write(tab)
    print(tab[-6000:])
    # stage timing with events
    stages = {}
    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        stages[name] = a.elapsed_time(b)
        return r
    from psalm_b200 import sequence as SEQ
    import torch.nn.functional as F
    for rep in range(2):
        toks, sizes = timed("swin", lambda: m.model.vision_tower.forward_tokens(images))
        h5, w5 = sizes[3]
        res5 = toks[3].view(B, h5, w5, -1).permute(0, 3, 1, 2)
        img_tok = timed("projector", lambda: m.model.mm_projector(res5))
        embeds = timed("embeds", lambda: SEQ.materialize_embeds(plan, m.model.embed_tokens, img_tok, m.seg_query))
        hidden = timed("phi", lambda: m.model.phi(embeds, None))
        seg_q = F.linear(SEQ.gather_seg_query(plan, hidden), *m.proj["seg_query_projector"])
        cls_emb = F.linear(SEQ.pool(plan.cls_pool, hidden), *m.proj["class_name_projector"])
        mf, ms, mss = timed("pixel_decoder", lambda: m.pixel_decoder.forward_tokens(toks, sizes))
        out = timed("predictor", lambda: m.predictor.forward_tokens(ms, mss, mf, sizes[0], seg_q, None, cls_emb))
        out["mask_size"] = sizes[0]
        timed("post_process", lambda: m.post_process(out, (1024, 1024), inp["seg_info"]))
    print("T =", plan.T, "stages (ms):", {k: round(v, 3) for k, v in stages.items()}, "sum", round(sum(stages.values()), 2))
    open(os.path.join(ROOT, "gpurun_out", "stage_times.txt"), "w").write(repr(stages))


if __name__ == "__main__":
    main()
