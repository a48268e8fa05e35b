"""`PSALM` — the drop-in object for the reference's segmentation inference path.

Keeps the operator surface of `PSALM(PhiForCausalLM, LlavaMetaForCausalLM)` that the eval scripts use
(reference language_model/llava_phi.py:146, psalm/model/llava_arch.py:49-63):
    model.eval_seg(input_ids=..., attention_mask=..., images=..., seg_info=..., class_name_ids=...,
                   cls_indices=..., class_name_embedding_indices=..., token_refer_id=...,
                   refer_embedding_indices=..., is_thing_list=...)      -> list[dict]   (llava_phi.py:1317)
    model.get_model(), model.get_vision_tower(), model.encode_images(images),
    model.get_vision_tower_feature(images), model.pixel_decoder.forward_features(dict),
    model.predictor(x, mask_features, None, seg_query, SEG_embedding, class_name_embedding, None)
and loads the reference checkpoint layout unchanged (layout.py / loader.py).

With `use_cuda_graph=True` the network and the device part of the task heads replay from one CUDA graph;
the returned tensors then alias the graph's static buffers and are valid until the next `eval_seg` call on
the same model (consume or clone them first, as the reference's eval loops do with `evaluator.process`).

How it differs from the reference on purpose (results are unchanged):
  * Swin runs ONCE per image (the reference runs it twice on identical input, llava_phi.py:449 and :223);
  * every image of the batch is post-processed (the reference returns inside the loop, llava_phi.py:1472);
  * no CPU / PyTorch fallback for the hot operators: a missing CUDA library raises.
"""
import contextlib
from collections import OrderedDict

import os

import torch
import torch.nn.functional as F

from . import postprocess as PP
from . import sequence as SEQ
from .layout import PsalmConfig
from .mask_decoder import MultiScaleMaskedTransformerDecoderForOPTPreTrain
from .phi import PhiModel
from .pixel_decoder import MSDeformAttnPixelDecoder
from .projector import ResNetSwin
from .swin import SwinTransformer


class StagedImages:
    """A batch of images whose host->device copy was enqueued on the model's copy stream by
    `PSALM.stage_images`; `PSALM.eval_seg(images=staged, ...)` waits for it on the compute stream."""

    def __init__(self, slot, ready):
        self.slot, self.ready = slot, ready          # slot = [device buffer, event "consumer is done"]
        self.tensor = slot[0]
        self.shape, self.dtype = slot[0].shape, slot[0].dtype


class PendingSeg:
    """Handle of a submitted `PSALM.eval_seg_async` call: the network + the device part of the task heads are queued
    (one graph replay), the few integers the host merge needs are on their way to pinned memory.  `result()` finishes
    the post-processing ON THE CALLER'S CURRENT STREAM (after making it wait for the pass), so a caller that wants the
    host work of image batch k to overlap the device work of batch k+1 submits k+1 first and calls `result()` of k
    under a side stream.  The tensors of the result live in the lane's static buffers: they stay valid until the next
    submission on the same lane."""

    def __init__(self, model, out, image_hw, seg_info, boxes, done, hostvecs, thing_list, thresholds):
        self.model, self.out, self.image_hw, self.seg_info, self.boxes = model, out, image_hw, seg_info, boxes
        self.done, self.hostvecs, self.thing_list, self.thresholds = done, hostvecs, thing_list, thresholds

    def result(self):
        m = self.model
        torch.cuda.current_stream(m.device).wait_event(self.done)
        self.done.synchronize()                              # pinned host vectors are complete
        keep = (getattr(m, "is_thing_list", None), m.object_mask_threshold, m.overlap_threshold)
        m.is_thing_list, (m.object_mask_threshold, m.overlap_threshold) = self.thing_list, self.thresholds
        try:
            return m.post_process(self.out, self.image_hw, self.seg_info, self.boxes, hostvecs=self.hostvecs)
        finally:
            m.is_thing_list, m.object_mask_threshold, m.overlap_threshold = keep


class PSALMModel:
    """`model.model` of the reference (PSALMModel(LlavaMetaModel, PhiModel), llava_phi.py:52): owns the
    LLM, the vision tower and the projector."""

    def __init__(self, sd, cfg, dtype, device):
        self.phi = PhiModel(sd, "model.", cfg.phi, dtype, device)
        self.embed_tokens = self.phi.embed_tokens
        self.vision_tower = SwinTransformer(sd, "model.vision_tower.", cfg.swin, dtype, device)
        self.mm_projector = ResNetSwin(sd, "model.mm_projector.", dtype, device)

    def get_vision_tower(self):
        return self.vision_tower

    def __call__(self, inputs_embeds=None, attention_mask=None, **_):
        return self.phi(inputs_embeds, attention_mask)


class PSALM:
    def __init__(self, state_dict, cfg: PsalmConfig = PsalmConfig(), dtype=torch.bfloat16, device="cuda",
                 seg_task="panoptic", use_cuda_graph=False):
        self._check_runtime(device)
        self.use_cuda_graph = use_cuda_graph
        # one fused kernel for the task heads (16-bit storage); fp32 parity runs keep the exact torch path
        self._fused_postprocess = dtype != torch.float32
        self.overlap_branches = not os.environ.get("PSALM_NO_OVERLAP")   # pixel decoder || LLM prefill on two streams
        self.cfg, self.dtype, self.device = cfg, dtype, torch.device(device)
        sd = state_dict
        cv = lambda t: t.to(device=device, dtype=dtype).contiguous()  # noqa: E731
        self.model = PSALMModel(sd, cfg, dtype, device)
        self.pixel_decoder = MSDeformAttnPixelDecoder(sd, "pixel_decoder.", cfg.mask, dtype, device)
        self.predictor = MultiScaleMaskedTransformerDecoderForOPTPreTrain(sd, "predictor.", cfg.mask, dtype, device)
        self.seg_query = cv(sd["seg_query"])
        self.proj = {n: (cv(sd[n + ".weight"]), cv(sd[n + ".bias"]))
                     for n in ("seg_query_projector", "SEG_token_projector", "class_name_projector", "region_projector")
                     if n + ".weight" in sd}     # region_projector is optional in the checkpoint contract (loader.py)
        # lm_head (no bias in the reference, llava_phi.py:191): only the chat / decode path reads it
        self.lm_head = None
        if "lm_head.weight" in sd:
            self.lm_head = (cv(sd["lm_head.weight"]), cv(sd["lm_head.bias"]) if "lm_head.bias" in sd else None)
        self.num_queries = cfg.mask.num_queries
        self.test_topk_per_image = cfg.mask.num_queries
        self.size_divisibility = cfg.mask.size_divisibility
        # panoptic thresholds: the reference hard-codes 0.8 / 0.8 (llava_phi.py:331-332)
        self.object_mask_threshold = cfg.mask.object_mask_threshold
        self.overlap_threshold = cfg.mask.overlap_threshold
        self.set_task(seg_task)

    @contextlib.contextmanager
    def _precision_scope(self):
        """fp32 models run true fp32 library GEMMs / convolutions (cuDNN and cuBLAS would otherwise pick TF32);
        the global switches are restored on exit, other models in the process are not affected."""
        if self.dtype != torch.float32:
            yield
            return
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            yield
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    @staticmethod
    def _check_runtime(device):
        """No CPU path and no silent fallback: refuse to construct without a GPU and the built kernels."""
        if torch.device(device).type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("psalm_b200.PSALM needs a CUDA device: there is no CPU implementation of the hot path")
        from . import _lib
        _lib.lib()

    @property
    def fused_postprocess(self):
        """Fused task-head kernels: 16-bit storage, and not the region task (its head is a [K,Q] score table over the
        plain thresholded masks, done with torch ops)."""
        return self._fused_postprocess and not getattr(self, "region_on", False)

    @fused_postprocess.setter
    def fused_postprocess(self, v):
        self._fused_postprocess = bool(v)

    # ---- configuration (llava_phi.py:268-301) ----------------------------------------------------
    def set_task(self, seg_task):
        if seg_task not in ("semantic", "instance", "panoptic", "referring", "region"):
            raise NotImplementedError("SEG_TASK %r (the video variant is outside this build's scope)" % seg_task)
        self.seg_task = seg_task
        self.semantic_on = seg_task in ("semantic", "panoptic")
        self.instance_on = seg_task in ("instance", "panoptic")
        self.panoptic_on = seg_task == "panoptic"
        self.referring_on = seg_task == "referring"
        self.region_on = seg_task == "region"
        self.sem_seg_postprocess_before_inference = self.instance_on or self.panoptic_on or self.referring_on or self.region_on

    @classmethod
    def from_state_dict(cls, sd, **kw):
        return cls(sd, **kw)

    @classmethod
    def from_pretrained(cls, model_path, mask_decoder_cfg=None, **kw):
        """Checkpoint directory of the reference (config.json + safetensors / bin shards), keys unchanged
        (psalm/model/builder.py:55 calls this on the reference class)."""
        from .builder import from_pretrained
        return from_pretrained(cls, model_path, mask_decoder_cfg, **kw)

    def generate(self, input_ids=None, images=None, max_new_tokens=32, do_sample=False, temperature=1.0, top_p=None,
                 eos_token_id=None, **_unused):
        """Chat / decoding path (psalm/serve/cli.py:89-96): prefill + autoregressive decode with a paged KV cache
        (psalm_b200/generate.py).  Returns the NEW token ids [B, n_new]."""
        from .generate import generate
        return generate(self, input_ids, images, max_new_tokens, do_sample, temperature, top_p, eos_token_id)

    # ---- LlavaMetaForCausalLM methods the training / serving scripts call by name -------------------
    def initialize_vision_tokenizer(self, model_args, tokenizer):
        """llava_arch.py:181-217.  The released checkpoints set mm_use_im_patch_token = mm_use_im_start_end = False, for
        which the reference method does nothing; growing the embedding table is a training-time operation."""
        if getattr(model_args, "mm_use_im_patch_token", False) or getattr(model_args, "mm_use_im_start_end", False):
            raise NotImplementedError("adding image patch / start / end tokens resizes the embedding table: training-time "
                                      "surface, outside this inference build")

    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images,
                                             class_name_embedding_indices=None, class_name_ids=None, cls_indices=None,
                                             instances=None, token_refer_id=None, refer_embedding_indices=None):
        """llava_phi.py:767-971: sentinel ids -> embeddings.  Returns the reference's tuple
        (input_ids=None, attention_mask, past_key_values, inputs_embeds, labels, seg_query_mask,
         class_name_embedding_indices, region_embedding_masks, refer_embedding_indices) for the image-prefill case."""
        with self._precision_scope():
            img_tok = self.encode_images(images.to(self.device))
        plan = self.make_plan(input_ids, attention_mask, images.shape[-2:], class_name_ids, cls_indices,
                              class_name_embedding_indices, token_refer_id, refer_embedding_indices)
        region_mask = None
        region_feat = None
        if plan.region_pos is not None:    # llava_phi.py:791-797: region features from the instances' region masks
            if instances is None:
                raise ValueError("<region> tokens in the prompt need `instances` with region_masks (llava_phi.py:791-792)")
            from .region import region_inputs
            plan.region_points, plan.region_image, counts = region_inputs([dict(instances=i) for i in instances])
            assert counts == plan.region_counts, "the munber of <region> tokens and regions needs to be same"
        plan = plan.to(self.device)
        if plan.region_pos is not None:
            region_feat = self._region_features(img_tok, plan)
        embeds = SEQ.materialize_embeds(plan, self.model.embed_tokens, img_tok, self.seg_query, region_feat)
        B, T = plan.B, plan.T
        flat = lambda pos: torch.zeros(B * T, device=self.device).index_fill_(0, pos, 1.0).view(B, T)   # noqa: E731
        seg_query_mask = flat(plan.seg_pos)
        cls_idx = None
        if plan.cls_pool is not None:     # class index (1-based) of every position, 0 elsewhere (:671-673)
            cls_idx = ((plan.cls_pool > 0).float() * torch.arange(1, plan.cls_pool.shape[1] + 1, device=self.device)
                       .view(1, -1, 1)).sum(1).long()
        ref_idx = (plan.refer_pool[:, 0] > 0).long() if plan.refer_pool is not None else None
        if plan.region_pos is not None:
            region_mask = flat(plan.region_pos)
        return None, plan.attention_mask, past_key_values, embeds, labels, seg_query_mask, cls_idx, region_mask, ref_idx

    def _region_features(self, img_tok, plan):
        """[R, hidden] pooled region features (region_pooling, context_cluster.py:333-400) from the projector tokens."""
        from . import kernels
        n_img = img_tok.shape[1]
        h = w = int(round(n_img ** 0.5))
        if h * w != n_img:      # context_cluster.py:355 takes h = w = int(sqrt(n)): square maps only
            raise ValueError("region prompts need a square projector map (got %d tokens)" % n_img)
        return kernels.region_pool(img_tok.contiguous(), plan.region_points, plan.region_image, h, w)

    # ---- LlavaMetaForCausalLM surface --------------------------------------------------------------
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.get_vision_tower()

    def encode_images(self, images):
        """llava_phi.py:448-451."""
        with self._precision_scope():
            feats = self.get_vision_tower()(images)
            return self.model.mm_projector(feats[-1])

    def get_vision_tower_feature(self, images):
        """llava_phi.py:222-230."""
        with self._precision_scope():
            f = self.get_vision_tower()(images)
        return dict(res2=f[0], res3=f[1], res4=f[2], res5=f[3])

    # ---- the hot path -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward_core(self, images, plan, trace=None):
        """Device-only part of eval_seg: images [B,3,H,W] on device, `plan` a SequencePlan on device.
        Returns dict(pred_masks [B,Q,H4*W4], mask_size, pred_class_name_logits, pred_SEG_logits).
        `trace`: optional dict that receives the stage outputs (token-major), for the per-stage parity tests."""
        with self._precision_scope():
            return self._forward_core(images, plan, trace)

    def _forward_core(self, images, plan, trace=None):
        toks, sizes = self.model.vision_tower.forward_tokens(images)                 # Swin, once
        h5, w5 = sizes[3]
        res5 = toks[3].view(toks[3].shape[0], h5, w5, -1).permute(0, 3, 1, 2)
        img_tok = self.model.mm_projector(res5)                                        # [B,n_img,hidden]
        # The pixel decoder needs only the Swin maps, the LLM only the projector tokens: the two branches run on two
        # streams (also inside a captured graph) and meet at the mask decoder.  The LLM branch is a chain of library GEMMs
        # at the tensor-core peak whose last waves leave SMs idle; the pixel decoder's memory-bound kernels fill them.
        branch, main = None, None
        if self.overlap_branches and toks[0].is_cuda and torch.cuda.is_current_stream_capturing():
            # (graph capture only: eager launches are host bound, two streams would buy nothing there)
            main = torch.cuda.current_stream(self.device)
            if not hasattr(self, "_branch_stream"):
                self._branch_stream = torch.cuda.Stream(device=self.device, priority=-1 if os.environ.get("PSALM_BRANCH_PRIO") else 0)
            self._branch_stream.wait_stream(main)
            with torch.cuda.stream(self._branch_stream):
                branch = self.pixel_decoder.forward_tokens(toks, sizes)
            for t in toks:
                t.record_stream(self._branch_stream)
        region_feat = None
        if plan.region_pos is not None:
            src_tok = img_tok
            if plan.vp_images is not None:   # DAVIS variant: pooled from the visual-prompt frame's map (llava_phi.py:1665-1670)
                vtoks, vsizes = self.model.vision_tower.forward_tokens(plan.vp_images)
                src_tok = self.model.mm_projector(vtoks[3].view(vtoks[3].shape[0], vsizes[3][0], vsizes[3][1], -1).permute(0, 3, 1, 2))
            region_feat = self._region_features(src_tok, plan)
        embeds = SEQ.materialize_embeds(plan, self.model.embed_tokens, img_tok, self.seg_query, region_feat)
        # (cutting the batch into groups on separate streams so that the prefill GEMMs fill each other's tail waves was
        # measured and is slower: 20.8 ms / 21.9 ms per step of 4 for 2 / 4 groups against 19.5 ms)
        hidden = self.model.phi(embeds, plan.attention_mask if plan.any_padding else None)
        seg_q = F.linear(SEQ.gather_seg_query(plan, hidden), *self.proj["seg_query_projector"])
        SEG_emb = cls_emb = None
        if plan.refer_pool is not None:
            SEG_emb = F.linear(SEQ.pool(plan.refer_pool, hidden), *self.proj["SEG_token_projector"])
        if plan.cls_pool is not None:
            cls_emb = F.linear(SEQ.pool(plan.cls_pool, hidden), *self.proj["class_name_projector"])
        region_emb = None
        if plan.region_pos is not None:     # llava_phi.py:1385-1388: hidden states at the <region> rows -> region_projector
            if "region_projector" not in self.proj:
                raise KeyError("region prompts need region_projector.* in the checkpoint")
            rows = F.linear(SEQ.gather_region_rows(plan, hidden), *self.proj["region_projector"])
            region_emb = list(torch.split(rows, list(plan.region_counts), 0))
        if branch is None:
            mask_features, ms, ms_sizes = self.pixel_decoder.forward_tokens(toks, sizes)
        else:   # join the pixel-decoder branch
            mask_features, ms, ms_sizes = branch
            main.wait_stream(self._branch_stream)
            for t in [mask_features] + list(ms):
                t.record_stream(main)
        out = self.predictor.forward_tokens(ms, ms_sizes, mask_features, sizes[0], seg_q, SEG_emb, cls_emb,
                                            region_embedding_list=region_emb)
        out["mask_size"] = sizes[0]
        if trace is not None:
            trace.update(region_features=region_feat, region_emb=region_emb)
            trace.update(swin=toks, swin_sizes=sizes, img_tok=img_tok, embeds=embeds, hidden=hidden, seg_query=seg_q,
                         SEG_emb=SEG_emb, cls_emb=cls_emb, mask_features=mask_features, ms=ms, ms_sizes=ms_sizes)
        return out

    # ---- CUDA-graph replay of the device-only part -------------------------------------------------
    def _post_device(self, out, image_hw, geoms=None):
        """Device part of the fused post-processing for every image (capturable); None if not applicable.
        geoms: per image (oh, ow, height, width) = un-padded box and output size; None = no crop, output = padded size."""
        Hi, Wi = image_hw
        d = self.size_divisibility
        Hp, Wp = (Hi + d - 1) // d * d, (Wi + d - 1) // d * d
        H4, W4 = out["mask_size"]
        B, Q = out["pred_masks"].shape[:2]
        cls = out["pred_class_name_logits"]
        if not (self.fused_postprocess and Hp >= 2 * H4 and Wp >= 2 * W4 and Q <= 104 and
                (cls is None or cls.shape[-1] - 1 <= 144)):
            return None
        from . import kernels
        thing = PP.thing_tensor(self.is_thing_list, self.device) if (self.panoptic_on and self.instance_on) else None
        pm = out["pred_masks"].view(B, Q, H4, W4)
        sizes, crops = [], []
        for b in range(B):
            oh, ow, height, width = geoms[b] if geoms is not None else (Hp, Wp, Hp, Wp)
            sizes.append((height, width))
            crops.append(None if (oh, ow, height, width) == (Hp, Wp, Hp, Wp) else (Hp, Wp, oh, ow))
        if os.environ.get("PSALM_NO_BATCH_POST"):
            return [PP.fused_device(kernels, pm[b], sizes[b][0], sizes[b][1], cls[b] if cls is not None else None,
                                    out["pred_SEG_logits"][b] if out["pred_SEG_logits"] is not None else None, thing,
                                    self.semantic_on, self.instance_on, self.panoptic_on, self.referring_on,
                                    self.test_topk_per_image, self.object_mask_threshold, crop=crops[b]) for b in range(B)]
        return PP.fused_device_batch(kernels, pm, sizes, cls, out["pred_SEG_logits"], thing, self.semantic_on,
                                     self.instance_on, self.panoptic_on, self.referring_on, self.test_topk_per_image,
                                     self.object_mask_threshold, crops=crops)

    MAX_GRAPHS = 8   # each entry owns static buffers + a private pool (hundreds of MB at 1024^2, B = 4)

    def forward_core_graphed(self, images, plan, lane=0, fuse_post=True):
        # fuse_post: True (no crop / resize), False (task heads outside the graph) or a tuple of per-image geometries
        """Same results as forward_core, replayed from a CUDA graph captured per (image size, prompt
        structure): the ~800 launches of one image become one graph launch (the reference issues them
        one by one from Python, plus ~150 extra tiny launches in its decoder).  `lane` selects an
        independent graph + static buffers so that several images can be in flight on different streams."""
        geoms = fuse_post if isinstance(fuse_post, tuple) else None
        key = (lane, fuse_post, self.seg_task, float(self.object_mask_threshold), tuple(getattr(self, "is_thing_list", None) or ()), tuple(images.shape), str(images.dtype), plan.B,
               plan.T, plan.n_img, plan.any_padding,
               None if plan.cls_pool is None else tuple(plan.cls_pool.shape), plan.refer_pool is not None,
               None if plan.pad_pos is None else int(plan.pad_pos.numel()))
        if not hasattr(self, "_graphs"):
            self._graphs = OrderedDict()
        ent = self._graphs.get(key)
        if ent is not None:
            self._graphs.move_to_end(key)
        tensors = ("tok_ids", "img_pos", "seg_pos", "pad_pos", "attention_mask", "cls_pool", "refer_pool")
        if ent is None:
            import copy
            static_img = images.clone()
            static_plan = copy.copy(plan)
            for n in tensors:
                t = getattr(plan, n)
                setattr(static_plan, n, None if t is None else t.clone())
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            hw = tuple(images.shape[-2:])
            with torch.cuda.stream(side):
                for _ in range(2):
                    o = self.forward_core(static_img, static_plan)
                    o["post"] = self._post_device(o, hw, geoms) if fuse_post else None
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with self._precision_scope(), torch.cuda.graph(g):
                static_out = self._forward_core(static_img, static_plan)
                # task heads' device part in the same graph (only when every image takes the fused path)
                static_out["post"] = self._post_device(static_out, hw, geoms) if fuse_post else None
                static_out["post_geoms"] = geoms
            ent = (g, static_img, static_plan, static_out)
            while len(self._graphs) >= self.MAX_GRAPHS:   # least recently used graph and its static buffers go
                self._graphs.popitem(last=False)
            self._graphs[key] = ent
        g, static_img, static_plan, static_out = ent
        static_img.copy_(images, non_blocking=True)
        for n in tensors:
            t = getattr(plan, n)
            if t is not None:
                getattr(static_plan, n).copy_(t, non_blocking=True)
        g.replay()
        return static_out

    # ---- input staging: overlap the upload of batch k+1 with the compute of batch k ----------------------
    def stage_images(self, images_host):
        """Enqueue the host->device copy of a (pinned) image batch on a dedicated copy stream and return a
        `StagedImages` handle for `eval_seg`.  Two device buffers per (shape, dtype) alternate; a buffer is reused
        only after the pass that consumed it has read it (event recorded by eval_seg)."""
        if not hasattr(self, "_stage"):
            self._stage = dict(stream=torch.cuda.Stream(device=self.device), rings={}, count={})
        st = self._stage
        key = (tuple(images_host.shape), images_host.dtype)
        ring = st["rings"].setdefault(key, [[torch.empty(images_host.shape, dtype=images_host.dtype, device=self.device),
                                             None] for _ in range(2)])
        n = st["count"].get(key, 0)
        st["count"][key] = n + 1
        slot = ring[n % 2]
        with torch.cuda.stream(st["stream"]):
            if slot[1] is not None:
                st["stream"].wait_event(slot[1])
            slot[0].copy_(images_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(st["stream"])
        return StagedImages(slot, ready)

    def make_plan(self, input_ids, attention_mask, image_hw, class_name_ids=None, cls_indices=None,
                  class_name_embedding_indices=None, token_refer_id=None, refer_embedding_indices=None):
        H, W = image_hw
        ps = self.cfg.swin.patch
        h, w = -(-H // ps), -(-W // ps)
        for _ in range(len(self.cfg.swin.depths) - 1):
            h, w = (h + 1) // 2, (w + 1) // 2
        n_img = ((h - 1) // 2 + 1) * ((w - 1) // 2 + 1)   # conv3x3 stride 2 pad 1 of the projector
        return SEQ.build_plan(input_ids, attention_mask, n_img, self.num_queries, class_name_ids, cls_indices,
                              class_name_embedding_indices, token_refer_id, refer_embedding_indices)

    def _cached_plan(self, input_ids, attention_mask, image_hw, class_name_ids, cls_indices,
                     class_name_embedding_indices, token_refer_id, refer_embedding_indices):
        """The sequence plan depends only on the prompt (ids / masks / class-name tables) and the image size;
        evaluation loops reuse one prompt for every image, so the device-resident plan is cached by content."""
        def key_of(t):
            if t is None:
                return None
            if isinstance(t, (list, tuple)):
                return tuple(key_of(x) for x in t)
            t = t.detach().cpu().contiguous()
            return (tuple(t.shape), str(t.dtype), t.numpy().tobytes())
        key = (tuple(image_hw),) + tuple(key_of(t) for t in (input_ids, attention_mask, class_name_ids, cls_indices,
                                                              class_name_embedding_indices, token_refer_id,
                                                              refer_embedding_indices))
        if not hasattr(self, "_plans"):
            self._plans = {}
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 16:
                self._plans.pop(next(iter(self._plans)))
            plan = self.make_plan(input_ids, attention_mask, image_hw, class_name_ids, cls_indices,
                                  class_name_embedding_indices, token_refer_id, refer_embedding_indices).to(self.device)
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def eval_seg(self, input_ids=None, attention_mask=None, past_key_values=None, inputs_embeds=None, labels=None,
                 use_cache=None, output_attentions=None, output_hidden_states=None, images=None, return_dict=None,
                 seg_info=None, class_name_ids=None, class_name_embedding_indices=None, cls_indices=None,
                 token_refer_id=None, refer_embedding_indices=None, is_thing_list=None, region_points=None, vp_images=None):
        if self.panoptic_on:
            assert is_thing_list is not None, "is_thing_list need to be given"   # llava_phi.py:1337-1339
            self.is_thing_list = is_thing_list
        return self.eval_seg_async(region_points=region_points, vp_images=vp_images, input_ids=input_ids, attention_mask=attention_mask, images=images, seg_info=seg_info,
                                   class_name_ids=class_name_ids, class_name_embedding_indices=class_name_embedding_indices,
                                   cls_indices=cls_indices, token_refer_id=token_refer_id,
                                   refer_embedding_indices=refer_embedding_indices, is_thing_list=is_thing_list).result()

    @torch.no_grad()
    def eval_seg_async(self, input_ids=None, attention_mask=None, images=None, seg_info=None, class_name_ids=None,
                       class_name_embedding_indices=None, cls_indices=None, token_refer_id=None,
                       refer_embedding_indices=None, is_thing_list=None, lane=0, region_points=None, vp_images=None):
        """Submit one `eval_seg` call and return a `PendingSeg`; `.result()` gives what `eval_seg` returns.  `lane`
        selects an independent CUDA graph + static output buffers, so that a caller alternating lanes 0 / 1 can finish
        batch k (host merge, read-back) while the device already runs batch k+1.  `region_points`: optional per-sample
        [K,256,2] sample points for <region> prompts (default: drawn like the reference, psalm_b200/region.py)."""
        if self.panoptic_on:
            assert is_thing_list is not None, "is_thing_list need to be given"   # llava_phi.py:1337-1339
            self.is_thing_list = is_thing_list
        staged = images if isinstance(images, StagedImages) else None
        if staged is not None:   # upload already in flight on the copy stream (stage_images)
            torch.cuda.current_stream(self.device).wait_event(staged.ready)
            images_d = staged.tensor
        else:
            # float images are the reference contract (already normalised by the mapper); uint8 images are raw pixel
            # values, normalised on the device (coco_panoptic_mapper.py:161) - 4x fewer bytes over PCIe
            images_d = images.to(self.device, non_blocking=True)
        plan = self._cached_plan(input_ids, attention_mask, images.shape[-2:], class_name_ids, cls_indices,
                                 class_name_embedding_indices, token_refer_id, refer_embedding_indices)
        has_regions = plan.region_pos is not None
        if has_regions:   # llava_phi.py:1346-1349: the regions come with the request (seg_info[i]['instances'].region_masks)
            import copy
            from .region import region_inputs
            pts, img, counts = region_inputs(seg_info, region_points, "region_masks" if vp_images is None else "vp_region_masks")
            assert counts == plan.region_counts, "the munber of <region> tokens and regions needs to be same"   # llava_phi.py:593
            plan = copy.copy(plan)
            plan.region_points, plan.region_image = pts.to(self.device), img.to(self.device)
            if vp_images is not None:
                plan.vp_images = vp_images.to(self.device)
        fused, boxes = self._fused_applies(images.shape[-2:], seg_info)
        if self.use_cuda_graph and not has_regions:   # the number of regions varies per request: eager launches
            out = self.forward_core_graphed(images_d, plan, lane=lane, fuse_post=fused)
        else:
            out = self.forward_core(images_d, plan)
        cur = torch.cuda.current_stream(self.device)
        if staged is not None:   # the staging buffer may be overwritten once this pass has read it
            staged.slot[1] = torch.cuda.Event()
            staged.slot[1].record(cur)
        hostvecs = None
        if out.get("post") is not None:   # the integers of the host merge: device -> pinned memory, behind the pass
            if not hasattr(self, "_hostvec_pins"):
                self._hostvec_pins = {}
            hostvecs = []
            for b, d in enumerate(out["post"]):
                hv = d["hostvec"]
                if hv is None:
                    hostvecs.append(None)
                    continue
                k = (lane, b, tuple(hv.shape), hv.dtype)
                pin = self._hostvec_pins.get(k)
                if pin is None:
                    pin = self._hostvec_pins[k] = torch.empty(hv.shape, dtype=hv.dtype, pin_memory=True)
                pin.copy_(hv, non_blocking=True)
                hostvecs.append(pin)
        done = torch.cuda.Event()
        done.record(cur)
        return PendingSeg(self, out, tuple(images.shape[-2:]), seg_info, boxes, done, hostvecs,
                          getattr(self, "is_thing_list", None), (self.object_mask_threshold, self.overlap_threshold))

    def _fused_applies(self, image_hw, seg_info):
        """(fuse_post, boxes): fuse_post is True when every image takes the fused task-head kernel without crop / resize,
        a tuple of per-image (oh, ow, height, width) when the composed up-sample -> crop -> resize kernel applies to all of
        them (the reference's mapper flow: padded 1024^2 input, original-size output), False otherwise (the graph is then
        captured without the task heads instead of running them for nothing)."""
        Hi, Wi = image_hw
        d = self.size_divisibility
        Hp, Wp = (Hi + d - 1) // d * d, (Wi + d - 1) // d * d
        boxes = [PP.unpadded_box(info["padding_mask"]) for info in seg_info]
        geoms = tuple((box[0], box[1], info.get("height", Hi), info.get("width", Wi)) for info, box in zip(seg_info, boxes))
        if all(g == (Hp, Wp, Hp, Wp) for g in geoms):
            return True, boxes
        if not (self.fused_postprocess and self.sem_seg_postprocess_before_inference):
            return False, boxes
        from . import kernels
        ps = self.cfg.swin.patch
        H4, W4 = -(-Hi // ps), -(-Wi // ps)
        ncls = 144
        ok = all(kernels.postproc_crop_supported(self.num_queries, H4, W4, Hp, Wp, g[0], g[1], g[2], g[3], ncls) or
                 g == (Hp, Wp, Hp, Wp) for g in geoms) and self.num_queries <= 104 and Hp >= 2 * H4 and Wp >= 2 * W4
        return (geoms if ok else False), boxes

    @torch.no_grad()
    def post_process(self, out, image_hw, seg_info, boxes=None, hostvecs=None):
        """llava_phi.py:1395-1472 for EVERY image of the batch.  `boxes`: un-padded (h, w) per image when the
        caller already derived them from the padding masks; `hostvecs`: the fused heads' host vectors when they were
        already copied to (pinned) host memory."""
        with self._precision_scope():
            return self._post_process(out, image_hw, seg_info, boxes, hostvecs)

    def _post_process(self, out, image_hw, seg_info, boxes=None, hostvecs=None):
        Hi, Wi = image_hw
        d = self.size_divisibility
        Hp, Wp = (Hi + d - 1) // d * d, (Wi + d - 1) // d * d     # ImageList.from_tensors(images, 32), :1400
        H4, W4 = out["mask_size"]
        B, Q = out["pred_masks"].shape[:2]
        pm = out["pred_masks"].view(B, Q, H4, W4)
        mask_pred = None
        results = []
        for b in range(B):
            info = seg_info[b]
            height, width = info.get("height", Hi), info.get("width", Wi)
            oh, ow = boxes[b] if boxes is not None else PP.unpadded_box(info["padding_mask"])
            cls_b = out["pred_class_name_logits"][b] if out["pred_class_name_logits"] is not None else None
            seg_b = out["pred_SEG_logits"][b] if out["pred_SEG_logits"] is not None else None
            trivial = (oh, ow) == (Hp, Wp) and (height, width) == (Hp, Wp)
            pg = out.get("post_geoms")
            if out.get("post") is not None and ((pg is None and trivial) or (pg is not None and pg[b] == (oh, ow, height, width))):
                results.append(PP.fused_host(out["post"][b], getattr(self, "is_thing_list", None), self.overlap_threshold,
                                             host=None if hostvecs is None or hostvecs[b] is None else hostvecs[b].numpy()))
                continue
            if (not trivial and self.fused_postprocess and self.sem_seg_postprocess_before_inference and Q <= 104 and
                    Hp >= 2 * H4 and Wp >= 2 * W4 and (cls_b is None or cls_b.shape[-1] - 1 <= 144)):
                from . import kernels
                if kernels.postproc_crop_supported(Q, H4, W4, Hp, Wp, oh, ow, height, width, 144):
                    results.append(PP.fused_postprocess(
                        kernels, pm[b], height, width, cls_b, seg_b, getattr(self, "is_thing_list", None), self.semantic_on,
                        self.instance_on, self.panoptic_on, self.referring_on, self.test_topk_per_image,
                        self.object_mask_threshold, self.overlap_threshold, crop=(Hp, Wp, oh, ow)))
                    continue
            if self.fused_postprocess and trivial and Hp >= 2 * H4 and Wp >= 2 * W4 and Q <= 104 and \
                    (cls_b is None or cls_b.shape[-1] - 1 <= 144):
                from . import kernels
                results.append(PP.fused_postprocess(
                    kernels, pm[b], Hp, Wp, cls_b, seg_b, getattr(self, "is_thing_list", None), self.semantic_on,
                    self.instance_on, self.panoptic_on, self.referring_on, self.test_topk_per_image,
                    self.object_mask_threshold, self.overlap_threshold))
                continue
            if mask_pred is None:
                mask_pred = F.interpolate(pm.float(), size=(Hp, Wp), mode="bilinear", align_corners=False)
            mp = mask_pred[b]
            r = {}
            if self.sem_seg_postprocess_before_inference:
                mp = PP.sem_seg_postprocess(mp, (oh, ow), height, width)
            cls = out["pred_class_name_logits"][b].float() if out["pred_class_name_logits"] is not None else None
            mp = mp.contiguous()
            sig = mp.sigmoid()   # shared by the task heads below (the reference recomputes it per head)
            tf32 = self.dtype != torch.float32
            if self.semantic_on:
                sem = PP.semantic_inference(cls, mp, sig, tf32)
                if not self.sem_seg_postprocess_before_inference:
                    sem = PP.sem_seg_postprocess(sem, (oh, ow), height, width)
                r["sem_seg"] = sem
            if self.instance_on:
                r["instances"] = PP.instance_inference(cls, mp, self.test_topk_per_image,
                                                       getattr(self, "is_thing_list", None), self.panoptic_on, sig)
            if self.panoptic_on:
                r["panoptic_seg"] = PP.panoptic_inference(cls, mp, self.is_thing_list,
                                                          self.object_mask_threshold,
                                                          self.overlap_threshold, sig)
            if self.referring_on:
                r["instances"] = PP.seg_instance_inference(out["pred_SEG_logits"][b].float(), mp,
                                                           self.test_topk_per_image, sig)
            if self.region_on:   # llava_phi.py:1457-1466
                r["instances"] = PP.region_inference(out["pred_region_logits"][b].float(), mp, sig)
                gt = info["instances"].gt_masks
                gt = gt.tensor if hasattr(gt, "tensor") else gt
                r["gt"] = PP.sem_seg_postprocess(gt.to(mp.device).float(), (oh, ow), height, width)
            results.append(r)
        return results


class PSALMForDAVISEval(PSALM):
    """Video-object-segmentation variant (llava_phi.py:1477-2012, builder.py:47 'psalm_video'): the <region> prompts of
    the current frame are pooled from a VISUAL-PROMPT frame (`vp_images`, usually the first frame of the clip) with
    `seg_info[i]['instances'].vp_region_masks`; everything after the sequence splice is PSALM.eval_seg.  The reference's
    `eval_seg` and `eval_video` of this class run the same computation."""

    def eval_seg(self, *args, vp_images=None, **kw):
        if vp_images is None:
            raise ValueError("PSALMForDAVISEval needs vp_images (the visual-prompt frames, llava_phi.py:1497)")
        return super().eval_seg(*args, vp_images=vp_images, **kw)

    eval_video = eval_seg
