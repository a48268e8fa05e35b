"""Chat / decode path: PSALM.generate = prefill + paged-KV decode + lm_head, against the oracle
recomputing the FULL causal prefill for every new token (no cache) - host logic on CPU with emulated kernels."""
import torch

import emu
from oracle import psalm_oracle as O
from psalm_b200 import synth
from psalm_b200.layout import PhiConfig, PsalmConfig

SMALL = PsalmConfig(phi=PhiConfig(hidden=256, layers=2, heads=4, inter=1024, vocab=1000))
SMALL_O = dict(hidden=256, layers=2, heads=4, inter=1024, eps=1e-5, theta=10000.0, rotary_frac=0.5)


def oracle_generate(sd, input_ids, images, n_new):
    """Greedy decoding the slow way: re-run the whole sequence for every token."""
    with torch.no_grad():
        img_tok = None
        if images is not None:
            feats = O.swin_forward(sd, "model.vision_tower.", images)
            img_tok = O.projector_forward(sd, "model.mm_projector.", feats[3])
        emb = sd["model.embed_tokens.weight"]
        B = input_ids.shape[0]
        seqs = []
        for b in range(B):
            row = input_ids[b]
            pos = (row == O.IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
            if pos:
                p = pos[0]
                seqs.append(torch.cat([emb[row[:p]], img_tok[b], emb[row[p + 1:]]], 0))
            else:
                seqs.append(emb[row])
        x = torch.stack(seqs, 0)
        toks = []
        for _ in range(n_new):
            am = torch.ones(x.shape[:2], dtype=torch.bool)
            h = O.phi_forward(sd, "model.", x, am, SMALL_O)
            logits = h[:, -1] @ sd["lm_head.weight"].t()
            t = logits.argmax(-1)
            toks.append(t)
            x = torch.cat([x, emb[t][:, None]], 1)
        return torch.stack(toks, 1)


def test_generate_matches_full_recompute(monkeypatch):
    from psalm_b200.psalm import PSALM
    emu.install(monkeypatch)

    class _EmuPSALM(PSALM):
        @staticmethod
        def _check_runtime(device):
            pass
    sd = synth.synth_state_dict(SMALL, seed=3, include_lm_head=True)
    m = _EmuPSALM(sd, SMALL, torch.float32, "cpu", "panoptic")
    g = torch.Generator().manual_seed(0)
    # text-only prompt, batch 2; and a prompt with an image
    ids = torch.randint(5, 1000, (2, 21), generator=g)
    got = m.generate(ids, max_new_tokens=6)
    assert torch.equal(got, oracle_generate(sd, ids, None, 6))
    ids2 = torch.cat([ids[:1, :7], torch.tensor([[-200]]), ids[:1, 7:12]], 1)
    img = torch.randn(1, 3, 64, 96, generator=g)
    got2 = m.generate(ids2, images=img, max_new_tokens=5, page_size=4)
    assert torch.equal(got2, oracle_generate(sd, ids2, img, 5))
    # eos stops early
    eos = int(got[0, 2])
    short = m.generate(ids[:1], max_new_tokens=6, eos_token_id=eos)
    assert short.shape[1] <= 3 and int(short[0, -1]) == eos
