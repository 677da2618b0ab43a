"""Conditioning encoders against the implementation the reference itself calls: the Hugging
Face ``transformers`` CLIP / ViT-MAE classes, EXECUTED here (fp32, same random weights) — transformers is installed in
this image, so unlike the diffusers blocks this parity is pinned to real third-party code, not to a restatement.
Tolerance: rel-L2 <= 2e-2 on the tensors the pipeline reads (bf16 activations through up to 32 layers)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")
bf16, f32 = torch.bfloat16, torch.float32
DEV = "cuda"


def _hf_text(cfg, with_proj):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    c = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                       num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                       max_position_embeddings=cfg.max_position_embeddings, hidden_act=cfg.hidden_act,
                       layer_norm_eps=cfg.layer_norm_eps, projection_dim=max(cfg.projection_dim, 1),
                       eos_token_id=cfg.eos_token_id, bos_token_id=0, pad_token_id=1)
    torch.manual_seed(0)
    m = (CLIPTextModelWithProjection if with_proj else CLIPTextModel)(c)
    _spread(m)
    return m.to(DEV).eval()


def _spread(m):
    """HF's default init (std 0.02) makes every layer nearly an identity; give the weights O(1/sqrt(fan_in)) scale so a
    wrong head split / activation / mask would actually show."""
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and "embedding" not in n and "cls_token" not in n and "position" not in n:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (0.7 * fan_in ** -0.5))
            elif n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            p.copy_(p.to(bf16).float())                      # both sides see bf16-representable weights


def _ids(B, L, vocab, eos, legacy):
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, vocab - 2, (B, L), generator=g)
    ends = torch.randint(3, L - 1, (B,), generator=g)
    for b in range(B):
        ids[b, ends[b]] = (vocab - 1) if legacy else eos     # legacy (eos_token_id == 2): EOS = the largest id
        ids[b, ends[b] + 1:] = 1 if not legacy else 0
    return ids


@pytest.mark.parametrize("name", ["tiny_quick_gelu", "tiny_proj_eos", "clip_l", "openclip_bigg"])
def test_clip_text_encoder_matches_transformers(name):
    import diffsensei_b200 as ds
    if name == "tiny_quick_gelu":
        cfg, B = ds.EncoderConfig(128, 3, 2, 256, "quick_gelu", vocab_size=1000, max_position_embeddings=77), 3
    elif name == "tiny_proj_eos":
        cfg, B = ds.EncoderConfig(192, 2, 3, 384, "gelu", vocab_size=1000, max_position_embeddings=77, projection_dim=96,
                                  eos_token_id=999), 2
    elif name == "clip_l":
        cfg, B = ds.CLIP_L_TEXT, 2
    else:
        cfg, B = ds.OPENCLIP_BIGG_TEXT, 2
    hf = _hf_text(cfg, cfg.projection_dim > 0)
    eng = ds.ClipTextEncoderEngine(cfg, DEV)
    eng.load_state_dict(hf.state_dict())
    ids = _ids(B, 77, cfg.vocab_size, cfg.eos_token_id, cfg.eos_token_id == 2)
    with torch.no_grad():
        want = hf(ids.to(DEV), output_hidden_states=True)
    got = eng(ids, output_hidden_states=True)
    e2 = rel_l2(got.hidden_states[-2].float(), want.hidden_states[-2])
    el = rel_l2(got.last_hidden_state.float(), want.last_hidden_state)
    print(f"{name}: hidden_states[-2] rel-L2 {e2:.3e}, last_hidden_state {el:.3e}")
    assert got.hidden_states[-2].shape == (B, 77, cfg.hidden_size)
    assert e2 < 2e-2 and el < 2e-2
    if cfg.projection_dim > 0:
        assert rel_l2(got.text_embeds.float(), want.text_embeds) < 2e-2 and got[0] is got.text_embeds
    else:
        assert rel_l2(got.pooler_output.float(), want.pooler_output) < 2e-2


@pytest.mark.parametrize("name", ["tiny_d80", "vit_h"])
def test_clip_vision_encoder_matches_transformers(name):
    import diffsensei_b200 as ds
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    if name == "tiny_d80":
        cfg, B = ds.EncoderConfig(160, 3, 2, 320, "gelu", image_size=56, patch_size=14), 3      # head_dim 80, 17 tokens
    else:
        cfg, B = ds.CLIP_VIT_H, 2
    c = CLIPVisionConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                         image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act=cfg.hidden_act,
                         layer_norm_eps=cfg.layer_norm_eps, projection_dim=64)
    torch.manual_seed(0)
    hf = CLIPVisionModelWithProjection(c)
    _spread(hf)
    hf = hf.to(DEV).eval()
    eng = ds.ClipVisionEncoderEngine(cfg, DEV)
    eng.load_state_dict(hf.state_dict())
    px = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(3))
    px = px.to(bf16).float()
    with torch.no_grad():
        want = hf(px.to(DEV), output_hidden_states=True).hidden_states[-2]
    got = eng(px, output_hidden_states=True).hidden_states[-2]
    err = rel_l2(got.float(), want)
    print(f"{name}: hidden_states[-2] {tuple(got.shape)} rel-L2 {err:.3e}")
    n_tok = (cfg.image_size // cfg.patch_size) ** 2 + 1
    assert got.shape == (B, n_tok, cfg.hidden_size) and err < 2e-2


@pytest.mark.parametrize("name", ["tiny", "magi_base"])
def test_vit_mae_encoder_matches_transformers(name):
    import diffsensei_b200 as ds
    from transformers import ViTMAEConfig, ViTMAEModel
    cfg, B = (ds.EncoderConfig(128, 2, 2, 256, "gelu", 1e-12, image_size=64, patch_size=16), 3) if name == "tiny" \
        else (ds.MAGI_VIT_MAE, 2)
    c = ViTMAEConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act=cfg.hidden_act,
                     layer_norm_eps=cfg.layer_norm_eps, mask_ratio=0.0)
    torch.manual_seed(0)
    hf = ViTMAEModel(c)
    _spread(hf)
    hf = hf.to(DEV).eval()
    eng = ds.VitMaeEncoderEngine(cfg, DEV)
    eng.load_state_dict(hf.state_dict())
    px = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(4)).to(bf16).float()
    with torch.no_grad():
        want = hf(px.to(DEV)).last_hidden_state[:, 0]       # what the pipeline reads (:128); shuffle-invariant
    got = eng(px).last_hidden_state[:, 0]
    err = rel_l2(got.float(), want)
    print(f"{name}: CLS embedding rel-L2 {err:.3e}")
    assert got.shape == (B, cfg.hidden_size) and err < 2e-2


def test_attention_small_and_embed_kernels():
    from diffsensei_b200 import ops
    g = torch.Generator().manual_seed(6)
    for (B, N, heads, d, causal) in ((2, 77, 3, 64, True), (3, 257, 2, 80, False), (1, 197, 4, 64, False), (2, 17, 2, 128, True),
                                     (1, 320, 1, 8, False)):
        qkv = (torch.randn(B, N, 3 * heads * d, generator=g) * 0.8).to(bf16)
        q, k, v = (t.float().view(B, N, heads, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        want = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B, N, -1)
        got = ops.attention_small(qkv.to(DEV), heads, causal).float().cpu()
        assert rel_l2(got, want) < 6e-3, (B, N, heads, d, causal)
    tok, pos = torch.randn(50, 64, generator=g).to(bf16), torch.randn(77, 64, generator=g).to(bf16)
    ids = torch.randint(0, 50, (3, 20), generator=g, dtype=torch.int32)
    got = ops.embed_tokens(ids.to(DEV), tok.to(DEV), pos.to(DEV)).float().cpu()
    assert torch.equal(got, (tok.float()[ids.long()] + pos.float()[:20]).to(bf16).float())


def test_pipeline_runs_from_token_ids_and_pixel_values():
    """__call__ with the encoders registered: token ids + pixel values in, images out (tiny configs end to end)."""
    import dataclasses
    import diffsensei_b200 as ds
    from diffsensei_b200.weights import (random_state_dict, resampler_param_shapes, unet_param_shapes,
                                         vae_decoder_param_shapes)
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection, CLIPVisionConfig, \
        CLIPVisionModelWithProjection, ViTMAEConfig, ViTMAEModel
    torch.manual_seed(0)
    # text encoders whose hidden sizes add up to TINY's cross_attention_dim (128) and whose projection matches pooled_text_dim (96)
    t1 = ds.EncoderConfig(64, 2, 1, 128, "quick_gelu", vocab_size=500, max_position_embeddings=77)
    t2 = ds.EncoderConfig(64, 2, 1, 128, "gelu", vocab_size=500, max_position_embeddings=77, projection_dim=96)
    mk = lambda c, proj: (CLIPTextModelWithProjection if proj else CLIPTextModel)(CLIPTextConfig(
        vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
        num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads, max_position_embeddings=77,
        hidden_act=c.hidden_act, projection_dim=max(c.projection_dim, 1), eos_token_id=2))
    e1, e2 = ds.ClipTextEncoderEngine(t1, DEV), ds.ClipTextEncoderEngine(t2, DEV)
    e1.load_state_dict(mk(t1, False).state_dict())
    e2.load_state_dict(mk(t2, True).state_dict())
    rc = ds.RESAMPLER_TINY                                                   # embedding_dim 64, magi 32, 33 tokens
    vcfg = ds.EncoderConfig(64, 2, 1, 128, "gelu", image_size=64, patch_size=16)     # 16 patches + CLS = 17 tokens
    mcfg = ds.EncoderConfig(32, 2, 1, 64, "gelu", 1e-12, image_size=64, patch_size=16)
    ve = ds.ClipVisionEncoderEngine(vcfg, DEV)
    ve.load_state_dict(CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=1, image_size=64, patch_size=16,
        projection_dim=16)).state_dict())
    me = ds.VitMaeEncoderEngine(mcfg, DEV)
    me.load_state_dict(ViTMAEModel(ViTMAEConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                                num_attention_heads=1, image_size=64, patch_size=16, mask_ratio=0.0)).state_dict())
    unet = ds.UNetMangaEngine(ds.TINY, DEV)
    unet.load_state_dict(random_state_dict(unet_param_shapes(ds.TINY), 0, DEV))
    res = ds.ResamplerEngine(**dataclasses.asdict(rc), device=DEV)
    res.load_state_dict(random_state_dict(resampler_param_shapes(rc), 1, DEV))
    vae = ds.VaeDecoderEngine(ds.TINY_VAE, DEV)
    vae.load_state_dict(random_state_dict(vae_decoder_param_shapes(ds.TINY_VAE), 2, DEV))
    pipe = ds.DiffSenseiPipeline(unet, vae=vae, text_encoder=e1, text_encoder_2=e2, image_encoder=ve)
    pipe.register_manga_modules(me, res)
    ids = _ids(1, 77, 500, 2, True)
    g = torch.Generator().manual_seed(9)
    out = pipe(prompt="ignored: ids given", height=128, width=128, num_inference_steps=2, guidance_scale=7.5,
               num_samples=1, generator=torch.Generator().manual_seed(0), ip_bbox=[[.1, .1, .6, .9], [.5, .2, .95, .9]],
               ip_scale=0.6, dialog_bbox=[[.05, .05, .3, .2]], prompt_input_ids=ids, prompt_input_ids_2=ids,
               clip_pixel_values=torch.randn(2, 3, 64, 64, generator=g), magi_pixel_values=torch.randn(2, 3, 64, 64, generator=g),
               output_type="pt")
    assert out.images.shape == (1, 3, 128, 128) and torch.isfinite(out.images).all()
