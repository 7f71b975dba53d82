"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU fp32 restatement of the CLIP ViT image-encoder guidance
of BASELINE configs[2] / SURVEY §8c.

The reference repository contains NO implementation, call site or test of this path (only prose, model-card.md:45-48);
the upstream module would be openai/CLIP `clip.load("ViT-B/16")`, which is neither vendored nor installed.  The spec
is therefore the one SURVEY §8c states, and the oracle is PINNED against `transformers.CLIPVisionModelWithProjection`
(hidden_act="quick_gelu", the published CLIP architecture; transformers 5.5.0 in this image) by oracle/make_golden.py:
    x in [-1,1] -> (x+1)/2 -> bilinear resize to the encoder's image size (align_corners=False)
      -> (. - mean)/std (CLIP constants) -> ViT -> L2-normalise -> s * <e_img, e_txt>,  gradient w.r.t. x.
State-dict keys are the transformers ones (vision_model.*, visual_projection.weight)."""
import torch as th
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess(x: th.Tensor, image_size: int) -> th.Tensor:
    y = (x + 1.0) / 2.0
    if y.shape[-1] != image_size or y.shape[-2] != image_size:
        y = F.interpolate(y, size=(image_size, image_size), mode="bilinear", align_corners=False)
    mean = th.tensor(CLIP_MEAN, dtype=y.dtype).view(1, 3, 1, 1)
    std = th.tensor(CLIP_STD, dtype=y.dtype).view(1, 3, 1, 1)
    return (y - mean) / std


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def image_embed(sd, pixels: th.Tensor, *, heads: int, layers: int, patch: int) -> th.Tensor:
    """transformers CLIPVisionModelWithProjection.forward: patch conv (no bias) + class token + learned positions,
    pre-LN, `layers` pre-norm blocks (MHA with q scaled by d^-1/2, QuickGELU MLP), post-LN of the class token,
    linear projection (no bias)."""
    p = "vision_model."
    h = F.conv2d(pixels, sd[p + "embeddings.patch_embedding.weight"], None, stride=patch)  # [B,H,gh,gw]
    bsz, hid = h.shape[0], h.shape[1]
    h = h.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].view(1, 1, hid).expand(bsz, 1, hid)
    h = th.cat([cls, h], 1) + sd[p + "embeddings.position_embedding.weight"].unsqueeze(0)
    h = _ln(h, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"])
    d = hid // heads
    for i in range(layers):
        q = f"{p}encoder.layers.{i}."
        y = _ln(h, sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"])
        t = y.shape[1]
        qq = F.linear(y, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"]) * d ** -0.5
        kk = F.linear(y, sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"])
        vv = F.linear(y, sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"])
        qq, kk, vv = (z.view(bsz, t, heads, d).transpose(1, 2) for z in (qq, kk, vv))
        a = th.softmax(qq @ kk.transpose(-1, -2), dim=-1) @ vv
        a = a.transpose(1, 2).reshape(bsz, t, hid)
        h = h + F.linear(a, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        y = _ln(h, sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"])
        y = F.linear(y, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])
        y = y * th.sigmoid(1.702 * y)
        h = h + F.linear(y, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    pooled = _ln(h[:, 0], sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"])
    return F.linear(pooled, sd["visual_projection.weight"])


def similarity(sd, x, text, scale, *, image_size, heads, layers, patch):
    e = image_embed(sd, preprocess(x, image_size), heads=heads, layers=layers, patch=patch)
    e = e / e.norm(dim=-1, keepdim=True)
    return scale * (e * text).sum(-1)


def guidance(sd, x, text, scale, **kw):
    """grad_x sum_b s * <e_img(x_b), e_txt_b>"""
    with th.enable_grad():
        xin = x.detach().requires_grad_(True)
        sim = similarity(sd, xin, text, scale, **kw)
        return th.autograd.grad(sim.sum(), xin)[0]
