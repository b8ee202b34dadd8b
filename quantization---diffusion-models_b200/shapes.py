"""Linear-layer inventories of the BASELINE.json configs: module names are the
diffusers names the reference's name-based logic sees (models/StableDiffusion1_x.py:121-137).
Each entry: (name, M, N, K, count) for one UNet / MMDiT forward at the config's effective batch."""

CTX_SD15, CTX_SDXL = 768, 2048
TXT_TOKENS = 77


def group_for(k, group_size=128):
    """quantize/fake_quant.py:34-37 fallback: K=320 -> 64."""
    while group_size > 0 and k % group_size:
        group_size -= 32
    return group_size


def basic_transformer_block(C, ctx, m_img, m_txt):
    """diffusers BasicTransformerBlock linears: (name, M, N, K)."""
    return [
        ("attn1.to_q", m_img, C, C), ("attn1.to_k", m_img, C, C), ("attn1.to_v", m_img, C, C),
        ("attn1.to_out.0", m_img, C, C),
        ("attn2.to_q", m_img, C, C), ("attn2.to_k", m_txt, C, ctx), ("attn2.to_v", m_txt, C, ctx),
        ("attn2.to_out.0", m_img, C, C),
        ("ff.net.0.proj", m_img, 8 * C, C), ("ff.net.2", m_img, C, 4 * C),
    ]


def _merge(layers):
    out = {}
    for name, m, n, k in layers:
        key = (name, m, n, k)
        out[key] = out.get(key, 0) + 1
    return [(name, m, n, k, c) for (name, m, n, k), c in out.items()]


def sd15_unet_linears(batch=8, cfg=True):
    """SD1.5 UNet at 512x512 (latent 64x64): blocks 320x5 @4096 tok, 640x5 @1024, 1280x5 @256, 1280x1 @64."""
    b = batch * (2 if cfg else 1)
    layers = []
    for C, nblk, tok in ((320, 5, 4096), (640, 5, 1024), (1280, 5, 256), (1280, 1, 64)):
        for _ in range(nblk):
            layers += [(f"C{C}.{n}", m, nn_, k) for n, m, nn_, k in basic_transformer_block(C, CTX_SD15, b * tok, b * TXT_TOKENS)]
    layers += [("time_embedding.linear_1", b, 1280, 320), ("time_embedding.linear_2", b, 1280, 1280)]
    for cout, cnt in ((320, 7), (640, 6), (1280, 9)):   # 22 resnet time_emb_proj
        layers += [(f"time_emb_proj.{cout}", b, cout, 1280)] * cnt
    return _merge(layers)


def sdxl_unet_linears(batch=4, cfg=True):
    """SDXL UNet at 1024x1024 (latent 128x128): 640 x 10 blocks @4096 tok, 1280 x 60 blocks @1024 tok."""
    b = batch * (2 if cfg else 1)
    layers = []
    for C, nblk, tok, nproj in ((640, 10, 4096, 5), (1280, 60, 1024, 6)):
        for _ in range(nblk):
            layers += [(f"C{C}.{n}", m, nn_, k) for n, m, nn_, k in basic_transformer_block(C, CTX_SDXL, b * tok, b * TXT_TOKENS)]
        layers += [(f"C{C}.proj_in", b * tok, C, C), (f"C{C}.proj_out", b * tok, C, C)] * nproj
    layers += [("time_embedding.linear_1", b, 1280, 320), ("time_embedding.linear_2", b, 1280, 1280),
               ("add_embedding.linear_1", b, 1280, 2816), ("add_embedding.linear_2", b, 1280, 1280)]
    for cout, cnt in ((320, 4), (640, 6), (1280, 7)):
        layers += [(f"time_emb_proj.{cout}", b, cout, 1280)] * cnt
    return _merge(layers)


def sd35_mmdit_linears(batch=1):
    """SD3.5-Large MMDiT at 1024x1024: 38 joint blocks, D=2432, FF 9728, 4096 image + 333 text tokens."""
    D, FF, mi, mt = 2432, 9728, batch * 4096, batch * 333
    layers = []
    for blk in range(38):
        last = blk == 37
        layers += [("norm1.linear", batch, 6 * D, D), ("norm1_context.linear", batch, (2 if last else 6) * D, D)]
        layers += [(f"attn.{n}", mi, D, D) for n in ("to_q", "to_k", "to_v", "to_out.0")]
        layers += [(f"attn.{n}", mt, D, D) for n in ("add_q_proj", "add_k_proj", "add_v_proj")]
        layers += [("ff.net.0.proj", mi, FF, D), ("ff.net.2", mi, D, FF)]
        if not last:
            layers += [("attn.to_add_out", mt, D, D), ("ff_context.net.0.proj", mt, FF, D), ("ff_context.net.2", mt, D, FF)]
    layers += [("context_embedder", mt, D, 4096), ("time_text_embed.t1", batch, D, 256), ("time_text_embed.t2", batch, D, D),
               ("time_text_embed.p1", batch, D, 2048), ("time_text_embed.p2", batch, D, D),
               ("norm_out.linear", batch, 2 * D, D), ("proj_out", mi, 64, D)]
    return _merge(layers)


def total_flops(layers):
    return sum(2.0 * m * n * k * c for _, m, n, k, c in layers)


def gemm_bytes_w4a16(m, n, k, g):
    """algorithmic bytes of one W4A16 call."""
    return 2 * m * k + 0.5 * k * n + 2.5 * (k // g) * n + 2 * m * n


def gemm_bytes_w8a8(m, n, k):
    return m * k + 4 * m + k * n + 4 * n + 2 * m * n
