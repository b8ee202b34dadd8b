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
    return [e[:4] for e in _transformer_block_inputs("", C, ctx, m_img, m_txt)]


def _transformer_block_inputs(blk, C, ctx, m_img, m_txt):
    """(name, M, N, K, input): `input` names the tensor the Linear reads -- Linears with the same input can run as one
    launch on the N-concatenated packed weight (fused_utils.fuse_linears; utils/fused_utils.py:87-96)."""
    return [
        ("attn1.to_q", m_img, C, C, blk + ".norm1"), ("attn1.to_k", m_img, C, C, blk + ".norm1"), ("attn1.to_v", m_img, C, C, blk + ".norm1"),
        ("attn1.to_out.0", m_img, C, C, blk + ".attn1"),
        ("attn2.to_q", m_img, C, C, blk + ".norm2"), ("attn2.to_k", m_txt, C, ctx, "context"), ("attn2.to_v", m_txt, C, ctx, "context"),
        ("attn2.to_out.0", m_img, C, C, blk + ".attn2"),
        ("ff.net.0.proj", m_img, 8 * C, C, blk + ".norm3"), ("ff.net.2", m_img, C, 4 * C, blk + ".ff"),
    ]


def _merge(layers):
    out = {}
    for name, m, n, k in layers:
        key = (name, m, n, k)
        out[key] = out.get(key, 0) + 1
    return [(name, m, n, k, c) for (name, m, n, k), c in out.items()]


_FUSED_NAMES = {"context": "attn2.to_kv(all blocks)", "temb": "time_emb_proj(all resnets)", "emb": "adaln.linear(all blocks)"}


def _fuse(layers):
    """[(name, M, N, K, input)] -> [(name, M, N_total, K, count, parts)]: members with one input (and M, K) become one entry
    whose `parts` are the members' N in call order; FLOPs are unchanged."""
    groups, order = {}, []
    for name, m, n, k, inp in layers:
        key = (inp, m, k)
        if key not in groups:
            groups[key] = []
            order.append(key)
        groups[key].append((name, n))
    out = {}
    for key in order:
        inp, m, k = key
        members = groups[key]
        parts = tuple(n for _, n in members)
        if len(members) == 1:
            name = members[0][0]
        elif inp in _FUSED_NAMES:
            name = _FUSED_NAMES[inp]
        else:   # q / k / v of one attention: "C320.attn1.to_q" ... -> "C320.attn1.to_qkv"
            first = members[0][0]
            name = first[:first.rfind("_") + 1] + "".join(nm[nm.rfind("_") + 1:].replace("proj", "") for nm, _ in members)
            name = name if not first.endswith("_proj") else first[:first.rfind("add_")] + "add_qkv_proj"
        e = (name, m, sum(parts), k, parts)
        out[e] = out.get(e, 0) + 1
    return [(name, m, n, k, c, parts) for (name, m, n, k, parts), c in out.items()]


def _sd15(batch, cfg):
    b = batch * (2 if cfg else 1)
    layers, i = [], 0
    for C, nblk, tok in ((320, 5, 4096), (640, 5, 1024), (1280, 5, 256), (1280, 1, 64)):
        for _ in range(nblk):
            layers += [(f"C{C}.{n}", m, nn_, k, inp) for n, m, nn_, k, inp in
                       _transformer_block_inputs(f"b{i}", C, CTX_SD15, b * tok, b * TXT_TOKENS)]
            i += 1
    layers += [("time_embedding.linear_1", b, 1280, 320, "t_sin"), ("time_embedding.linear_2", b, 1280, 1280, "t_l1")]
    for cout, cnt in ((320, 7), (640, 6), (1280, 9)):   # 22 resnet time_emb_proj
        layers += [(f"time_emb_proj.{cout}", b, cout, 1280, "temb")] * cnt
    return layers


def sd15_unet_linears(batch=8, cfg=True):
    """SD1.5 UNet at 512x512 (latent 64x64): blocks 320x5 @4096 tok, 640x5 @1024, 1280x5 @256, 1280x1 @64."""
    return _merge([e[:4] for e in _sd15(batch, cfg)])


def _sdxl(batch, cfg):
    b = batch * (2 if cfg else 1)
    layers, i = [], 0
    for C, nblk, tok, nproj in ((640, 10, 4096, 5), (1280, 60, 1024, 6)):
        for _ in range(nblk):
            layers += [(f"C{C}.{n}", m, nn_, k, inp) for n, m, nn_, k, inp in
                       _transformer_block_inputs(f"b{i}", C, CTX_SDXL, b * tok, b * TXT_TOKENS)]
            i += 1
        for j in range(nproj):
            layers += [(f"C{C}.proj_in", b * tok, C, C, f"pi{C}.{j}"), (f"C{C}.proj_out", b * tok, C, C, f"po{C}.{j}")]
    layers += [("time_embedding.linear_1", b, 1280, 320, "t_sin"), ("time_embedding.linear_2", b, 1280, 1280, "t_l1"),
               ("add_embedding.linear_1", b, 1280, 2816, "a_in"), ("add_embedding.linear_2", b, 1280, 1280, "a_l1")]
    for cout, cnt in ((320, 4), (640, 6), (1280, 7)):
        layers += [(f"time_emb_proj.{cout}", b, cout, 1280, "temb")] * cnt
    return layers


def sdxl_unet_linears(batch=4, cfg=True):
    """SDXL UNet at 1024x1024 (latent 128x128): 640 x 10 blocks @4096 tok, 1280 x 60 blocks @1024 tok."""
    return _merge([e[:4] for e in _sdxl(batch, cfg)])


def _sd35(batch):
    D, FF, mi, mt = 2432, 9728, batch * 4096, batch * 333
    layers = []
    for blk in range(38):
        last = blk == 37
        layers += [("norm1.linear", batch, 6 * D, D, "emb"), ("norm1_context.linear", batch, (2 if last else 6) * D, D, "emb")]
        layers += [(f"attn.{n}", mi, D, D, f"b{blk}.xn" if n != "to_out.0" else f"b{blk}.xo") for n in ("to_q", "to_k", "to_v", "to_out.0")]
        layers += [(f"attn.{n}", mt, D, D, f"b{blk}.cn") for n in ("add_q_proj", "add_k_proj", "add_v_proj")]
        layers += [("ff.net.0.proj", mi, FF, D, f"b{blk}.n2"), ("ff.net.2", mi, D, FF, f"b{blk}.ff")]
        if not last:
            layers += [("attn.to_add_out", mt, D, D, f"b{blk}.co"), ("ff_context.net.0.proj", mt, FF, D, f"b{blk}.n2c"),
                       ("ff_context.net.2", mt, D, FF, f"b{blk}.ffc")]
    layers += [("context_embedder", mt, D, 4096, "ctx_in"), ("time_text_embed.t1", batch, D, 256, "t_sin"),
               ("time_text_embed.t2", batch, D, D, "t_l1"), ("time_text_embed.p1", batch, D, 2048, "pooled"),
               ("time_text_embed.p2", batch, D, D, "p_l1"), ("norm_out.linear", batch, 2 * D, D, "emb"), ("proj_out", mi, 64, D, "out")]
    return layers


def sd35_mmdit_linears(batch=1):
    """SD3.5-Large MMDiT at 1024x1024: 38 joint blocks, D=2432, FF 9728, 4096 image + 333 text tokens."""
    return _merge([e[:4] for e in _sd35(batch)])


def sd15_unet_linears_fused(batch=8, cfg=True):
    """The same Linears with the same-input ones as one launch: (name, M, N_total, K, count, parts).  attn1 q/k/v per block,
    attn2 k/v of all 16 blocks (prompt embedding), the 22 time_emb_proj (silu(temb)): 184 calls -> 100 launches."""
    return _fuse(_sd15(batch, cfg))


def sdxl_unet_linears_fused(batch=4, cfg=True):
    return _fuse(_sdxl(batch, cfg))


def sd35_mmdit_linears_fused(batch=1):
    """q/k/v and add_q/k/v per block; the 76 AdaLN modulation Linears + norm_out (all read silu(emb)) as one launch."""
    return _fuse(_sd35(batch))


def total_flops(layers):
    return sum(2.0 * e[1] * e[2] * e[3] * e[4] for e in layers)


def gemm_bytes_w4a16(m, n, k, g):
    """algorithmic bytes of one W4A16 call."""
    return 2 * m * k + 0.5 * k * n + 2.5 * (k // g) * n + 2 * m * n


def gemm_bytes_w8a8(m, n, k):
    return m * k + 4 * m + k * n + 4 * n + 2 * m * n
