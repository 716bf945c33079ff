"""TEST INFRASTRUCTURE -- the CPU oracle for the hot path.  Not product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module.

A functional fp32 restatement (plain torch CPU tensor ops on ``state_dict`` entries,
no ``nn.Module``) of the reference's algorithm for

  * ``CausalConditionalDecoder.forward``        speech/cosyvoice/flow/decoder.py:405-496
  * ``ConditionalCFM.solve_euler`` + CFG         speech/cosyvoice/flow/flow_matching.py:74-126
  * ``CausalConditionalCFM.forward``             speech/cosyvoice/flow/flow_matching.py:323-348
  * ``DACVAE.decode``                            dac-vae/model.py:485-488 (+ :107-143, :237-379)

Parity status: the reference holds NO golden vectors / tests for this path
(SURVEY.md §4, §8c).  This restatement is therefore pinned against outputs of the
reference itself, imported unmodified through ``oracle/ref_import.py`` in the build
container, with weights from ``minimax-speech_b200/synth.py`` loaded into the
reference modules (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``,
``tests/test_oracle_golden.py``).  The third-party ``diffusers==0.29.0`` attention /
GELU arithmetic is restated from its published algorithm (see ref_import.py).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Estimator (CausalConditionalDecoder with channels=[C])
# --------------------------------------------------------------------------------------


def sinusoidal_pos_emb(t, dim=320, scale=1000.0):
    """matcha/models/components/decoder.py:14-29."""
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000.0) / (half - 1)))
    e = scale * t.float().unsqueeze(1) * f.unsqueeze(0)
    return torch.cat([e.sin(), e.cos()], dim=-1)


def time_mlp(sd, t, in_channels=320):
    """decoder.py:422-423 -> TimestepEmbedding (matcha decoder.py:73-117): Linear, SiLU, Linear."""
    e = sinusoidal_pos_emb(t, in_channels).to(t.dtype)
    h = F.linear(e, sd["time_mlp.linear_1.weight"], sd["time_mlp.linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, sd["time_mlp.linear_2.weight"], sd["time_mlp.linear_2.bias"])


def causal_conv(x, w, b):
    """decoder.py:59-62: left-pad k-1 zeros, no right context."""
    return F.conv1d(F.pad(x, (w.shape[-1] - 1, 0)), w, b)


def causal_block(sd, p, x, mask):
    """CausalBlock1D decoder.py:65-78: conv3 -> LayerNorm over channels -> Mish, masked."""
    h = causal_conv(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"])
    h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd[p + ".block.2.weight"], sd[p + ".block.2.bias"], 1e-5)
    h = F.mish(h.transpose(1, 2))
    return h * mask


def block1d(sd, p, x, mask):
    """Block1D matcha decoder.py:32-43 (the non-causal ConditionalDecoder's block): Conv1d(k 3, pad 1) -> GroupNorm(8) ->
    Mish, masked.  Restated with PER-UTTERANCE semantics (one reference call per utterance at its own length, batch 1 as
    the reference's solve_euler requires): the GroupNorm statistics of utterance b run over its valid frames only."""
    h = F.conv1d(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    out = torch.zeros_like(h)
    for b in range(h.shape[0]):
        n = int(mask[b, 0].sum().item())
        if n:
            out[b:b + 1, :, :n] = F.group_norm(h[b:b + 1, :, :n], 8, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], 1e-5)
    return F.mish(out) * mask


def is_causal(sd):
    """CausalConditionalDecoder keeps its LayerNorm at block.2 (after a Transpose), ConditionalDecoder its GroupNorm at block.1."""
    return "final_block.block.2.weight" in sd


def resnet_block(sd, p, x, mask, temb):
    """ResnetBlock1D.forward matcha decoder.py:56-61 with causal (flow/decoder.py:81-85) or plain blocks."""
    blk = causal_block if is_causal(sd) else block1d
    h = blk(sd, p + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"]).unsqueeze(-1)
    h = blk(sd, p + ".block2", h, mask)
    return h + F.conv1d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])


def attention_bias(mask, streaming, chunk):
    """add_optional_chunk_mask mask.py:161-236 + mask_to_bias common.py:160-168.
    mask [B,1,T] (0/1) -> additive bias [B,T,T]."""
    m = mask.bool()
    B, _, T = m.shape
    if streaming and chunk > 0:
        pos = torch.arange(T, device=mask.device)
        block_end = (torch.div(pos, chunk, rounding_mode="trunc") + 1) * chunk  # mask.py:154-157
        cm = pos.unsqueeze(0) < block_end.unsqueeze(1)
        m = m & cm.unsqueeze(0)
    else:
        m = m.repeat(1, T, 1)
    dead = m.sum(dim=-1) == 0  # mask.py:233-235
    m = m.clone()
    m[dead] = True
    return (1.0 - m.float()) * -1.0e10


def transformer_block(sd, p, u, bias, heads):
    """BasicTransformerBlock.forward transformer.py:243-316 (self-attn + FF, pre-LN)."""
    B, T, C = u.shape
    n = F.layer_norm(u, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    q = F.linear(n, sd[p + ".attn1.to_q.weight"])
    k = F.linear(n, sd[p + ".attn1.to_k.weight"])
    v = F.linear(n, sd[p + ".attn1.to_v.weight"])
    hd = q.shape[-1] // heads
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5) + bias.unsqueeze(1)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = o.transpose(1, 2).reshape(B, T, heads * hd)
    u = u + F.linear(o, sd[p + ".attn1.to_out.0.weight"], sd[p + ".attn1.to_out.0.bias"])
    n = F.layer_norm(u, (C,), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"], 1e-5)
    h = F.gelu(F.linear(n, sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"]))  # exact erf
    return u + F.linear(h, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"])


def _count(sd, fmt):
    n = 0
    while (fmt.format(n)) in sd:
        n += 1
    return n


def estimator_forward(sd, x, mask, mu, t, spks, cond, streaming=False, heads=8, chunk=50, taps=None):
    """CausalConditionalDecoder.forward decoder.py:405-496 for channels=[C].
    x,mu,cond [R,80,T]; mask [R,1,T]; t [R]; spks [R,80] -> [R,80,T].
    ``taps`` (optional dict) receives named intermediates for kernel-level tests."""
    n_blocks = _count(sd, "down_blocks.0.1.{}.norm1.weight")
    n_mid = _count(sd, "mid_blocks.{}.0.mlp.1.weight")
    temb = time_mlp(sd, t, sd["time_mlp.linear_1.weight"].shape[1])
    T = x.shape[-1]
    h = torch.cat([x, mu, spks.unsqueeze(-1).expand(-1, -1, T), cond], dim=1)  # decoder.py:427-433
    bias = attention_bias(mask, streaming, chunk)

    def group(h, prefix):
        h = resnet_block(sd, prefix + ".0", h, mask, temb)
        if taps is not None:
            taps[prefix + ".0"] = h
        u = h.transpose(1, 2)
        for j in range(n_blocks):
            u = transformer_block(sd, f"{prefix}.1.{j}", u, bias, heads)
            if taps is not None:
                taps[f"{prefix}.1.{j}"] = u
        return u.transpose(1, 2)

    causal = is_causal(sd)
    conv3 = causal_conv if causal else (lambda xx, w, b: F.conv1d(xx, w, b, padding=1))  # decoder.py:141 / :186
    h = group(h, "down_blocks.0")
    skip = h
    h = conv3(h * mask, sd["down_blocks.0.2.weight"], sd["down_blocks.0.2.bias"])
    for i in range(n_mid):
        h = group(h, f"mid_blocks.{i}")
    h = torch.cat([h, skip], dim=1)
    h = group(h, "up_blocks.0")
    h = conv3(h * mask, sd["up_blocks.0.2.weight"], sd["up_blocks.0.2.bias"])
    h = (causal_block if causal else block1d)(sd, "final_block", h, mask)
    out = F.conv1d(h * mask, sd["final_proj.weight"], sd["final_proj.bias"])
    return out * mask


# --------------------------------------------------------------------------------------
# CFM Euler solve with classifier-free guidance
# --------------------------------------------------------------------------------------


def cosine_t_span(n_timesteps, dtype=torch.float32):
    """flow_matching.py:345-347."""
    t = torch.linspace(0, 1, n_timesteps + 1, dtype=dtype)
    return 1 - torch.cos(t * 0.5 * torch.pi)


def solve_euler(sd, z, t_span, mu, mask, spks, cond, cfg_rate=0.7, streaming=False, heads=8, chunk=50):
    """ConditionalCFM.solve_euler flow_matching.py:74-126, generalised to B>=1 with the
    reference's per-utterance semantics (the reference hard-codes B=1): row b of
    the conditional half sees (mu,spks,cond)[b]; the unconditional half sees zeros."""
    x = z.clone()
    B = x.shape[0]
    t, dt = t_span[0], t_span[1] - t_span[0]
    zeros3, zeros2 = torch.zeros_like(mu), torch.zeros_like(spks)
    for step in range(1, len(t_span)):
        xin = torch.cat([x, x], 0)
        v = estimator_forward(sd, xin, torch.cat([mask, mask], 0), torch.cat([mu, zeros3], 0),
                              t.reshape(1).expand(2 * B), torch.cat([spks, zeros2], 0),
                              torch.cat([cond, zeros3], 0), streaming, heads, chunk)
        v = (1.0 + cfg_rate) * v[:B] - cfg_rate * v[B:]  # :118-119
        x = x + dt * v
        t = t + dt
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t
    return x.float()


def cfm_forward(sd, noise, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, streaming=False,
                cfg_rate=0.7, heads=8, chunk=50):
    """CausalConditionalCFM.forward flow_matching.py:323-348.  ``noise`` = rand_noise [1,80,>=T]."""
    z = noise[:, :, :mu.shape[2]].to(device=mu.device, dtype=mu.dtype).expand(mu.shape[0], -1, -1) * temperature
    t_span = cosine_t_span(n_timesteps, mu.dtype).to(mu.device)
    return solve_euler(sd, z, t_span, mu, mask, spks, cond, cfg_rate, streaming, heads, chunk)


def cfm_forward_cached(sd, z, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, prompt_len=0, cache=None,
                       cfg_rate=0.7, heads=8, chunk=50):
    """The non-causal twin ``ConditionalCFM.forward`` flow_matching.py:39-72 (batch 1): ``z`` is the injected
    ``torch.randn_like(mu)`` sample (before temperature); the first ``cache.shape[2]`` frames of z and mu are replaced by
    the cached prompt + overlap frames (:60-64), the new cache is [z | mu] over the first ``prompt_len`` and the last 34
    frames (:65-67).  Returns (latent, cache [1,80,prompt_len+34,2]); ``mu`` is NOT modified in place here."""
    z = z.to(mu.dtype) * temperature
    mu = mu.clone()
    if cache is not None and cache.shape[2] != 0:
        n = cache.shape[2]
        z = z.clone()
        z[:, :, :n] = cache[:, :, :, 0]
        mu[:, :, :n] = cache[:, :, :, 1]
    z_cache = torch.cat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2)
    mu_cache = torch.cat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)
    new_cache = torch.stack([z_cache, mu_cache], dim=-1)
    t_span = cosine_t_span(n_timesteps, mu.dtype).to(mu.device)
    return solve_euler(sd, z, t_span, mu, mask, spks, cond, cfg_rate, False, heads, chunk), new_cache


# --------------------------------------------------------------------------------------
# DAC-VAE decoder
# --------------------------------------------------------------------------------------


def wn_weight(sd, p):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but 0
    (dim 0 is Cin for ConvTranspose1d) -- layers.py:9-14."""
    v, g = sd[p + ".weight_v"], sd[p + ".weight_g"]
    return v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1))


def snake(x, alpha):
    """layers.py:18-24."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def wn_conv_lrelu(sd, p, x, dilation=1, padding=0):
    """Generator Conv1d = Sequential(weight_norm(Conv1d), LeakyReLU(0.1)) -- the shadowing
    WNConv1d at dac-vae/model.py:509-514."""
    return F.leaky_relu(F.conv1d(x, wn_weight(sd, p + ".0"), sd[p + ".0.bias"], dilation=dilation,
                                 padding=padding), 0.1)


def residual_unit(sd, p, x, dilation):
    """ResidualUnit model.py:107-143."""
    y = snake(x, sd[p + ".block.0.alpha"])
    y = wn_conv_lrelu(sd, p + ".block.1", y, dilation, 3 * dilation)
    y = snake(y, sd[p + ".block.2.alpha"])
    y = wn_conv_lrelu(sd, p + ".block.3", y)
    return x + y


def dac_decode(sd, z, taps=None):
    """DACVAE.decode model.py:485-488: de_conv_pre -> Decoder (model.py:326-379). z [B,80,L]."""
    rates = []
    while f"decoder.model.{len(rates) + 1}.block.1.weight_v" in sd:
        w = sd[f"decoder.model.{len(rates) + 1}.block.1.weight_v"]
        rates.append(w.shape[-1] // 2)
    x = wn_conv_lrelu(sd, "de_conv_pre", z)
    x = wn_conv_lrelu(sd, "decoder.model.0", x, 1, 3)
    for i, s in enumerate(rates):
        p = f"decoder.model.{i + 1}.block"
        x = snake(x, sd[p + ".0.alpha"])
        x = F.conv_transpose1d(x, wn_weight(sd, p + ".1"), sd[p + ".1.bias"], stride=s,
                               padding=math.ceil(s / 2), output_padding=s % 2)  # model.py:255-262
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{p}.{j + 2}", x, d)
        if taps is not None:
            taps[f"stage{i + 1}"] = x
    n = len(rates)
    x = snake(x, sd[f"decoder.model.{n + 1}.alpha"])
    x = wn_conv_lrelu(sd, f"decoder.model.{n + 2}", x, 1, 3)
    return torch.tanh(x)


def dac_encode(sd, audio, noise=None):
    """DACVAE.encode model.py:469-483 after Encoder (model.py:146-234): audio [B,1,S] (S a multiple of the hop) ->
    (z, m, logs), each [B,latent,S/hop].  ``noise`` ([B,latent,L], optional) replaces torch.randn_like for parity
    (z = m + noise * exp(logs)); without it z = m."""
    rates = []
    while f"encoder.block.{len(rates) + 1}.block.4.0.weight_v" in sd:
        rates.append(sd[f"encoder.block.{len(rates) + 1}.block.4.0.weight_v"].shape[-1] // 2)
    x = wn_conv_lrelu(sd, "encoder.block.0", audio, 1, 3)
    for i, s in enumerate(rates):
        p = f"encoder.block.{i + 1}.block"
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{p}.{j}", x, d)
        x = snake(x, sd[p + ".3.alpha"])
        x = F.leaky_relu(F.conv1d(x, wn_weight(sd, p + ".4.0"), sd[p + ".4.0.bias"], stride=s,
                                  padding=math.ceil(s / 2)), 0.1)  # model.py:183-189 (+ the shadow's LeakyReLU)
    n = len(rates)
    x = snake(x, sd[f"encoder.block.{n + 1}.alpha"])
    x = wn_conv_lrelu(sd, f"encoder.block.{n + 2}", x, 1, 1)
    x = F.leaky_relu(x)  # model.py:475 (slope 0.01)
    x = wn_conv_lrelu(sd, "en_conv_post", x)
    latent = x.shape[1] // 2
    m, logs = torch.split(x, latent, dim=1)
    logs = torch.clamp(logs, min=-14.0, max=14.0)
    z = m if noise is None else m + noise * torch.exp(logs)
    return z, m, logs


def dac_decode_varlen(sd, z, lengths):
    """Per-utterance semantics for a right-padded batch: utterance b is decoded alone at its own
    length (this is what a loop over the reference's ``decode`` gives); padding is zero."""
    hop = 1
    k = 1
    while f"decoder.model.{k}.block.1.weight_v" in sd:
        hop *= sd[f"decoder.model.{k}.block.1.weight_v"].shape[-1] // 2
        k += 1
    out = torch.zeros(z.shape[0], 1, z.shape[2] * hop, device=z.device)
    for b, n in enumerate(lengths):
        out[b, :, :n * hop] = dac_decode(sd, z[b:b + 1, :, :n])[0]
    return out


# --------------------------------------------------------------------------------------
# token -> mu front half: flow/flow.py:437-511, transformer/upsample_encoder.py
# --------------------------------------------------------------------------------------
def rel_pos_emb(T, d):
    """EspnetRelPositionalEncoding (transformer/embedding.py:224-300): [1, 2T-1, d], row n = relative position T-1-n."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def rel_attention(sd, p, x, pos_emb, key_mask, heads, chunk=0):
    """RelPositionMultiHeadedAttention.forward (transformer/attention.py:249-330) + forward_attention (:84-127).
    x [B,T,d]; key_mask [B,T] bool; bd[i,j] = (q_i + v) . p[T-1-i+j] (the rel_shift of :225-247 in closed form)."""
    B, T, d = x.shape
    dk = d // heads
    q = F.linear(x, sd[p + ".linear_q.weight"], sd[p + ".linear_q.bias"]).view(B, T, heads, dk)
    k = F.linear(x, sd[p + ".linear_k.weight"], sd[p + ".linear_k.bias"]).view(B, T, heads, dk).transpose(1, 2)
    v = F.linear(x, sd[p + ".linear_v.weight"], sd[p + ".linear_v.bias"]).view(B, T, heads, dk).transpose(1, 2)
    pp = F.linear(pos_emb, sd[p + ".linear_pos.weight"]).view(1, -1, heads, dk).transpose(1, 2)  # [1,h,2T-1,dk]
    qu = (q + sd[p + ".pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[p + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd_full = torch.matmul(qv, pp.transpose(-2, -1))  # [B,h,T,2T-1]
    idx = (T - 1 - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0)).to(x.device)  # [T,T]
    bd = torch.gather(bd_full, 3, idx.expand(B, heads, T, T))
    scores = (ac + bd) / math.sqrt(dk)
    dead = ~key_mask.view(B, 1, 1, T)
    if chunk > 0:  # add_optional_chunk_mask (utils/mask.py:161-236): key j visible to query i iff j < (i/chunk+1)*chunk
        pos = torch.arange(T, device=x.device)
        vis = pos.unsqueeze(0) < ((torch.div(pos, chunk, rounding_mode="trunc") + 1) * chunk).unsqueeze(1)
        dead = dead | ~vis.view(1, 1, T, T)
    attn = torch.softmax(scores.masked_fill(dead, float("-inf")), dim=-1).masked_fill(dead, 0.0)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, d)
    return F.linear(o, sd[p + ".linear_out.weight"], sd[p + ".linear_out.bias"])


def conformer_layer(sd, p, x, pos_emb, key_mask, heads, chunk=0):
    """ConformerEncoderLayer.forward (transformer/encoder_layer.py:109-) for normalize_before, no macaron, no conv
    module (config.yaml:73-88): x += attn(LN(x)); x += FF(LN(x)), FF = w_2(swish(w_1(.)))."""
    n = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm_mha.weight"], sd[p + ".norm_mha.bias"], 1e-5)
    x = x + rel_attention(sd, p + ".self_attn", n, pos_emb, key_mask, heads, chunk)
    n = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm_ff.weight"], sd[p + ".norm_ff.bias"], 1e-5)
    h = F.silu(F.linear(n, sd[p + ".feed_forward.w_1.weight"], sd[p + ".feed_forward.w_1.bias"]))
    return x + F.linear(h, sd[p + ".feed_forward.w_2.weight"], sd[p + ".feed_forward.w_2.bias"])


def upsample_conformer_encode(sd, xs, lens, heads=8, prefix="encoder.", context=None, streaming=False, chunk=25):
    """UpsampleConformerEncoder.forward (upsample_encoder.py:266-318).  xs [B,T,512] (embedded tokens), lens [B] ->
    ([B,2T,512], 2*lens).  ``context`` [B,3,512]: the look-ahead tokens of a non-final chunk (flow.py:482-489), else the
    input is right-padded with zeros; ``streaming``: block-causal attention, chunk 25 tokens then 50 frames."""
    P = prefix
    B, T, d = xs.shape
    key_mask = torch.arange(T, device=xs.device).unsqueeze(0) < lens.unsqueeze(1)

    def embed(pp, x):  # LinearNoSubsampling (subsampling.py:69-113) + rel-pos encoding scale (embedding.py:268)
        x = F.linear(x, sd[pp + ".out.0.weight"], sd[pp + ".out.0.bias"])
        x = F.layer_norm(x, (d,), sd[pp + ".out.1.weight"], sd[pp + ".out.1.bias"], 1e-5)
        return x * math.sqrt(d), rel_pos_emb(x.shape[1], d).to(x.device)

    x, pos = embed(P + "embed", xs)
    # PreLookaheadLayer (upsample_encoder.py:66-107) without context: right-pad 3, conv k=4, leaky_relu, causal conv k=3
    if context is not None:  # context goes through the same embed (offset only moves the unused pos_emb)
        ctx, _ = embed(P + "embed", context)
        y = torch.cat([x, ctx], dim=1).transpose(1, 2)
    else:
        y = F.pad(x.transpose(1, 2), (0, 3))
    y = F.leaky_relu(F.conv1d(y, sd[P + "pre_lookahead_layer.conv1.weight"], sd[P + "pre_lookahead_layer.conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd[P + "pre_lookahead_layer.conv2.weight"], sd[P + "pre_lookahead_layer.conv2.bias"])
    x = y.transpose(1, 2) + x
    i = 0
    while f"{P}encoders.{i}.norm_mha.weight" in sd:
        x = conformer_layer(sd, f"{P}encoders.{i}", x, pos, key_mask, heads, chunk if streaming else 0)
        i += 1
    # Upsample1D (upsample_encoder.py:37-63): nearest x2, left-pad 4, conv k=5
    y = F.interpolate(x.transpose(1, 2), scale_factor=2.0, mode="nearest")
    y = F.conv1d(F.pad(y, (4, 0)), sd[P + "up_layer.conv.weight"], sd[P + "up_layer.conv.bias"])
    lens2 = lens * 2
    key_mask = torch.arange(2 * T, device=xs.device).unsqueeze(0) < lens2.unsqueeze(1)
    x, pos = embed(P + "up_embed", y.transpose(1, 2))
    i = 0
    while f"{P}up_encoders.{i}.norm_mha.weight" in sd:
        x = conformer_layer(sd, f"{P}up_encoders.{i}", x, pos, key_mask, heads, 2 * chunk if streaming else 0)
        i += 1
    x = F.layer_norm(x, (d,), sd[P + "after_norm.weight"], sd[P + "after_norm.bias"], 1e-5)
    return x, lens2


def tokens_to_mu(sd, token, embedding, finalize=True, streaming=False, pre_lookahead_len=3):
    """CausalMaskedDiffWithXvec.inference front half (flow.py:461-489) for one utterance, no prompt:
    (mu [1,80,2T], spks [1,80]); finalize=False: the last 3 tokens are look-ahead context only (T shrinks by 3)."""
    spks = F.linear(F.normalize(embedding, dim=1), sd["spk_embed_affine_layer.weight"], sd["spk_embed_affine_layer.bias"])
    x = F.embedding(torch.clamp(token, min=0), sd["input_embedding.weight"])
    lens = torch.tensor([token.shape[1]])
    if finalize:
        h, _ = upsample_conformer_encode(sd, x, lens, streaming=streaming)
    else:
        h, _ = upsample_conformer_encode(sd, x[:, :-pre_lookahead_len], lens, context=x[:, -pre_lookahead_len:],
                                         streaming=streaming)
    mu = F.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"])
    return mu.transpose(1, 2).contiguous(), spks


# --------------------------------------------------------------------------------------
# speaker encoder: llm/llm.py:34-96, transformer/arch_util.py:80-123
# --------------------------------------------------------------------------------------
def speaker_encode(sd, mel, heads=8):
    """LearnableSpeakerEncoder.forward: mel [B,80,T] -> L2-normalised embedding [B,192] (first-frame pooling)."""
    h = F.conv1d(mel, sd["init.weight"], sd["init.bias"])
    i = 0
    while f"attn.{i}.norm.weight" in sd:
        p = f"attn.{i}"
        B, C, T = h.shape
        n = F.group_norm(h, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)  # normalization(): 32 groups
        qkv = F.conv1d(n, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
        ch = C // heads
        q, k, v = qkv.reshape(B * heads, 3 * ch, T).split(ch, dim=1)  # QKVAttentionLegacy: head-major [q|k|v] blocks
        w = torch.softmax(torch.einsum("bct,bcs->bts", q, k) / math.sqrt(ch), dim=-1)
        a = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T)
        h = h + F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
        i += 1
    out = F.linear(h[:, :, 0], sd["output_proj.weight"], sd["output_proj.bias"])
    return F.normalize(out, p=2, dim=1)


def flow_inference(sd_flow, sd_est, noise, token, prompt_token, prompt_feat, embedding=None, reference_mels=None, sd_spk=None,
                   streaming=False, finalize=False, n_timesteps=10, pre_lookahead_len=3):
    """CausalMaskedDiffWithXvec.inference (flow/flow.py:437-511), batch 1: speaker embedding (given, or from reference
    mels [1,80,T] / [1,N,80,T] through the speaker encoder, flow.py:336-366), [prompt_token | token] -> mu, cond = prompt_feat
    on the first frames, CFM solve (10 steps), the prompt frames cut off.  -> latents [1,80,2*T_token(-6)]."""
    if reference_mels is not None:
        if reference_mels.dim() == 4:
            embs = [speaker_encode(sd_spk, reference_mels[:, i]) for i in range(reference_mels.shape[1])]
            embedding = torch.stack(embs, dim=1).mean(dim=1)
        else:
            embedding = speaker_encode(sd_spk, reference_mels)
    elif embedding is None:
        embedding = torch.zeros(1, sd_flow["spk_embed_affine_layer.weight"].shape[1])
    tok = torch.cat([prompt_token, token], dim=1)
    mu, spks = tokens_to_mu(sd_flow, tok, embedding, finalize=finalize, streaming=streaming, pre_lookahead_len=pre_lookahead_len)
    mel_len1 = prompt_feat.shape[1]
    cond = torch.zeros_like(mu)
    cond[:, :, :mel_len1] = prompt_feat.transpose(1, 2)
    mask = torch.ones(1, 1, mu.shape[2])
    feat = cfm_forward(sd_est, noise, mu, mask, n_timesteps, 1.0, spks, cond, streaming=streaming)
    return feat[:, :, mel_len1:]


# --------------------------------------------------------------------------------------
# Parity metrics
# --------------------------------------------------------------------------------------


def fsq_encode(weight, bias, x):
    """FSQCodebook.encode tools/S3Tokenizer/s3tokenizer/model_v2.py:99-112: project_down -> tanh -> * 0.999 -> round, + 1,
    base-3 digits.  x [B, T, D] -> int32 [B, T]."""
    h = torch.tanh(F.linear(x.reshape(-1, x.shape[-1]), weight, bias).float()) * 0.9990000128746033
    h = h.round() + 1
    powers = torch.pow(3, torch.arange(8, dtype=h.dtype))
    return torch.sum(h * powers.unsqueeze(0), dim=-1).reshape(x.shape[0], x.shape[1]).int()



def s3_encode(sd, mel, mel_len):
    """AudioEncoderV2.forward (tools/S3Tokenizer/s3tokenizer/model_v2.py:320-351) with FSMNMultiHeadAttention
    (:152-249), ResidualAttentionBlock (:252-287), the rotary embedding (:37-70) and the masks of utils.py
    (make_non_pad_mask, mask_to_bias).  mel [B, n_mels, T] fp32, mel_len [B] -> (hidden [B, T', n_state], code_len [B]);
    T' = ((T - 1) // 2 + 1 - 1) // 2 + 1.  Frames past code_len carry whatever the reference computes there."""
    def non_pad(lengths, T):
        return (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1))

    n_state = sd["encoder.conv1.weight"].shape[0]
    H = n_state // 64
    mel_len = mel_len.to(torch.int64)
    x = mel.float() * non_pad(mel_len, mel.shape[2]).unsqueeze(1)
    x = F.gelu(F.conv1d(x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], stride=2, padding=1))
    l1 = (mel_len + 2 - 2 - 1) // 2 + 1
    x = x * non_pad(l1, x.shape[2]).unsqueeze(1)
    x = F.gelu(F.conv1d(x, sd["encoder.conv2.weight"], sd["encoder.conv2.bias"], stride=2, padding=1))
    l2 = (l1 + 2 - 2 - 1) // 2 + 1
    x = x.permute(0, 2, 1)  # [B, T', n_state]
    B, T, _ = x.shape
    mask_pad = non_pad(l2, T).unsqueeze(2).float()          # [B, T', 1]
    bias = (1.0 - non_pad(l2, T).float()) * -1.0e10          # [B, T'] additive, per key
    # rotary table: angle(t, d) = t * 10000^(-2 (d mod 32) / 64)  (precompute_freqs_cis(64, .) concatenated with itself)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    ang = torch.outer(torch.arange(T).float(), inv)
    cos = torch.cat((ang.cos(), ang.cos()), -1)[None, :, None, :]
    sin = torch.cat((ang.sin(), ang.sin()), -1)[None, :, None, :]

    def rot(a):  # [B, T, H, 64]
        return a * cos + torch.cat((-a[..., 32:], a[..., :32]), -1) * sin

    i = 0
    while f"encoder.blocks.{i}.attn.query.weight" in sd:
        p = f"encoder.blocks.{i}"
        a = F.layer_norm(x, (n_state,), sd[p + ".attn_ln.weight"], sd[p + ".attn_ln.bias"], 1e-6)
        q = F.linear(a, sd[p + ".attn.query.weight"], sd[p + ".attn.query.bias"]).view(B, T, H, 64)
        k = F.linear(a, sd[p + ".attn.key.weight"]).view(B, T, H, 64)
        v = F.linear(a, sd[p + ".attn.value.weight"], sd[p + ".attn.value.bias"])
        q, k = rot(q), rot(k)
        vm = v * mask_pad  # forward_fsmn: depthwise conv over time (kernel 31, 15 + 15 zero padding) + its input, masked
        w = sd[p + ".attn.fsmn_block.weight"]
        ks = w.shape[-1]
        mem = F.conv1d(F.pad(vm.transpose(1, 2), ((ks - 1) // 2, ks - 1 - (ks - 1) // 2)), w, groups=n_state).transpose(1, 2)
        mem = (mem + vm) * mask_pad
        scale = 64 ** -0.25
        qk = torch.einsum("bthd,bshd->bhts", q * scale, k * scale) + bias[:, None, None, :]
        o = torch.einsum("bhts,bshd->bthd", torch.softmax(qk.float(), -1), v.view(B, T, H, 64)).reshape(B, T, n_state)
        x = x + F.linear(o, sd[p + ".attn.out.weight"], sd[p + ".attn.out.bias"]) + mem
        m = F.layer_norm(x, (n_state,), sd[p + ".mlp_ln.weight"], sd[p + ".mlp_ln.bias"], 1e-5)
        x = x + F.linear(F.gelu(F.linear(m, sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"])), sd[p + ".mlp.2.weight"],
                         sd[p + ".mlp.2.bias"])
        i += 1
    return x, l2.to(torch.int32)


def s3_quantize(sd, mel, mel_len):
    """S3TokenizerV2.quantize (model_v2.py:386-415): encoder trunk, then the FSQ head; batches holding a clip longer than
    30 s (3000 mel frames) go through _quantize_mixed_batch (:417-588) with merge_tokenized_segments (utils.py:367-390)."""
    w, b = sd["quantizer._codebook.project_down.weight"], sd["quantizer._codebook.project_down.bias"]
    lens = [int(v) for v in mel_len]
    if max(lens) <= 3000:
        hidden, code_len = s3_encode(sd, mel, mel_len)
        return fsq_encode(w, b, hidden), code_len
    window, overlap_s = 3000, 4
    stride = window - overlap_s * 100
    pieces, piece_len, of = [], [], []
    for i, n in enumerate(lens):
        start = 0
        while True:  # one pass for a short clip; windows every `stride` frames for a long one
            seg = mel[i, :, start:min(start + window, n)]
            piece_len.append(seg.shape[1])
            pieces.append(F.pad(seg, (0, window - seg.shape[1])))
            of.append(i)
            start += stride
            if n <= 3000 or start >= n:
                break
    hidden, code_len = s3_encode(sd, torch.stack(pieces), torch.tensor(piece_len))
    codes = fsq_encode(w, b, hidden)
    drop = (overlap_s // 2) * 25
    out = []
    for i, n in enumerate(lens):
        idx = [j for j, o in enumerate(of) if o == i]
        toks = []
        for pos, j in enumerate(idx):
            t = codes[j, :int(code_len[j])].tolist()
            if n > 3000:
                t = t[(0 if pos == 0 else drop):(len(t) if pos == len(idx) - 1 else len(t) - drop)]
            toks += t
        out.append(toks)
    res = torch.zeros(len(lens), max(len(t) for t in out), dtype=torch.long)
    for i, t in enumerate(out):
        res[i, :len(t)] = torch.tensor(t, dtype=torch.long)
    return res, torch.tensor([len(t) for t in out], dtype=torch.long)

def rel_l2(y, ref):
    y, ref = y.double(), ref.double()
    return float((y - ref).norm() / ref.norm().clamp_min(1e-30))


def snr_db(y, ref):
    y, ref = y.double(), ref.double()
    return float(10.0 * torch.log10(ref.pow(2).sum() / (ref - y).pow(2).sum().clamp_min(1e-30)))
