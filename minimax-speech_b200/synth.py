"""Deterministic synthetic weights and inputs for the hot path.

The reference ships no checkpoints and its constructors are not
available on the GPU box, so weights are generated here with the *same key schema
and shapes* as the reference ``state_dict``s from a seed,
using numpy's PCG64 stream (stable across machines for one numpy version).

``init="reference"`` follows the reference initialisers' distributions
(speech/cosyvoice/flow/decoder.py:196-208: kaiming-normal conv/linear weights, zero
biases, unit LayerNorm; dac-vae/model.py:17-104 + torch defaults: weight-norm'd
convs keep torch's default kaiming-uniform ``weight_v`` with ``weight_g=|v|``, zero
bias, Snake alpha ~ xavier-normal).  ``init="test"`` additionally randomises
biases, LayerNorm affine parameters and ``weight_g`` so that parity tests exercise
every term of the arithmetic.
"""
import math
import zlib

import numpy as np
import torch

ESTIMATOR_CFG = dict(in_channels=320, out_channels=80, channels=256, n_blocks=4, num_mid_blocks=12,
                     num_heads=8, head_dim=64, static_chunk_size=50)  # speech/config.yaml:105-116
DAC_CFG = dict(latent_dim=80, decoder_dim=1536, decoder_rates=(5, 4, 4, 3, 2), d_out=1,
               sample_rate=24000)  # dac-vae/configs/configx2.yml
TRAINED_GAIN = 1.5  # weight_g multiplier of init="trained" (DAC): activations of O(10) at the Snake inputs
CFG_RATE = 0.7  # speech/config.yaml:99
NOISE_FRAMES = 50 * 300  # flow_matching.py:321


def _rng(seed, name):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def _normal(seed, name, shape, std, mean=0.0):
    a = _rng(seed, name).standard_normal(int(np.prod(shape)), dtype=np.float32).reshape(shape)
    return torch.from_numpy(a * np.float32(std) + np.float32(mean))


def _uniform(seed, name, shape, bound):
    a = _rng(seed, name).random(int(np.prod(shape)), dtype=np.float32).reshape(shape)
    return torch.from_numpy((a * 2 - 1) * np.float32(bound))


def estimator_state_dict(seed=1986, init="reference", in_channels=320, out_channels=80, channels=256,
                         n_blocks=4, num_mid_blocks=12, num_heads=8, head_dim=64, causal=True, **_):
    """Keys/shapes of ``CausalConditionalDecoder.state_dict()`` for ``channels=[C]``; ``causal=False``: the non-causal
    ``ConditionalDecoder`` (speech/cosyvoice/flow/decoder.py:88-291), whose blocks are Conv1d(pad 1) -> GroupNorm(8) ->
    Mish (matcha Block1D): the norm parameters sit at ``block.1`` instead of ``block.2``."""
    nk = "2" if causal else "1"
    C, inner, temb = channels, num_heads * head_dim, channels * 4
    test = init == "test"
    sd = {}

    def lin(name, out_f, in_f, bias=True, k=None):
        shape = (out_f, in_f) if k is None else (out_f, in_f, k)
        fan_in = in_f * (k or 1)
        sd[name + ".weight"] = _normal(seed, name + ".weight", shape, math.sqrt(2.0 / fan_in))
        if bias:
            sd[name + ".bias"] = (_normal(seed, name + ".bias", (out_f,), 0.1) if test
                                  else torch.zeros(out_f))

    def ln(name, n):
        sd[name + ".weight"] = (_normal(seed, name + ".weight", (n,), 0.1, 1.0) if test else torch.ones(n))
        sd[name + ".bias"] = (_normal(seed, name + ".bias", (n,), 0.1) if test else torch.zeros(n))

    def resnet(p, cin):
        lin(p + ".mlp.1", C, temb)
        lin(p + ".block1.block.0", C, cin, k=3)
        ln(p + ".block1.block." + nk, C)
        lin(p + ".block2.block.0", C, C, k=3)
        ln(p + ".block2.block." + nk, C)
        lin(p + ".res_conv", C, cin, k=1)

    def tblock(p):
        ln(p + ".norm1", C)
        for n in ("to_q", "to_k", "to_v"):
            lin(p + ".attn1." + n, inner, C, bias=False)
        lin(p + ".attn1.to_out.0", C, inner)
        ln(p + ".norm3", C)
        lin(p + ".ff.net.0.proj", 4 * C, C)
        lin(p + ".ff.net.2", C, 4 * C)

    lin("time_mlp.linear_1", temb, in_channels)
    lin("time_mlp.linear_2", temb, temb)
    resnet("down_blocks.0.0", in_channels)
    for j in range(n_blocks):
        tblock(f"down_blocks.0.1.{j}")
    lin("down_blocks.0.2", C, C, k=3)
    for i in range(num_mid_blocks):
        resnet(f"mid_blocks.{i}.0", C)
        for j in range(n_blocks):
            tblock(f"mid_blocks.{i}.1.{j}")
    resnet("up_blocks.0.0", 2 * C)
    for j in range(n_blocks):
        tblock(f"up_blocks.0.1.{j}")
    lin("up_blocks.0.2", C, C, k=3)
    lin("final_block.block.0", C, C, k=3)
    ln("final_block.block." + nk, C)
    lin("final_proj", out_channels, C, k=1)
    return sd


def dac_decoder_state_dict(seed=0, init="reference", latent_dim=80, decoder_dim=1536,
                           decoder_rates=(5, 4, 4, 3, 2), d_out=1, **_):
    """Keys/shapes of the ``decoder.*`` + ``de_conv_pre.*`` part of ``DACVAE.state_dict()``."""
    trained = init == "trained"  # "test" + the regime of a trained checkpoint: Snake alpha = O(1), activations = O(10)
    test = init == "test" or trained
    sd = {}

    def wn(name, w_shape, fan_in, n_bias, transpose=False):
        v = _uniform(seed, name + ".weight_v", w_shape, 1.0 / math.sqrt(fan_in))
        g = v.reshape(w_shape[0], -1).norm(dim=1).reshape(w_shape[0], 1, 1)
        if test:
            g = g * _uniform(seed, name + ".weight_g", (w_shape[0], 1, 1), 0.3).add(1.0)
        if trained:
            g = g * TRAINED_GAIN
        sd[name + ".bias"] = (_uniform(seed, name + ".bias", (n_bias,), 1.0 / math.sqrt(fan_in)) if test
                              else torch.zeros(n_bias))
        sd[name + ".weight_g"] = g
        sd[name + ".weight_v"] = v

    def snake(name, c):
        if trained:  # alpha in [0.5, 2]: |alpha * x| reaches tens of radians (the range-reduction regime of sin)
            sd[name + ".alpha"] = _uniform(seed, name + ".alpha", (1, c, 1), 0.75).add(1.25)
        else:
            sd[name + ".alpha"] = _normal(seed, name + ".alpha", (1, c, 1), math.sqrt(2.0 / (c + 1)))

    def conv(name, cout, cin, k):  # WNConv1d shadow adds the trailing ".0" (dac-vae/model.py:509-514)
        wn(name + ".0", (cout, cin, k), cin * k, cout)

    conv("decoder.model.0", decoder_dim, latent_dim, 7)
    c = decoder_dim
    for i, s in enumerate(decoder_rates):
        p = f"decoder.model.{i + 1}.block"
        snake(p + ".0", c)
        # ConvTranspose1d weight is [Cin, Cout, k]; torch's fan_in for it is Cout*k
        wn(p + ".1", (c, c // 2, 2 * s), (c // 2) * 2 * s, c // 2)
        c //= 2
        for j in range(3):
            q = f"{p}.{j + 2}.block"
            snake(q + ".0", c)
            conv(q + ".1", c, c, 7)
            snake(q + ".2", c)
            conv(q + ".3", c, c, 1)
    n = len(decoder_rates)
    snake(f"decoder.model.{n + 1}", c)
    conv(f"decoder.model.{n + 2}", d_out, c, 7)
    conv("de_conv_pre", latent_dim, latent_dim, 1)
    return sd


def dac_encoder_state_dict(seed=0, init="reference", encoder_dim=64, encoder_rates=(2, 3, 4, 4, 5), latent_dim=80,
                           d_in=1, **_):
    """Keys/shapes of the ``encoder.*`` + ``en_conv_post.*`` part of ``DACVAE.state_dict()`` (dac-vae/model.py:146-234,
    440-442; every WNConv1d is the shadowing Sequential(conv, LeakyReLU), hence the ".0")."""
    test = init == "test"
    sd = {}

    def conv(name, cout, cin, k):
        w_shape, fan_in = (cout, cin, k), cin * k
        v = _uniform(seed, name + ".0.weight_v", w_shape, 1.0 / math.sqrt(fan_in))
        g = v.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1)
        if test:
            g = g * _uniform(seed, name + ".0.weight_g", (cout, 1, 1), 0.3).add(1.0)
        sd[name + ".0.bias"] = (_uniform(seed, name + ".0.bias", (cout,), 1.0 / math.sqrt(fan_in)) if test
                                else torch.zeros(cout))
        sd[name + ".0.weight_g"] = g
        sd[name + ".0.weight_v"] = v

    def snake(name, c):
        sd[name + ".alpha"] = _normal(seed, name + ".alpha", (1, c, 1), math.sqrt(2.0 / (c + 1)))

    conv("encoder.block.0", encoder_dim, d_in, 7)
    c = encoder_dim
    for i, st in enumerate(encoder_rates):
        p = f"encoder.block.{i + 1}.block"
        for j in range(3):
            q = f"{p}.{j}.block"
            snake(q + ".0", c)
            conv(q + ".1", c, c, 7)
            snake(q + ".2", c)
            conv(q + ".3", c, c, 1)
        snake(p + ".3", c)
        conv(p + ".4", 2 * c, c, 2 * st)
        c *= 2
    n = len(encoder_rates)
    snake(f"encoder.block.{n + 1}", c)
    conv(f"encoder.block.{n + 2}", latent_dim, c, 3)
    conv("en_conv_post", 2 * latent_dim, latent_dim, 1)
    return sd


def conformer_encoder_state_dict(seed=7, d=512, heads=8, ff=2048, num_blocks=6, num_up_blocks=4, out_dim=80, vocab=6561,
                                 spk_dim=192, **_):
    """Keys/shapes of the token -> mu front half of CausalMaskedDiffWithXvec (flow/flow.py:437-511): input_embedding,
    encoder.* (UpsampleConformerEncoder, transformer/upsample_encoder.py), encoder_proj, spk_embed_affine_layer.
    Test initialisation only (every tensor non-trivial)."""
    sd = {}

    def lin(name, n, k, bias=True):
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (n, k), 1.0 / math.sqrt(k))
        if bias:
            sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), 1.0 / math.sqrt(k))

    def ln(name, n):
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (n,), 0.2).add(1.0)
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), 0.1)

    def conv(name, n, c, k):
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (n, c, k), 1.0 / math.sqrt(c * k))
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), 1.0 / math.sqrt(c * k))

    def layer(p):
        sd[p + ".self_attn.pos_bias_u"] = _uniform(seed, p + ".u", (heads, d // heads), 0.1)
        sd[p + ".self_attn.pos_bias_v"] = _uniform(seed, p + ".v", (heads, d // heads), 0.1)
        for nm in ("q", "k", "v", "out"):
            lin(f"{p}.self_attn.linear_{nm}", d, d)
        lin(p + ".self_attn.linear_pos", d, d, bias=False)
        lin(p + ".feed_forward.w_1", ff, d)
        lin(p + ".feed_forward.w_2", d, ff)
        ln(p + ".norm_ff", d)
        ln(p + ".norm_mha", d)

    sd["input_embedding.weight"] = _normal(seed, "input_embedding.weight", (vocab, d), 1.0)
    lin("encoder.embed.out.0", d, d)
    ln("encoder.embed.out.1", d)
    conv("encoder.pre_lookahead_layer.conv1", d, d, 4)
    conv("encoder.pre_lookahead_layer.conv2", d, d, 3)
    for i in range(num_blocks):
        layer(f"encoder.encoders.{i}")
    conv("encoder.up_layer.conv", d, d, 5)
    lin("encoder.up_embed.out.0", d, d)
    ln("encoder.up_embed.out.1", d)
    for i in range(num_up_blocks):
        layer(f"encoder.up_encoders.{i}")
    ln("encoder.after_norm", d)
    lin("encoder_proj", out_dim, d)
    lin("spk_embed_affine_layer", out_dim, spk_dim)
    return sd


def speaker_encoder_state_dict(seed=13, mel_dim=80, model_dim=512, output_dim=192, num_blocks=6, **_):
    """Keys/shapes of LearnableSpeakerEncoder (llm/llm.py:34-96; AttentionBlock: transformer/arch_util.py:80-123).
    Test initialisation (the reference zero-initialises proj_out; here every tensor is non-trivial)."""
    sd = {}

    def conv(name, n, c):
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (n, c, 1), 1.0 / math.sqrt(c))
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), 1.0 / math.sqrt(c))

    conv("init", model_dim, mel_dim)
    for i in range(num_blocks):
        sd[f"attn.{i}.norm.weight"] = _uniform(seed, f"attn.{i}.norm.weight", (model_dim,), 0.2).add(1.0)
        sd[f"attn.{i}.norm.bias"] = _uniform(seed, f"attn.{i}.norm.bias", (model_dim,), 0.1)
        conv(f"attn.{i}.qkv", 3 * model_dim, model_dim)
        conv(f"attn.{i}.proj_out", model_dim, model_dim)
    sd["output_proj.weight"] = _uniform(seed, "output_proj.weight", (output_dim, model_dim), 1.0 / math.sqrt(model_dim))
    sd["output_proj.bias"] = _uniform(seed, "output_proj.bias", (output_dim,), 1.0 / math.sqrt(model_dim))
    return sd



def s3_tokenizer_state_dict(seed=21, n_mels=128, n_state=1280, n_head=20, n_layer=6, kernel_size=31, **_):
    """Keys/shapes of S3TokenizerV2 (tools/S3Tokenizer/s3tokenizer/model_v2.py:290-379: AudioEncoderV2 + the FSQ head).
    Test initialisation: nn.Linear / nn.Conv1d style bounds, LayerNorm affine near (1, 0), every tensor non-trivial."""
    assert n_state == 64 * n_head, "the reference precomputes its rotary table for 64-wide heads (model_v2.py:307)"
    sd = {}

    def lin(name, n, k, bias=True, kk=1):
        bound = 1.0 / math.sqrt(k * kk)
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (n, k, kk) if kk > 1 else (n, k), bound)
        if bias:
            sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), bound)

    lin("encoder.conv1", n_state, n_mels, kk=3)
    lin("encoder.conv2", n_state, n_state, kk=3)
    for i in range(n_layer):
        b = f"encoder.blocks.{i}"
        lin(b + ".attn.query", n_state, n_state)
        lin(b + ".attn.key", n_state, n_state, bias=False)
        lin(b + ".attn.value", n_state, n_state)
        lin(b + ".attn.out", n_state, n_state)
        sd[b + ".attn.fsmn_block.weight"] = _uniform(seed, b + ".fsmn", (n_state, 1, kernel_size), 1.0 / math.sqrt(kernel_size))
        for ln in ("attn_ln", "mlp_ln"):
            sd[f"{b}.{ln}.weight"] = _uniform(seed, f"{b}.{ln}.weight", (n_state,), 0.2).add(1.0)
            sd[f"{b}.{ln}.bias"] = _uniform(seed, f"{b}.{ln}.bias", (n_state,), 0.1)
        lin(b + ".mlp.0", 4 * n_state, n_state)
        lin(b + ".mlp.2", n_state, 4 * n_state)
    lin("quantizer._codebook.project_down", 8, n_state)
    return sd


def s3_mel(index, frames, n_mels=128):
    """Synthetic log-mel input of the tokenizer (100 frames per second; whisper-style values in about [-1, 1.5])."""
    return _normal(9000 + index, "s3mel", (1, n_mels, frames), 0.6, 0.2)

def reference_mel(index, frames, mel_dim=80):
    return _normal(8000 + index, "mel", (1, mel_dim, frames), 1.0)


def token_inputs(index, n_tokens, vocab=6561, spk_dim=192):
    """Synthetic FSQ tokens (25 Hz) and a speaker embedding."""
    r = _rng(7000 + index, "tokens")
    tok = torch.from_numpy(r.integers(0, vocab, size=(1, n_tokens)).astype(np.int64))
    emb = _normal(7000 + index, "xvec", (1, spk_dim), 1.0)
    return tok, emb


PIPE_EST = dict(n_blocks=1, num_mid_blocks=1)


def pipeline_inputs(case):
    """Inputs of the two whole-pipeline golden cases (oracle/gen_golden.py pipeline_golden and the tests)."""
    if case == "a":  # given x-vector, 8 prompt tokens / 16 prompt frames, final chunk
        tok, emb = token_inputs(30, 30)
        ptok, _ = token_inputs(31, 8)
        pfeat = dac_latents(32, 16).transpose(1, 2).contiguous()
        return dict(token=tok, prompt_token=ptok, prompt_feat=pfeat, embedding=emb, reference_mels=None, streaming=False,
                    finalize=True)
    tok, _ = token_inputs(40, 53)  # two reference clips -> speaker encoder; non-final streaming chunk
    ptok, _ = token_inputs(41, 10)
    pfeat = dac_latents(42, 20).transpose(1, 2).contiguous()
    mels = torch.stack([reference_mel(50 + i, 40) for i in range(2)], dim=1)  # [1,2,80,40]
    return dict(token=tok, prompt_token=ptok, prompt_feat=pfeat, embedding=None, reference_mels=mels, streaming=True,
                finalize=False)


def pipeline_state_dicts():
    return (conformer_encoder_state_dict(7), estimator_state_dict(7, init="test", **PIPE_EST),
            speaker_encoder_state_dict(13))


def audio_clip(index, samples):
    """Synthetic mono audio in (-1, 1): a few sinusoids plus noise, deterministic per index."""
    t = torch.arange(samples, dtype=torch.float32) / 24000.0
    f = 380.0 + 300.0 * _uniform(9000 + index, "audio.f", (4,), 1.0)
    a = 0.2 + 0.1 * _uniform(9000 + index, "audio.a", (4,), 1.0)
    x = sum(a[i] * torch.sin(2 * math.pi * f[i] * t + i) for i in range(4))
    x = x + 0.05 * _normal(9000 + index, "audio.n", (samples,), 1.0)
    return x.clamp(-0.99, 0.99).reshape(1, 1, samples)


def fixed_noise(frames=NOISE_FRAMES, channels=80):
    """``CausalConditionalCFM.rand_noise`` (flow_matching.py:320-321): seed-0 torch randn.
    Does not disturb the caller's global RNG state (the reference does)."""
    g = torch.Generator().manual_seed(0)
    return torch.randn([1, channels, NOISE_FRAMES], generator=g)[:, :, :frames]


def utterance_inputs(index, frames, channels=80):
    """Synthetic (mu, spks, cond) for utterance ``index``."""
    mu = _normal(1234 + index, "mu", (1, channels, frames), 1.0)
    spks = _normal(1234 + index, "spks", (1, channels), 1.0)
    cond = torch.zeros(1, channels, frames)
    p = min(150, frames // 5)
    cond[:, :, :p] = _normal(1234 + index, "cond", (1, channels, p), 1.0)
    return mu, spks, cond


def batch_inputs(lengths, channels=80, first_index=0):
    """Right-padded batch: mu[B,80,T], mask[B,1,T], spks[B,80], cond[B,80,T]."""
    B, T = len(lengths), max(lengths)
    mu = torch.zeros(B, channels, T)
    cond = torch.zeros(B, channels, T)
    spks = torch.zeros(B, channels)
    mask = torch.zeros(B, 1, T)
    for i, n in enumerate(lengths):
        m, s, c = utterance_inputs(first_index + i, n, channels)
        mu[i, :, :n], spks[i], cond[i, :, :n], mask[i, :, :n] = m[0], s[0], c[0], 1.0
    return mu, mask, spks, cond


def dac_latents(index, frames, channels=80):
    return _normal(4321 + index, "z", (1, channels, frames), 1.0)


def mixed_lengths(n=256, lo_s=2, hi_s=30, frame_rate=50, seed=0):
    """BASELINE config 5: ``n`` utterance lengths, uniform in whole seconds."""
    import random
    r = random.Random(seed)
    return [frame_rate * r.randint(lo_s, hi_s) for _ in range(n)]


def checksum(sd):
    """Order-independent fingerprint of a state dict (guards golden fixtures)."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float(v.abs().sum())
    return tot
