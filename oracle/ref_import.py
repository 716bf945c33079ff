"""TEST INFRASTRUCTURE -- not product code.

Imports the *unmodified* reference hot-path modules from /root/reference by
pre-populating ``sys.modules`` with stubs for the third-party packages that are
absent from this image.  Only usable inside the build
container (``/root/reference`` does not exist on the GPU box); it is used by
``oracle/gen_golden.py`` to produce the fixtures under ``tests/golden/`` that pin
``oracle/restatement.py``.

Third-party arithmetic restated here (not present under /root/reference):
``diffusers==0.29.0`` (speech/../requirements.txt): ``Attention`` with
``AttnProcessor2_0`` (q/k/v Linear(bias=False), ``to_out=[Linear(bias=True),
Dropout]``, mask ``[B,L,L] -> repeat_interleave(heads) -> [B,H,L,L]``,
``F.scaled_dot_product_attention``), ``GELU`` (= ``F.gelu(Linear(x),
approximate='none')``), ``LoRACompatibleLinear`` (= ``nn.Linear``) and
``get_activation('silu')``.  Call sites: matcha transformer.py:110,126,196-204,
266-271 and matcha decoder.py:92.
"""
import logging
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# /root/reference in the build container; on the GPU box the unmodified hot-path modules staged by oracle/stage_ref.py
REF_ROOT = os.environ.get("LS_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/speech/cosyvoice") else _STAGED)


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "speech", "cosyvoice"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Attention(nn.Module):
    """diffusers 0.29.0 ``Attention`` restricted to what BasicTransformerBlock uses
    (self-attention, no norm, no added kv), processor = AttnProcessor2_0."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0,
                 bias=False, upcast_attention=False, **kw):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=True), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        b, l, _ = hidden_states.shape
        if attention_mask is not None:
            # Attention.prepare_attention_mask: [B,Lq,Lk] -> [B*H,Lq,Lk] -> view [B,H,Lq,Lk]
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
            attention_mask = attention_mask.view(b, self.heads, -1, attention_mask.shape[-1])
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.to_q(hidden_states)
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        hd = self.inner_dim // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.inner_dim).to(q.dtype)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        return o


class _GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("not on the hot path")


def _get_activation(name):
    name = name.lower()
    return {"silu": nn.SiLU, "swish": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}[name]()


class _DictConfig(dict):
    def __init__(self, content=None, **kw):
        super().__init__(content or {}, **kw)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    # matcha.utils pulls hydra/lightning/gdown/... -> stub the package, keep the real subpackages
    mu = _mod("matcha.utils")
    mu.__path__ = []
    _mod("matcha.utils.pylogger", get_pylogger=lambda name=__name__: logging.getLogger(name))
    _mod("conformer", ConformerBlock=type("ConformerBlock", (nn.Module,), {}))
    _mod("omegaconf", DictConfig=_DictConfig)
    d = _mod("diffusers")
    d.__path__ = []
    _mod("diffusers.models").__path__ = []
    _mod("diffusers.utils").__path__ = []
    _mod("diffusers.models.activations", get_activation=_get_activation)
    _mod("diffusers.models.attention", GELU=_GELU, GEGLU=_Unused, ApproximateGELU=_Unused,
         AdaLayerNorm=_Unused, AdaLayerNormZero=_Unused)
    _mod("diffusers.models.attention_processor", Attention=_Attention)
    _mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    _mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)

    # dac-vae: stub the vendored audiotools (needs flatten_dict, julius, soundfile ...)
    class BaseModel(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

    ml = _mod("audiotools.ml", BaseModel=BaseModel)
    at = _mod("audiotools", AudioSignal=object, STFTParams=object, ml=ml)
    at.__path__ = []

    sys.path.insert(0, os.path.join(REF_ROOT, "speech"))
    sys.path.insert(0, os.path.join(REF_ROOT, "dac-vae"))
    _installed = True


def load_flow_classes():
    """-> (CausalConditionalCFM, ConditionalCFM, CausalConditionalDecoder, ConditionalDecoder, DictConfig)"""
    install_stubs()
    from cosyvoice.flow.flow_matching import CausalConditionalCFM, ConditionalCFM
    from cosyvoice.flow.decoder import CausalConditionalDecoder, ConditionalDecoder
    return CausalConditionalCFM, ConditionalCFM, CausalConditionalDecoder, ConditionalDecoder, _DictConfig


def load_dac_module():
    install_stubs()
    import model as dac_model  # dac-vae/model.py
    return dac_model


CFM_PARAMS = dict(sigma_min=1e-06, solver="euler", t_scheduler="cosine", training_cfg_rate=0.2,
                  inference_cfg_rate=0.7, reg_loss_type="l1", use_immiscible=True, immiscible_k=8,
                  use_contrastive_fm=True, contrastive_lambda=0.05)  # speech/config.yaml:92-104

DAC_CFG_X2 = dict(sample_rate=24000, encoder_dim=64, latent_dim=80, encoder_rates=[2, 3, 4, 4, 5],
                  decoder_dim=1536, decoder_rates=[5, 4, 4, 3, 2], d_in=1, d_out=1, weight_init="xavier",
                  activation="snake", gain=1.0)  # dac-vae/configs/configx2.yml:2-13


def build_reference_flow(estimator_kwargs=None):
    """Reference CausalConditionalCFM + CausalConditionalDecoder per speech/config.yaml:89-116."""
    CCFM, _, CDec, _, DC = load_flow_classes()
    kw = dict(in_channels=320, out_channels=80, channels=[256], dropout=0.0, attention_head_dim=64,
              n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn="gelu", static_chunk_size=50,
              num_decoding_left_chunks=-1)
    kw.update(estimator_kwargs or {})
    est = CDec(**kw)
    cfm = CCFM(in_channels=240, n_spks=1, spk_emb_dim=80, cfm_params=DC(content=CFM_PARAMS), estimator=est)
    return cfm.eval()


def build_reference_noncausal_estimator():
    """The reference's non-causal ConditionalDecoder (speech/cosyvoice/flow/decoder.py:88-291) at config.yaml's geometry
    (channels=[256]: no down/up-sampling level)."""
    _, _, _, Dec, _ = load_flow_classes()
    return Dec(in_channels=320, out_channels=80, channels=[256], dropout=0.0, attention_head_dim=64, n_blocks=4,
               num_mid_blocks=12, num_heads=8, act_fn="gelu").eval()


def load_fsq_codebook_class():
    """s3tokenizer.model_v2.FSQCodebook (speech/tools/S3Tokenizer); onnx / torchaudio, imported by its siblings, are stubbed."""
    import importlib
    import types
    root = os.path.join(REF_ROOT, "speech", "tools", "S3Tokenizer")
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("onnx", "onnx.numpy_helper", "torchaudio", "torchaudio.compliance", "torchaudio.compliance.kaldi", "tqdm"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    return importlib.import_module("s3tokenizer.model_v2").FSQCodebook



def build_reference_s3_tokenizer(n_mels=128, n_state=1280, n_head=20, n_layer=6):
    """The unmodified S3TokenizerV2 (speech/tools/S3Tokenizer/s3tokenizer/model_v2.py:354-415), same stubs as above."""
    load_fsq_codebook_class()
    import importlib
    m2 = importlib.import_module("s3tokenizer.model_v2")
    cfg = m2.ModelConfig(n_mels=n_mels, n_audio_state=n_state, n_audio_head=n_head, n_audio_layer=n_layer)
    return m2.S3TokenizerV2("speech_tokenizer_v2_25hz", cfg).eval()

def build_reference_dac(cfg=None):
    dm = load_dac_module()
    return dm.DACVAE(**(cfg or DAC_CFG_X2)).eval()


def build_reference_speaker_encoder():
    """The unmodified LearnableSpeakerEncoder (speech/cosyvoice/llm/llm.py:34-96); cosyvoice.transformer.xtransformers
    (only used by other classes of arch_util) is stubbed."""
    import importlib
    import types
    install_stubs()
    if "cosyvoice.transformer.xtransformers" not in sys.modules:
        xt = types.ModuleType("cosyvoice.transformer.xtransformers")

        class _Absent:
            def __init__(self, *a, **k):
                raise RuntimeError("xtransformers is stubbed")

        xt.ContinuousTransformerWrapper = _Absent
        xt.RelativePositionBias = _Absent
        sys.modules["cosyvoice.transformer.xtransformers"] = xt
    llm = importlib.import_module("cosyvoice.llm.llm")
    return llm.LearnableSpeakerEncoder(mel_dim=80, model_dim=512, output_dim=192, num_blocks=6, num_heads=8).eval()


def build_reference_pipeline(estimator_kwargs=None):
    """The unmodified CausalMaskedDiffWithXvec (speech/cosyvoice/flow/flow.py:201-511) per speech/config.yaml:61-116 with its
    own UpsampleConformerEncoder, CausalConditionalCFM and LearnableSpeakerEncoder (use_speaker_encoder=True)."""
    import importlib
    build_reference_speaker_encoder()  # installs the xtransformers stub before cosyvoice.llm.llm is imported
    um = importlib.import_module("cosyvoice.transformer.upsample_encoder")
    fl = importlib.import_module("cosyvoice.flow.flow")
    enc = um.UpsampleConformerEncoder(
        input_size=512, output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
        positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True, input_layer="linear",
        pos_enc_layer_type="rel_pos_espnet", selfattention_layer_type="rel_selfattn", use_cnn_module=False,
        macaron_style=False, static_chunk_size=25)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints its config
        m = fl.CausalMaskedDiffWithXvec(input_size=512, output_size=80, spk_embed_dim=192, output_type="mel", vocab_size=6561,
                                        input_frame_rate=25, only_mask_loss=True, token_latent_ratio=2, pre_lookahead_len=3,
                                        use_speaker_encoder=True, encoder=enc, decoder=build_reference_flow(estimator_kwargs))
    return m.eval()
