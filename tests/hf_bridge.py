"""Map MLX-layout Whisper weights onto the independent `transformers` implementation (test helper).

SURVEY.md section 8c: the HF model is the second, executable source that pins the oracle.  Name /
layout differences are listed there (conv weight (out,k,in) vs (out,in,k); attn.query vs
self_attn.q_proj; ...).
"""
import torch


def build_hf(dims: dict, w: dict):
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    cfg = WhisperConfig(
        vocab_size=dims["n_vocab"], num_mel_bins=dims["n_mels"], d_model=dims["n_audio_state"],
        encoder_layers=dims["n_audio_layer"], encoder_attention_heads=dims["n_audio_head"],
        decoder_layers=dims["n_text_layer"], decoder_attention_heads=dims["n_text_head"],
        encoder_ffn_dim=4 * dims["n_audio_state"], decoder_ffn_dim=4 * dims["n_text_state"],
        max_source_positions=dims["n_audio_ctx"], max_target_positions=dims["n_text_ctx"],
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        tie_word_embeddings=True,
    )
    cfg._attn_implementation = "eager"
    hf = WhisperForConditionalGeneration(cfg).eval().float()
    sd = {}

    def attn(src, dst):
        for a, b in (("query", "q_proj"), ("key", "k_proj"), ("value", "v_proj"), ("out", "out_proj")):
            sd[f"{dst}.{b}.weight"] = w[f"{src}.{a}.weight"]
            if a != "key":
                sd[f"{dst}.{b}.bias"] = w[f"{src}.{a}.bias"]

    def ln(src, dst):
        sd[dst + ".weight"] = w[src + ".weight"]
        sd[dst + ".bias"] = w[src + ".bias"]

    sd["model.encoder.conv1.weight"] = w["encoder.conv1.weight"].permute(0, 2, 1)
    sd["model.encoder.conv1.bias"] = w["encoder.conv1.bias"]
    sd["model.encoder.conv2.weight"] = w["encoder.conv2.weight"].permute(0, 2, 1)
    sd["model.encoder.conv2.bias"] = w["encoder.conv2.bias"]
    ln("encoder.ln_post", "model.encoder.layer_norm")
    for i in range(dims["n_audio_layer"]):
        s, d = f"encoder.blocks.{i}", f"model.encoder.layers.{i}"
        attn(s + ".attn", d + ".self_attn")
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        ln(s + ".mlp1", d + ".fc1")
        ln(s + ".mlp2", d + ".fc2")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
    sd["model.decoder.embed_tokens.weight"] = w["decoder.token_embedding.weight"]
    sd["model.decoder.embed_positions.weight"] = w["decoder.positional_embedding"]
    ln("decoder.ln", "model.decoder.layer_norm")
    for i in range(dims["n_text_layer"]):
        s, d = f"decoder.blocks.{i}", f"model.decoder.layers.{i}"
        attn(s + ".attn", d + ".self_attn")
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        attn(s + ".cross_attn", d + ".encoder_attn")
        ln(s + ".cross_attn_ln", d + ".encoder_attn_layer_norm")
        ln(s + ".mlp1", d + ".fc1")
        ln(s + ".mlp2", d + ".fc2")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
    missing, unexpected = hf.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    missing = [m for m in missing if "embed_positions" not in m and "proj_out" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


def hf_to_mlx(hf_state_dict: dict, dtype=torch.float16) -> dict:
    """transformers Whisper state dict -> MLX-layout weights, following the PUBLISHED conversion recipe of
    mlx-examples `whisper/convert.py` (the script that produced the `mlx-community/whisper-*-mlx` repos): HF names are
    first mapped onto the OpenAI names (`layers` -> `blocks`, `fc1`/`fc2` -> `mlp.0`/`mlp.2`, `final_layer_norm` ->
    `mlp_ln`, `self_attn.q_proj` -> `attn.query`, `encoder_attn` -> `cross_attn`, `embed_tokens` ->
    `token_embedding`, `embed_positions.weight` -> `positional_embedding`, `layer_norm` -> `ln_post` / `ln`), then
    OpenAI -> MLX (`mlp.0` -> `mlp1`, `mlp.2` -> `mlp2`; conv weights (out, in, k) -> (out, k, in)); the tied
    `proj_out` and the encoder's fixed sinusoid table are dropped (MLX recomputes the sinusoids).
    Written from that description, NOT from whisper-mlx_b200/load_models.py: this is the independent side of the
    loader test."""
    rules = [("model.", ""), (".layers.", ".blocks."), (".fc1.", ".mlp.0."), (".fc2.", ".mlp.2."),
             (".final_layer_norm.", ".mlp_ln."), (".self_attn.q_proj.", ".attn.query."), (".self_attn.k_proj.", ".attn.key."),
             (".self_attn.v_proj.", ".attn.value."), (".self_attn_layer_norm.", ".attn_ln."),
             (".self_attn.out_proj.", ".attn.out."), (".encoder_attn.q_proj.", ".cross_attn.query."),
             (".encoder_attn.k_proj.", ".cross_attn.key."), (".encoder_attn.v_proj.", ".cross_attn.value."),
             (".encoder_attn_layer_norm.", ".cross_attn_ln."), (".encoder_attn.out_proj.", ".cross_attn.out."),
             ("decoder.layer_norm.", "decoder.ln."), ("encoder.layer_norm.", "encoder.ln_post."),
             ("embed_tokens", "token_embedding"), ("decoder.embed_positions.weight", "decoder.positional_embedding"),
             (".mlp.0.", ".mlp1."), (".mlp.2.", ".mlp2.")]
    out = {}
    for k, v in hf_state_dict.items():
        if k.startswith("proj_out") or k == "model.encoder.embed_positions.weight":
            continue
        for a, b in rules:
            k = k.replace(a, b)
        if k in ("encoder.conv1.weight", "encoder.conv2.weight"):
            v = v.permute(0, 2, 1)
        out[k] = v.detach().to(dtype).contiguous()
    return out


def mlx_quantize_dict(weights: dict, group_size: int = 64, bits: int = 4) -> dict:
    """What `nn.quantize(model, group_size, bits)` does to a saved MLX Whisper: every Linear / Embedding `X.weight`
    becomes `X.weight` (uint32 words, `32 / bits` values each, little end first), `X.scales`, `X.biases`
    ((out, in / group_size), the checkpoint dtype), with w ~ q * scale + bias, scale = (max - min) / (2^bits - 1),
    bias = min per group (mx.quantize).  Conv weights, norms, biases and the positional embedding stay as they are."""
    out = {}
    for k, w in weights.items():
        is_linear = k.endswith(".weight") and w.ndim == 2 and ("_ln." not in k and ".ln" not in k) and w.shape[1] % group_size == 0
        if not is_linear:
            out[k] = w
            continue
        o, i = w.shape
        g = w.float().view(o, i // group_size, group_size)
        lo, hi = g.min(-1).values, g.max(-1).values
        scale = ((hi - lo) / (2 ** bits - 1)).clamp_min(1e-8)
        q = torch.round((g - lo[..., None]) / scale[..., None]).clamp(0, 2 ** bits - 1).to(torch.int64).view(o, i)
        per = 32 // bits
        words = (q.view(o, i // per, per) << (torch.arange(per) * bits)).sum(-1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        base = k[: -len(".weight")]
        out[k] = words
        out[base + ".scales"] = scale.to(w.dtype)
        out[base + ".biases"] = lo.to(w.dtype)
    return out
