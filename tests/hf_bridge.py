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
