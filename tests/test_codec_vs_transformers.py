"""The codec decoder module (faster_qwen3_tts/codec.py `Code2Wav`, the torch reference the K4 kernels are checked
against) versus the Hugging Face `Qwen3OmniMoeCode2Wav` of the in-image transformers -- the public analogue of the
Qwen3-TTS 12 Hz tokenizer decoder that SURVEY 8(c) names (code-offset embedding mean -> sliding-window pre-transformer
with layer scale -> 2 x (ConvTranspose k=2 + ConvNeXt) -> conv7 -> 4 x [SnakeBeta, causal ConvTranspose(2r, r),
3 residual units] -> SnakeBeta -> conv7 -> clamp).  Same weights (copied by name), same codes, fp32: the two
implementations must agree to float rounding.

One deliberate difference: the analogue's `CausalTransConvNet` trims `kernel - stride` samples on BOTH sides, which
yields (T-1)*r samples per stage; the Qwen3-TTS decoder the reference drives returns exactly 1920 samples per frame
(model.py:935-937 relies on it, every committed sample wav is a multiple of 1920), i.e. the causal trim is on the right
only -- which is what codec.py and the K4 kernels implement.  The test patches the analogue's trim accordingly and
compares everything else."""
import pytest
import torch

pytest.importorskip("transformers")

from oracle import prompt_cases  # noqa: F401,E402  (puts the package on sys.path)

from faster_qwen3_tts import codec  # noqa: E402


def _name_map(n_layers):
    m = {"code_embedding.weight": "code_embedding.weight", "norm.weight": "pre_transformer.norm.weight"}
    for i in range(n_layers):
        a, b = f"layers.{i}.", f"pre_transformer.layers.{i}."
        for mine, hf in (("q", "self_attn.q_proj"), ("k", "self_attn.k_proj"), ("v", "self_attn.v_proj"), ("o", "self_attn.o_proj"),
                         ("gate", "mlp.gate_proj"), ("up", "mlp.up_proj"), ("down", "mlp.down_proj"),
                         ("ln1", "input_layernorm"), ("ln2", "post_attention_layernorm")):
            m[a + mine + ".weight"] = b + hf + ".weight"
        m[a + "s1"] = b + "self_attn_layer_scale.scale"
        m[a + "s2"] = b + "mlp_layer_scale.scale"
    return m


@pytest.mark.parametrize("T", [1, 7, 90])   # 90 > sliding_window 72: the window mask matters
def test_code2wav_matches_hf_analogue(T, monkeypatch):
    from transformers.models.qwen3_omni_moe import configuration_qwen3_omni_moe as Cf, modeling_qwen3_omni_moe as M
    kw = dict(codebook_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
              decoder_dim=256)
    hcfg = Cf.Qwen3OmniMoeCode2WavConfig(num_key_value_heads=4, num_quantizers=16, sliding_window=72,
                                         upsample_rates=(8, 5, 4, 3), upsampling_ratios=(2, 2), **kw)
    hcfg._attn_implementation = "eager"

    def right_trim_forward(self, hidden_state):   # causal: keep the first T*stride samples
        hidden_state = self.conv(hidden_state)
        return hidden_state[..., : hidden_state.shape[-1] - self.right_pad].contiguous()

    monkeypatch.setattr(M.Qwen3OmniMoeCausalTransConvNet, "forward", right_trim_forward)
    torch.manual_seed(0)
    hf = M.Qwen3OmniMoeCode2Wav(hcfg).eval()
    with torch.no_grad():   # variance-preserving random weights so that a real signal reaches the output
        g = torch.Generator().manual_seed(1)
        for n, p in hf.named_parameters():
            if n.endswith((".alpha", ".beta")):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))            # SnakeBeta exponents
            elif n.endswith(".scale") or n.endswith(".gamma"):
                p.copy_(0.5 + 0.2 * torch.randn(p.shape, generator=g))      # layer scales
            elif p.dim() == 1:
                p.copy_((1.0 if "norm" in n else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
            elif "code_embedding" in n:
                p.copy_(torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(0.55 * torch.randn(p.shape, generator=g) / (fan_in ** 0.5))   # gain ~1 per stage
        peak = hf(torch.randint(0, 64, (1, 16, 8), generator=g)).abs().max()   # keep the output inside the clamp
    mine = codec.Code2Wav(codec.Code2WavConfig(**kw)).eval()
    hp = dict(hf.named_parameters())
    nm = _name_map(2)
    rest_mine = [n for n, _ in mine.named_parameters() if n not in nm]
    rest_hf = [n for n in hp if n not in nm.values()]
    assert len(rest_mine) == len(rest_hf)
    nm.update(dict(zip(rest_mine, rest_hf)))     # the convolutional part is declared in the same order on both sides
    with torch.no_grad():
        for n, p in mine.named_parameters():
            assert p.shape == hp[nm[n]].shape, (n, nm[n])
            p.copy_(hp[nm[n]])
    codes = torch.randint(0, 64, (1, 16, T), generator=torch.Generator().manual_seed(T))
    with torch.inference_mode():
        want = hf(codes)
        got = mine(codes)
    assert want.shape[-1] == 1920 * T and got.reshape(-1).shape[0] == 1920 * T
    err = (got.reshape(-1) - want.reshape(-1)).abs().max().item()
    print(f"T={T}: max abs PCM diff {err:.2e} (|pcm| max {want.abs().max():.3f})")
    assert want.abs().max() > 0.05 and err < 5e-6   # O(1) activations at every stage: float rounding only
    # the functional fp32 ORACLE (oracle/codec_oracle.py, the checker of the 1e-3 PCM bar) on the same weights
    from oracle import codec_oracle
    with torch.inference_mode():
        ora = codec_oracle.decode(mine.state_dict(), codes[0].transpose(0, 1), codebook_size=64, num_attention_heads=4)
    err_o = (ora - want.reshape(-1)).abs().max().item()
    print(f"T={T}: oracle vs HF analogue max abs PCM diff {err_o:.2e}")
    assert ora.shape[0] == 1920 * T and err_o < 5e-6
