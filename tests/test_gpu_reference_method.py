"""The reference-METHOD stand-in that bench.py times beside the engine (baseline/reference_method.py: static KV + mask
table + CUDA graphs + the reference's per-frame eager glue) must itself be a correct generator: greedy tokens equal to
the CPU oracle on the tiny geometry in fp32, with and without left padding."""
import pytest
import torch

from oracle import qwen3_tts_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pad", [0, 5])
def test_reference_method_generates_the_oracle_tokens(pad):
    from util_models import Pair
    from baseline.reference_method import build_reference_method, ref_generate_streaming
    cfg = O.cfg_tiny()
    p = Pair(cfg, seed=2, dtype=torch.float32, max_seq_len=64)
    pg, tg = build_reference_method(p.talker, p.syn, device="cuda", dtype=torch.float32, max_seq_len=64, prefill_len=10)
    pg.do_sample = False
    pg.capture()   # re-capture with greedy sampling baked in (the reference bakes the sampling parameters at capture)
    P, n = 14, 12
    tie, tth, tpe = O.make_inputs(cfg, P, 3, seed=1)
    tam = torch.ones(1, P, dtype=torch.long)
    if pad:
        tie[:pad] = 0
        tam[0, :pad] = 0
    with torch.inference_mode():
        want = O.generate(p.om, tie, tth, tpe, max_new_tokens=n, sp_talker=O.SamplingParams(do_sample=False, repetition_penalty=1.05),
                          sp_pred=O.SamplingParams(do_sample=False), max_seq_len=64, n_left_pad=pad)
    chunks = [c.cpu() for c, t in ref_generate_streaming(p.talker, tie[None].cuda(), tam.cuda(), tth[None].cuda(),
                                                         tpe[None, None].cuda(), p.config, pg, tg, max_new_tokens=n,
                                                         do_sample=False, repetition_penalty=1.05, chunk_size=5)]
    got = torch.cat(chunks)
    print("rows equal:", int((got == want[: got.shape[0]]).all(dim=1).sum()), "of", want.shape[0])
    assert torch.equal(got, want)
    assert [c.shape[0] for c in chunks][:-1] == [5] * (len(chunks) - 1)
