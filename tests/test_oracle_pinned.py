"""The CPU oracle against fixtures produced by EXECUTING the reference's own sampling.py / generate.py /
streaming.py (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import qwen3_tts_oracle as O


@pytest.fixture(scope="module")
def samp(golden_dir):
    return np.load(os.path.join(golden_dir, "sampling.npz"))


@pytest.fixture(scope="module")
def loop(golden_dir):
    return np.load(os.path.join(golden_dir, "loop.npz"), allow_pickle=False)


def _case(samp, i):
    pre = f"c{i}_"
    dt = torch.bfloat16 if int(samp[pre + "bf16"]) else torch.float32
    lg = torch.from_numpy(samp[pre + "logits"]).to(dt)
    T, k, p, u, eos, sup = samp[pre + "params"]
    V = lg.numel()
    mask = torch.zeros(V, dtype=torch.bool)
    mask[(V - 1024 if V > 1024 else V - 32):] = True
    mask[int(eos)] = False
    return lg, dict(temperature=float(T), top_k=int(k), top_p=float(p)), float(u), mask, \
        (None if sup < 0 else [int(sup)]), pre


def test_filtered_probs_bit_exact_vs_reference(samp):
    for i in range(int(samp["n_cases"])):
        lg, kw, u, mask, sup, pre = _case(samp, i)
        pr = O.filtered_probs(lg, suppress_mask=mask, suppress_tokens=sup, **kw)
        ref = samp[pre + "probs"]
        assert np.array_equal(pr.float().numpy(), ref), f"case {i}"


def test_sample_token_vs_reference_with_noise_contract(samp):
    for i in range(int(samp["n_cases"])):
        lg, kw, u, mask, sup, pre = _case(samp, i)
        tok = O.sample_token(lg, do_sample=True, u=u, suppress_mask=mask, suppress_tokens=sup, **kw)
        assert tok == int(samp[pre + "token"]), f"case {i}"
        g = O.sample_token(lg, do_sample=False, u=u, suppress_mask=mask, suppress_tokens=sup, **kw)
        assert g == int(samp[pre + "greedy"]), f"case {i}"


def test_repetition_penalty_known_answers(samp):
    # the reference's own KAT (tests/test_sampling.py:10-21)
    logits = torch.zeros(1, 1, 10)
    logits[..., 7] = 1.0
    logits[..., 8] = -1.0
    others = [0, 1, 2, 3, 4, 5, 6, 8, 9]
    hist = torch.tensor([7] + [others[i % len(others)] for i in range(1, 60)])
    out = O.apply_repetition_penalty(logits.clone(), hist, 1.1)
    assert out[0, 0, 7].item() == pytest.approx(1.0 / 1.1, rel=1e-6)
    assert out[0, 0, 8].item() == pytest.approx(-1.1, rel=1e-6)
    for j, dt in enumerate((torch.float32, torch.bfloat16)):
        lg = torch.from_numpy(samp[f"pen{j}_logits"]).to(dt)
        o = O.apply_repetition_penalty(lg.clone(), torch.from_numpy(samp[f"pen{j}_hist"]), 1.05)
        assert np.array_equal(o.float().numpy(), samp[f"pen{j}_out"])


def test_inverse_cdf_edges():
    p = torch.tensor([0.0, 0.25, 0.0, 0.75])
    assert O.draw_inverse_cdf(p, 0.0) == 1
    assert O.draw_inverse_cdf(p, 0.2499) == 1
    assert O.draw_inverse_cdf(p, 0.25) == 3
    assert O.draw_inverse_cdf(p, 0.99999994) == 3


@pytest.mark.parametrize("idx", range(7))
def test_loop_matches_reference_schedulers(loop, idx):
    name = str(loop["names"][idx])
    wseed, P, Tt, max_new, min_new, do_sample, pen, max_seq, chunk, boost, nseed = loop[name + "_params"]
    cfg = O.cfg_tiny()
    W = O.make_weights(cfg, seed=int(wseed), eos_boost=float(boost))
    om = O.OracleModel(cfg, W)
    tie, tth, tpe = O.make_inputs(cfg, int(P), int(Tt), seed=int(wseed))
    uniforms = np.random.default_rng(int(nseed)).random((int(max_new) + 1, 16), dtype=np.float32)
    with torch.inference_mode():
        codes, chunks = O.generate(
            om, tie, tth, tpe, max_new_tokens=int(max_new), min_new_tokens=int(min_new),
            sp_talker=O.SamplingParams(do_sample=bool(do_sample), repetition_penalty=float(pen)),
            sp_pred=O.SamplingParams(do_sample=bool(do_sample), repetition_penalty=1.0),
            max_seq_len=int(max_seq), uniforms=uniforms, chunk_size=int(chunk))
    assert np.array_equal(codes.numpy(), loop[name + "_codes"]), name
    assert chunks == loop[name + "_chunks"].tolist(), name
