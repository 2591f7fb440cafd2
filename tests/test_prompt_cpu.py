"""Prompt assembly against fixtures recorded from the REFERENCE's own
`_build_talker_inputs_local` (model.py:583-805) executed on the synthetic module tree (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import prompt_cases as PC

from faster_qwen3_tts.prompt import build_talker_inputs  # noqa: E402  (path set up by prompt_cases)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "prompt.npz")


@pytest.fixture(scope="module")
def base():
    return PC.build_base(seed=0)


def test_every_reference_case_is_reproduced(base):
    gold = np.load(GOLD)
    cases = PC.cases(base)
    assert sorted(cases) == sorted(gold["names"].tolist())
    with torch.inference_mode():
        for name, kw in cases.items():
            tie, tam, tth, tpe = build_talker_inputs(base.model, **kw)
            assert tuple(tie.shape) == gold[name + "_tie"].shape, name
            assert np.array_equal(tam.numpy(), gold[name + "_tam"]), name          # left padding, bit for bit
            assert tam.dtype == torch.long
            # same fp32 ops on the same rows; only the batching of the text projection differs -> float tolerance 1e-6
            for got, key in ((tie, "_tie"), (tth, "_tth"), (tpe, "_tpe")):
                want = gold[name + key]
                assert got.shape == want.shape, (name, key)
                assert np.abs(got.numpy() - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), (name, key)


def test_left_padding_and_trailing_fill(base):
    kw = PC.cases(base)["custom_batch2_left_pad"]
    with torch.inference_mode():
        tie, tam, tth, tpe = build_talker_inputs(base.model, **kw)
    pads = (tam == 0).sum(1).tolist()
    assert pads[0] > 0 and pads[1] == 0
    assert torch.all(tie[0, : pads[0]] == 0)                     # zero embeddings under mask 0
    assert torch.equal(tth[0, -1], tpe.reshape(-1))              # shorter trailing stream is filled with tts_pad


def test_unknown_speaker_and_language_raise_like_the_reference(base):
    kw = dict(PC.cases(base)["custom_speaker_stream"])
    with pytest.raises(NotImplementedError, match="Speaker nobody not implemented"):
        build_talker_inputs(base.model, **{**kw, "speakers": ["nobody"]})
    with pytest.raises(NotImplementedError, match="Language Klingon not implemented"):
        build_talker_inputs(base.model, **{**kw, "languages": ["Klingon"]})


def test_tokenizer_layout_matches_the_slices_the_reference_takes(base):
    ids = base._tokenize_texts([base._build_assistant_text("hello there")])[0]
    assert ids[0, :3].tolist() == [1, 4, 3] and ids[0, -5:].tolist() == [2, 3, 1, 4, 3]
    ref = base._tokenize_texts([base._build_ref_text("hello there")])[0]
    assert ref[0, :3].tolist() == [1, 4, 3] and ref[0, -2:].tolist() == [2, 3]
    assert torch.equal(ids[0, 3:-5], ref[0, 3:-2])
