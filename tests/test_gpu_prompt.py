"""Prompt assembly ON THE DEVICE the request runs on: the same nine cases recorded from the
reference's own `_build_talker_inputs_local` (tests/golden/prompt.npz), with the synthetic module tree and every input
tensor on cuda:0 -- the configuration the e2e bench leg and the servers use (gather / projection / adds are device ops
feeding fq3_prefill directly, no host round trip)."""
import os

import numpy as np
import pytest
import torch

from oracle import prompt_cases as PC

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "prompt.npz")


def _to(x, dev):
    if torch.is_tensor(x):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: _to(v, dev) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to(v, dev) for v in x)
    return x


def test_prompt_assembly_on_device_matches_reference_recording():
    from faster_qwen3_tts.prompt import build_talker_inputs
    torch.backends.cuda.matmul.allow_tf32 = False
    base = PC.build_base(seed=0)
    cases = PC.cases(base)          # ids / voice prompts built on the host exactly like the fixture generator did
    base.model.talker.to("cuda")
    gold = np.load(GOLD)
    with torch.inference_mode():
        for name, kw in cases.items():
            tie, tam, tth, tpe = build_talker_inputs(base.model, **_to(kw, "cuda"))
            assert tie.is_cuda and tam.is_cuda and tth.is_cuda and tpe.is_cuda, name
            assert np.array_equal(tam.cpu().numpy(), gold[name + "_tam"]), name
            for got, key in ((tie, "_tie"), (tth, "_tth"), (tpe, "_tpe")):
                want = gold[name + key]
                assert tuple(got.shape) == want.shape, (name, key)
                assert np.abs(got.float().cpu().numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), (name, key)
