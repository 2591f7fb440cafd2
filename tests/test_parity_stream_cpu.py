"""parity_mode streaming (the reference's dynamic-cache baseline, streaming.py:192-359): the product's restatement,
driven with the decode-step talker double of oracle/parity_cases.py, against the fixture recorded by EXECUTING the
reference's own ``parity_generate_streaming`` with the same double (oracle/make_golden.py -> tests/golden/parity_stream.npz):
same codes, same chunk boundaries, same is_final / total_steps_so_far sequences, same number of talker calls and the
same last call (token, mask length, cache position, subtalker kwargs), same timing keys."""
import os

import numpy as np
import pytest
import torch

from oracle import parity_cases as PC


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "parity_stream.npz"), allow_pickle=False)


@pytest.mark.parametrize("case", PC.CASES, ids=[c[0] for c in PC.CASES])
def test_parity_stream_matches_reference_recording(golden, case):
    from faster_qwen3_tts.streaming import parity_generate_streaming
    chunks, timings, calls = PC.run_case(parity_generate_streaming, case)
    name = case[0]
    codes = torch.cat(chunks).numpy() if chunks else np.zeros((0, 16), dtype=np.int64)
    assert np.array_equal(codes, golden[name + "_codes"])
    assert [c.shape[0] for c in chunks] == golden[name + "_chunks"].tolist()
    assert [int(t["is_final"]) for t in timings] == golden[name + "_final"].tolist()
    assert [t["total_steps_so_far"] for t in timings] == golden[name + "_total"].tolist()
    assert len(calls) == int(golden[name + "_ncalls"][0])
    assert str(calls[-1]) == str(golden[name + "_lastcall"][0])
    if timings:
        assert sorted(timings[0].keys()) == golden["timing_keys"].tolist()
        assert timings[0]["prefill_ms"] >= 0 and all(t["prefill_ms"] == 0 for t in timings[1:])


def test_parity_mode_needs_the_upstream_decode_step():
    """the offline synthetic talker is prefill-only: parity_mode says so instead of silently taking the fast path"""
    from faster_qwen3_tts.streaming import parity_generate_streaming

    class PrefillOnly:
        pass

    with pytest.raises(NotImplementedError):
        next(parity_generate_streaming(PrefillOnly(), torch.zeros(1, 3, 4), torch.ones(1, 3, dtype=torch.long),
                                       torch.zeros(1, 1, 4), torch.zeros(1, 1, 4), None))
