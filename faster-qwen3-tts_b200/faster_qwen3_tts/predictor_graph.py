"""PredictorGraph -- same public surface as the reference class (faster_qwen3_tts/predictor_graph.py:23-214).
``run`` executes the whole 15-pass loop (2-token prefill + 14 single-token decodes, per-pass head + sampling) in one
launch of the persistent kernel.  Sampling parameters are attributes, mutable at any time (the reference bakes
them at capture; tests/test_e2e_parity.py:211-214 mutates them before capture)."""
from __future__ import annotations

from typing import Optional

import torch

from .engine import Engine, SamplingParams


class PredictorGraph:
    def __init__(self, code_predictor=None, pred_config=None, talker_hidden_size=None, device="cuda",
                 dtype=torch.bfloat16, do_sample=True, top_k=50, top_p=1.0, temperature=0.9,
                 engine: Optional[Engine] = None):
        self.device = device
        self.dtype = dtype
        self.do_sample = do_sample
        self.top_k = top_k
        self.top_p = top_p
        self.temperature = temperature
        self.num_code_groups = getattr(pred_config, "num_code_groups", 16) if pred_config is not None else 16
        self.num_codebooks = self.num_code_groups - 1
        self.max_seq = 2 + self.num_codebooks
        self.engine = engine
        self.captured = False
        self.slot = 0   # request slot whose predictor cache run() uses
        self.generator: Optional[torch.Generator] = None

    def sampling(self) -> SamplingParams:
        return SamplingParams(do_sample=self.do_sample, top_k=self.top_k, temperature=self.temperature,
                              top_p=self.top_p, repetition_penalty=1.0)

    def _need_engine(self) -> Engine:
        if self.engine is None or not self.engine.loaded:
            raise RuntimeError("PredictorGraph has no loaded fq3 engine: construct it through "
                               "FasterQwen3TTS.from_pretrained(...) or pass engine=Engine(...)")
        return self.engine

    @torch.inference_mode()
    def capture(self, num_warmup=3):
        eng = self._need_engine()
        x = torch.zeros(2, eng.H, dtype=eng.dtype, device=eng.device)
        eng.predictor_run(x, SamplingParams(do_sample=False))
        torch.cuda.synchronize()
        self.captured = True

    @torch.inference_mode()
    def run(self, pred_input: torch.Tensor, uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pred_input [1,2,H_talker] -> LongTensor[15] (fresh tensor)."""
        eng = self._need_engine()
        if self.do_sample and uniforms is None:
            uniforms = torch.rand(self.num_codebooks, device=eng.device, generator=self.generator)
        return eng.predictor_run(pred_input, self.sampling(), uniforms, slot=self.slot)
