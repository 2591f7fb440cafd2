"""Shared helpers for the GPU parity tests: one set of seeded weights -> (CPU oracle, fq3 engine)."""
import types

import numpy as np
import torch

from oracle import qwen3_tts_oracle as O


def syn_cfg_from(cfg: O.ModelCfg):
    def stack(c: O.StackCfg):
        return types.SimpleNamespace(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                                     num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                                     num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim,
                                     rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta, vocab_size=c.vocab_size,
                                     sliding_window=None)
    t = stack(cfg.talker)
    t.codec_eos_token_id = cfg.codec_eos_token_id
    t.num_code_groups = cfg.num_code_groups
    p = stack(cfg.predictor)
    p.num_code_groups = cfg.num_code_groups
    return types.SimpleNamespace(talker_config=t, code_predictor_config=p, has_mtp=cfg.has_mtp_projection)


class Pair:
    def __init__(self, cfg: O.ModelCfg, seed=0, dtype=torch.float32, max_seq_len=256, eos_boost=1.0, num_ctas=0,
                 max_batch=1, oracle_device="cpu"):
        from faster_qwen3_tts import synthetic
        from faster_qwen3_tts.predictor_graph import PredictorGraph
        from faster_qwen3_tts.talker_graph import TalkerGraph
        from faster_qwen3_tts.weights import engine_for_talker
        self.cfg = cfg
        self.dtype = dtype
        self.W = O.make_weights(cfg, seed=seed, dtype=dtype, eos_boost=eos_boost)
        # the oracle is plain torch: full-size cases host it on the accelerator (torch eager = the reference's own
        # arithmetic on this GPU) so that they finish in seconds
        self.om = O.OracleModel(cfg, self.W if oracle_device == "cpu" else {k: v.to(oracle_device) for k, v in self.W.items()})
        self.syn = syn_cfg_from(cfg)
        self.base = synthetic.build_base_model(self.syn, self.W, dtype=dtype, device="cuda")
        self.talker = self.base.model.talker
        self.engine = engine_for_talker(self.talker, dtype=dtype, device="cuda", max_seq_len=max_seq_len,
                                        num_ctas=num_ctas, max_batch=max_batch)
        self.pg = PredictorGraph(self.talker.code_predictor, self.syn.code_predictor_config, cfg.talker.hidden_size,
                                 dtype=dtype, engine=self.engine)
        self.tg = TalkerGraph(self.talker.model, self.syn.talker_config, dtype=dtype, max_seq_len=max_seq_len,
                              engine=self.engine)
        self.config = self.syn.talker_config


def report(name, a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    print(f"  {name:14s} max|d|={d.max().item():.3e} max|ref|={b.abs().max().item():.3e} "
          f"argmax={int(d.argmax())} of {a.numel()}", flush=True)
    return d.max().item()
