"""Shared helpers for the parity tests (test infrastructure)."""
import os

import torch

from oracle import routeformer_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def case_from_golden(gold):
    """(cfg, spec, state_dict, batch) regenerated from the seeds stored in a golden file."""
    cfg = O.OracleConfig(**gold["cfg"])
    spec = O.BackboneSpec(**gold["spec"]) if gold["spec"] else None
    sd = O.fill_state_dict(O.state_dict_template(cfg, spec), gold["wseed"])
    batch = O.synthetic_batch(gold["B"], cfg, gold["shapes"], seed=gold["dseed"])
    return cfg, spec, sd, batch


def targets_for(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, cfg.pred_len, 2, generator=g) * 5.0 + 80.0,
            torch.randn(B, cfg.pred_len, cfg.image_embedding_size, generator=g))


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
