"""CPU, build container only: oracle vs the UNMODIFIED reference imported from /root/reference (skipped elsewhere)."""
import dataclasses

import pytest
import torch

from oracle import reference_shim as R
from oracle import routeformer_oracle as O

pytestmark = pytest.mark.skipif(not R.available(), reason="reference not mounted (GPU box)")

SMALL = O.OracleConfig(d_model=64, n_heads=4, e_layers=3, d_ff=128, with_video=True, with_gaze=True,
                       dense_prediction=True, encoder_layers=2, encoder_d_ff=64, image_embedding_size=32,
                       encoder_hidden_size=32)
SPEC = O.BackboneSpec(image_size=32, patch=8, channels=48)


@pytest.mark.parametrize("B,seed", [(1, 1), (2, 1), (3, 2), (4, 3)])
def test_forward_bit_level(B, seed):
    model = R.build_reference_model(SMALL, SPEC).eval()
    sd = O.fill_state_dict(O.state_dict_template(SMALL, SPEC), 7 + seed)
    model.load_state_dict(sd)
    batch = O.synthetic_batch(B, SMALL, "tiny", seed=seed)
    torch.manual_seed(12345)
    with torch.no_grad():
        m_ref, v_ref = model.preprocess_batch(batch)
        torch.manual_seed(12345)
        wp_ref, dense_ref = model(batch)
    orc = O.Routeformer(sd, SMALL, SPEC)
    torch.manual_seed(12345)
    m, v = orc.preprocess(batch, False, O.CpuRandint())
    assert torch.equal(m, m_ref) and torch.equal(v, v_ref)  # the visual path is bit-identical
    torch.manual_seed(12345)
    wp, dense = orc.forward(batch)
    assert (wp - wp_ref).abs().max() <= 1e-6 * wp_ref.abs().max()
    assert (dense - dense_ref).abs().max() <= 2e-5 * dense_ref.abs().max()


def test_view_and_gaze_dropout_rng_order():
    """Training-mode torch.rand(1) draws interleave with the randint draws exactly as in routeformer.py:301,405-408."""
    cfg = dataclasses.replace(SMALL, view_dropout=0.6, gaze_dropout=0.2)
    model = R.build_reference_model(cfg, SPEC).train()
    sd = O.fill_state_dict(O.state_dict_template(cfg, SPEC), 3)
    model.load_state_dict(sd)
    batch = O.synthetic_batch(2, cfg, "tiny", seed=5)
    for s in range(6):  # different seeds exercise drop-left / drop-right / drop-gaze / none
        torch.manual_seed(s)
        wp_ref, _ = model(batch)
        torch.manual_seed(s)
        wp, _ = O.Routeformer(sd, cfg, SPEC).forward(batch, training=True)
        assert (wp - wp_ref).abs().max() <= 1e-5 * wp_ref.abs().max(), s


def test_target_pass_shapes():
    """preprocess_batch(target, training=False) on the 30 target frames (full_comparison.py:482)."""
    model = R.build_reference_model(SMALL, SPEC).eval()
    sd = O.fill_state_dict(O.state_dict_template(SMALL, SPEC), 4)
    model.load_state_dict(sd)
    tgt = O.synthetic_batch(2, SMALL, "tiny", seed=8, T=30)
    torch.manual_seed(1)
    with torch.no_grad():
        _, v_ref = model.preprocess_batch(tgt, training=False)
    torch.manual_seed(1)
    _, v = O.Routeformer(sd, SMALL, SPEC).preprocess(tgt, False, O.CpuRandint())
    assert v.shape == v_ref.shape and torch.equal(v, v_ref)
