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


def test_random_configurations_against_the_reference():
    """The oracle against the unmodified reference away from the golden configurations: random draws over decoder_mode (vanilla /
    recursive / smart), modalities (GPS only, scene, gaze, both), dense prediction, rotate / normalise, distillation, activation,
    layer / head counts, history and horizon lengths, batch sizes.  Waypoints agree to 1e-6, dense features to 2e-5; where the
    reference itself cannot run a configuration (decoder_mode="recursive" with dense_prediction adds a 37-column input row to a
    34-column output, routeformer.py:244-245) the oracle raises the same RuntimeError."""
    import random

    rnd = random.Random(0)
    ran = raised = 0
    for i in range(16):
        with_video = rnd.random() < 0.8
        with_scene = rnd.random() < 0.7
        with_gaze = with_video and rnd.random() < 0.7
        if with_video and not (with_scene or with_gaze):
            with_gaze = True
        cfg = O.OracleConfig(
            seq_len=rnd.choice([20, 40]), pred_len=rnd.choice([10, 30]), d_model=rnd.choice([32, 64]), n_heads=rnd.choice([2, 4]),
            e_layers=rnd.choice([1, 2, 3]), d_layers=rnd.choice([1, 2]), d_ff=rnd.choice([64, 128]), factor=rnd.choice([3, 4, 5]),
            distil=rnd.random() < 0.7, activation=rnd.choice(["relu", "gelu"]), decoder_mode=rnd.choice(["vanilla", "recursive", "smart"]),
            with_video=with_video, with_scene=with_scene, with_gaze=with_gaze, dense_prediction=with_video and rnd.random() < 0.6,
            image_embedding_size=32, encoder_hidden_size=32, encoder_heads=rnd.choice([4, 8]), encoder_layers=rnd.choice([1, 2]),
            encoder_d_ff=64, cross_modal_decoder_heads=rnd.choice([4, 8]), cross_modal_decoder_layers=rnd.choice([1, 2]),
            rotate_motion=rnd.random() < 0.5, normalize_motion=rnd.random() < 0.5, motion_mean=1.8, motion_std=0.9)
        spec = SPEC if with_video else None
        model = R.build_reference_model(cfg, spec).eval()
        sd = O.fill_state_dict(O.state_dict_template(cfg, spec), 20 + i)
        model.load_state_dict(sd)
        batch = O.synthetic_batch(rnd.choice([1, 2, 3]), cfg, "tiny", seed=i)
        torch.manual_seed(12345)
        try:
            with torch.no_grad():
                ref = model(batch)
        except RuntimeError as err:
            assert cfg.decoder_mode == "recursive" and cfg.dense_prediction, (i, err)
            torch.manual_seed(12345)
            with pytest.raises(RuntimeError, match="must match the size of tensor b"):
                O.Routeformer(sd, cfg, spec).forward(batch)
            raised += 1
            continue
        torch.manual_seed(12345)
        out = O.Routeformer(sd, cfg, spec).forward(batch)
        ref = ref if isinstance(ref, tuple) else (ref, None)
        out = out if isinstance(out, tuple) else (out, None)
        assert (out[0] - ref[0]).abs().max() <= 1e-6 * ref[0].abs().max(), (i, cfg)
        if ref[1] is not None:
            assert (out[1] - ref[1]).abs().max() <= 2e-5 * ref[1].abs().max(), (i, cfg)
        ran += 1
    assert ran >= 12 and raised >= 1, (ran, raised)


def test_product_module_tree_mirrors_the_reference():
    """the product `Routeformer` against the imported reference model object itself -- `state_dict()` keys in the same
    ORDER with the same shapes / dtypes, `named_parameters()` in the same order (optimizer parameter groups are built from it,
    full_comparison.py:683-691), the reference's tensors load with strict=True, and the sub-module attribute names of the checkpoint
    contract exist (routeformer.py:57-117)."""
    from tests.helpers import build_product

    gps_only = O.OracleConfig()  # the paper's GPS backbone, no video
    for cfg, spec in ((SMALL, SPEC), (gps_only, None)):
        ref = R.build_reference_model(cfg, spec)
        mine = build_product(cfg, spec)
        rsd, msd = ref.state_dict(), mine.state_dict()
        assert list(rsd) == list(msd)
        assert all(rsd[k].shape == msd[k].shape and rsd[k].dtype == msd[k].dtype for k in rsd)
        assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
        assert [n for n, _ in ref.named_buffers()] == [n for n, _ in mine.named_buffers()]
        mine.load_state_dict(rsd, strict=True)
        assert all(torch.equal(v, rsd[k]) for k, v in mine.state_dict().items())
    full_ref, full_mine = R.build_reference_model(SMALL, SPEC), build_product(SMALL, SPEC)
    for name in ("video_backbone", "frame_encoder", "left_video_embedding", "right_video_embedding", "gaze_video_embedding",
                 "video_output_embedding", "video_encoder", "gaze_encoder", "gaze_video_decoder", "gps_backbone"):
        assert hasattr(full_ref, name) and hasattr(full_mine, name), name
