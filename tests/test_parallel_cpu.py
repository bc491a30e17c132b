"""CPU (gloo, world_size 2): host-side logic of the data-parallel path -- batch sharding, bucketed flat all-reduce,
and the arena bookkeeping the trainer relies on.  (The CUDA optimiser / model step is covered by the -m gpu tests.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from routeformer_b200.parallel import allreduce_flat, bucket_bounds, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        w = torch.randn(5000, 3)
        data = {"x": torch.randn(8, 5000), "y": torch.randn(8, 3)}
        local = shard_batch(data, rank, world)
        # stand-in for the model: per-rank gradient of a sum-reduced loss over the local shard, written into a flat arena
        wl = w.clone().requires_grad_()
        (((local["x"] @ wl) - local["y"]) ** 2).sum().backward()
        flat = wl.grad.reshape(-1).clone()
        for h in allreduce_flat(flat, n_buckets=3):
            h.wait()
        if rank == 0:
            wf = w.clone().requires_grad_()
            (((data["x"] @ wf) - data["y"]) ** 2).sum().backward()
            out.put(float((flat - wf.grad.reshape(-1)).abs().max() / wf.grad.abs().max()))
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_allreduce_to_full_batch_gradient():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) < 1e-5


def test_bucket_bounds_cover_in_reverse_order():
    for n, k in [(10, 4), (100000, 4), (4097, 3), (1, 2)]:
        b = bucket_bounds(n, k, align=16)
        assert b[0][1] == n and b[-1][0] == 0
        assert sorted(b) == b[::-1]
        assert all(lo < hi for lo, hi in b)
        cover = sorted(b)
        assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
    assert bucket_bounds(0, 4) == []


def test_shard_batch():
    batch = {"gps": torch.arange(16).view(8, 2), "left_video": torch.zeros(8, 3, 2)}
    s1 = shard_batch(batch, 1, 4)
    assert s1["gps"].tolist() == [[4, 5], [6, 7]] and s1["left_video"].shape == (2, 3, 2)
    try:
        shard_batch(batch, 0, 3)
        assert False
    except ValueError:
        pass


def test_arena_round_trip_and_grad_reattach():
    """Arena keeps state_dict round-trips, fuses q/k/v storage and re-attaches gradient views after zero_grad(set_to_none)."""
    from oracle import routeformer_oracle as O
    from routeformer_b200.arena import Arena
    from routeformer_b200.functional import _fused
    from tests.helpers import build_product, case_from_golden, load_golden

    gold = load_golden("full_small_eval")
    cfg, spec, sd, _ = case_from_golden(gold)
    model = build_product(cfg, spec)
    model.load_state_dict(sd)
    arena = Arena.ensure(model)
    assert arena.valid() and arena.n_trainable <= arena.size
    assert all(torch.equal(v, sd[k]) for k, v in model.state_dict().items())
    att = model.video_encoder.encoder.attn_layers[0].attention
    w = _fused(att.query_projection.weight, att.key_projection.weight, att.value_projection.weight)
    assert w is not None and w.shape == (3 * 128, 128) and torch.equal(w[128:256], att.key_projection.weight)
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert frozen == ["video_backbone.proj.weight", "video_backbone.proj.bias"]
    model.zero_grad(set_to_none=True)
    assert Arena.ensure(model) is arena
    g = att.out_projection.weight.grad
    assert g is not None and g.data_ptr() == arena.grad.data_ptr() + 4 * arena.offsets[id(att.out_projection.weight)]
    model.load_state_dict(O.fill_state_dict(sd, 99))  # in-place copy keeps the arena
    assert arena.valid()


def test_deferred_draws_fill_the_same_pinned_tables():
    """prepare_draws(refill_only="defer") + commit_deferred_draws() == prepare_draws(refill_only=True): same CPU RNG order, same
    staging-buffer contents, and nothing is written before the commit (the previous graph replay may still read the buffer)."""
    import torch

    from tests.helpers import build_product, case_from_golden, load_golden

    cfg, spec, sd, batch = case_from_golden(load_golden("full_small_train"))
    model = build_product(cfg, spec)
    torch.manual_seed(5)
    model.prepare_draws(batch, training=True, refill_only=True)
    ref = {k: v[0].clone() for k, v in model._idx_slots.items()}
    for v in model._idx_slots.values():
        v[0].zero_()
    torch.manual_seed(5)
    model.prepare_draws(batch, training=True, refill_only="defer")
    assert all(bool((v[0] == 0).all()) for v in model._idx_slots.values())
    assert model.commit_deferred_draws() and not model.commit_deferred_draws()
    assert ref and all(torch.equal(ref[k], v[0]) for k, v in model._idx_slots.items())


def test_sample_draw_plan_equals_sequential_plans():
    """prepare_sample_draws(batch, n) (eval_step's n stochastic forwards as one pass): the same CPU draws in the same order as n
    sequential eval forwards, and table s of every stacked [n, L_Q, U] entry is exactly forward s's table."""
    import torch

    from tests.helpers import build_product, case_from_golden, load_golden

    for case in ("full_small_eval", "dreyeve_small", "gps_only_paper"):
        cfg, spec, sd, batch = case_from_golden(load_golden(case))
        model = build_product(cfg, spec).eval()
        n = 3
        torch.manual_seed(7)
        log, seq = [], []
        for _ in range(n):
            plan = model.prepare_draws(batch, False)
            log += model.last_draw_log
            seq.append(([t.clone() for t in plan["visual"]["tables"]] if plan["visual"] else [], [t.clone() for t in plan["backbone"][1]]))
        torch.manual_seed(7)
        plan = model.prepare_sample_draws(batch, n)
        assert model.last_draw_log == log
        stacked = (plan["visual"]["tables"] if plan["visual"] else [], plan["backbone"][1])
        for part in (0, 1):
            assert len(stacked[part]) == len(seq[0][part])
            for i, t in enumerate(stacked[part]):
                assert t.shape[0] == n
                for s_ in range(n):
                    assert torch.equal(t[s_], seq[s_][part][i][0]), (case, part, i, s_)


def test_early_allreduce_ranges_partition_the_gradient_arena():
    """The overlapped all-reduce sends the GPS backbone's slices of the flat gradient arena early and the rest after backward:
    the two range lists must partition [0, n_trainable) and the early one must hold exactly the backbone's parameters."""
    from oracle import routeformer_oracle as O
    from routeformer_b200.parallel import DataParallelTrainer
    from tests.helpers import build_product

    cfg = O.OracleConfig(d_model=64, n_heads=4, e_layers=3, d_ff=128, with_video=True, with_gaze=True, dense_prediction=True,
                         encoder_layers=2, encoder_d_ff=64, image_embedding_size=32, encoder_hidden_size=32)
    model = build_product(cfg, O.BackboneSpec(image_size=32, patch=8, channels=48))
    trainer = DataParallelTrainer(model, None)
    early, late = trainer._split_ranges("gps_backbone.")
    arena = trainer.arena
    covered = sorted(early + late)
    assert covered[0][0] == 0 and covered[-1][1] == arena.n_trainable
    assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))
    assert len(early) <= 3  # the q/k/v groups and the remaining parameters of the backbone: a few large contiguous slices
    inside = lambda o: any(lo <= o < hi for lo, hi in early)
    for name, p in model.named_parameters():
        if p.requires_grad:
            assert inside(arena.offsets[id(p)]) == name.startswith("gps_backbone."), name
    n_early = sum(hi - lo for lo, hi in early)
    assert n_early >= sum(p.numel() for n, p in model.named_parameters() if n.startswith("gps_backbone.") and p.requires_grad)


def test_host_draw_plan_equals_the_reference_random_stream():
    """The product's host-side plan of one eval forward (every `torch.randint` of the ProbSparse attentions, made up front and
    uploaded as index tables) against the oracle executing the reference algorithm with the same CPU seed: the same calls in the
    same order (the golden `draws` list) and the SAME index tables, entry for entry."""
    import torch

    from oracle import routeformer_oracle as O
    from tests.helpers import build_product, case_from_golden, load_golden

    for case in ("full_small_eval", "dreyeve_small", "no_gaze_small", "no_scene_small", "gps_only_paper"):
        gold = load_golden(case)
        cfg, spec, sd, batch = case_from_golden(gold)
        torch.manual_seed(12345)
        draw = O.CpuRandint()
        with torch.no_grad():
            O.Routeformer(sd, cfg, spec).forward(batch, training=False, draw=draw)
        model = build_product(cfg, spec).eval()
        torch.manual_seed(12345)
        plan = model.prepare_draws(batch, False)
        assert [(lk, (lq, u)) for lk, lq, u in model.last_draw_log] == [tuple(d) for d in gold["draws"]], case
        tables = (list(plan["visual"]["tables"]) if plan["visual"] else []) + list(plan["backbone"][1])
        assert len(tables) == len(draw.drawn), case
        for i, (t, ref) in enumerate(zip(tables, draw.drawn)):
            assert t.dtype == torch.int32 and tuple(t.shape) == (1,) + tuple(ref.shape), (case, i)
            assert torch.equal(t[0].long(), ref), (case, i)


def test_state_dict_layout_equals_the_reference_layout():
    """`state_dict()` of the product speaks the reference's layout -- the same keys, shapes and dtypes (persistent
    buffers included) as recorded from the reference model in every golden file, from the small cases to the paper configuration
    (639 entries), and `load_state_dict(strict=True)` of the reference's tensors round-trips bit for bit."""
    import torch

    from tests.helpers import build_product, case_from_golden, load_golden

    for case in ("gps_only_paper", "full_small_eval", "dreyeve_small", "no_gaze_small", "no_scene_small", "sparse_small",
                 "normalized_small", "autoregressive_small", "full_paper_eval"):
        gold = load_golden(case)
        cfg, spec, sd, _ = case_from_golden(gold)
        model = build_product(cfg, spec)
        mine = model.state_dict()
        assert {(k, tuple(v.shape), str(v.dtype)) for k, v in mine.items()} == set(map(tuple, gold["layout"])), case
        missing, unexpected = model.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        assert all(torch.equal(v, sd[k]) for k, v in model.state_dict().items()), case
    assert len(load_golden("full_paper_eval")["layout"]) == 639


def _worker_real_model(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import routeformer_oracle as O
        from routeformer_b200.parallel import DataParallelTrainer
        from tests.helpers import build_product

        cfg = O.OracleConfig(d_model=64, n_heads=4, e_layers=2, d_ff=128, with_video=True, with_gaze=True, dense_prediction=True,
                             encoder_layers=1, encoder_d_ff=64, image_embedding_size=32, encoder_hidden_size=32)
        torch.manual_seed(100 + rank)  # different initial weights per rank: broadcast_parameters must make them rank 0's
        model = build_product(cfg, O.BackboneSpec(image_size=32, patch=8, channels=48))
        trainer = DataParallelTrainer(model, None)
        assert trainer.world == world
        arena = trainer.arena
        before = arena.param.clone()
        trainer.broadcast_parameters()
        gathered = [torch.empty_like(arena.param) for _ in range(world)]
        dist.all_gather(gathered, arena.param)
        same_params = all(torch.equal(g, gathered[0]) for g in gathered) and (rank == 0) == torch.equal(before, arena.param)
        # the step's two all-reduce phases (GPS backbone ranges early, the rest late) over the real gradient arena == one flat
        # all-reduce == the sum of the ranks' gradients
        g = torch.Generator().manual_seed(7 + rank)
        arena.grad.copy_(torch.randn(arena.grad.shape, generator=g))
        mine = arena.grad.clone()
        early, late = trainer._split_ranges("gps_backbone.")
        for lo, hi in early + late:
            dist.all_reduce(arena.grad[lo:hi], op=dist.ReduceOp.SUM)
        flat = mine.clone()
        for h in allreduce_flat(flat, n_buckets=4):
            h.wait()
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        total = sum(both)
        n = arena.n_trainable
        ok = torch.equal(arena.grad[:n], flat[:n]) and torch.allclose(flat[:n], total[:n], rtol=0, atol=1e-6)
        # the views the autograd blocks write through are views of that same arena
        p = model.gps_backbone.decoder.projection.weight
        views_ok = p.grad is not None and p.grad.data_ptr() == arena.grad.data_ptr() + 4 * arena.offsets[id(p)]
        out.put((rank, bool(same_params), bool(ok), bool(views_ok)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_real_model_broadcast_and_two_phase_allreduce():
    """World size 2 over gloo with the REAL model's flat parameter / gradient arenas: `broadcast_parameters` makes every rank hold
    rank 0's weights, and the trainer's early (GPS backbone) + late all-reduce ranges together reproduce one flat sum-all-reduce of
    the whole gradient arena, i.e. the sum of the ranks' gradients."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_real_model, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    results = sorted(out.get(timeout=10) for _ in range(2))
    assert results == [(0, True, True, True), (1, True, True, True)], results
