"""Bisects what breaks the capture of ParallelTrainerSteps.training_step (debug helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import routeformer_b200 as R  # noqa: E402
from routeformer_b200.parallel import DataParallelTrainer  # noqa: E402
from tests.test_gpu_model import _steps_case  # noqa: E402


def variant(kind):
    gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
    model.train()
    steps = R.ParallelTrainerSteps(model)
    lossf = steps.trajectory_loss

    def step_fn(b, _t):
        inp, target = b["train"], b["target"]
        wp, dense = model(inp)
        tv = None
        if kind >= 1:
            with torch.no_grad():
                _, tv = model.preprocess_batch(target, training=False)
        loss = lossf(wp, target["gps"].float())
        if kind >= 2:
            loss = loss + steps.dense_loss(dense, tv[:, :dense.shape[1]].detach())
        if kind >= 3:
            steps.last_metrics = {"train_ade": R.ade(wp, target["gps"].float())}
        return loss

    def draw_fn(b, refill):
        model.prepare_draws(b["train"], training=True, refill_only=refill)
        pat = model.last_pattern
        if kind >= 1:
            model.prepare_draws(b["target"], training=False, refill_only=refill, backbone=False)
        return pat

    trainer = DataParallelTrainer(model, None, use_cuda_graph=True, step_fn=step_fn, draw_fn=draw_fn)
    try:
        loss = trainer._replay(dev_batch, ())
        torch.cuda.synchronize()
        print(f"variant {kind}: OK loss {loss.item():.5f}")
    except Exception as e:  # noqa: BLE001
        print(f"variant {kind}: FAILED {str(e).splitlines()[0][:120]}")
        torch.cuda.synchronize()


def real(epoch, with_metrics=True):
    gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
    model.train()
    steps = R.ParallelTrainerSteps(model)
    if not with_metrics:
        import routeformer_b200.experiment as E
        E.ade = lambda a, b: a.sum() * 0
        E.fde = lambda a, b: a.sum() * 0
    step_fn, draw_fn = steps.graph_hooks(current_epoch=epoch)
    trainer = DataParallelTrainer(model, None, use_cuda_graph=True, step_fn=step_fn, draw_fn=draw_fn)
    try:
        loss = trainer._replay(dev_batch, ())
        torch.cuda.synchronize()
        print(f"real epoch {epoch} metrics {with_metrics}: OK loss {loss.item():.5f}")
    except Exception as e:  # noqa: BLE001
        print(f"real epoch {epoch} metrics {with_metrics}: FAILED {str(e).splitlines()[0][:120]}")
        torch.cuda.synchronize()


real(0)
real(10)
