"""Bisects which stream overlap breaks the capture of ParallelTrainerSteps.training_step (debug helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import routeformer_b200 as R  # noqa: E402
from routeformer_b200.parallel import DataParallelTrainer  # noqa: E402
from tests.test_gpu_model import _steps_case  # noqa: E402

for branches in (True, False):
    for wgrad in (True, False):
        gold, cfg, spec, sd, batch, model, dev_batch = _steps_case()
        model.train()
        R.Routeformer.overlap_branches = branches
        steps = R.ParallelTrainerSteps(model)
        step_fn, draw_fn = steps.graph_hooks(current_epoch=10)
        trainer = DataParallelTrainer(model, None, use_cuda_graph=True, step_fn=step_fn, draw_fn=draw_fn, overlap_wgrad=wgrad)
        try:
            loss = trainer._replay(dev_batch, ())
            torch.cuda.synchronize()
            print(f"branches={branches} wgrad={wgrad}: OK loss {loss.item():.5f}")
        except Exception as e:  # noqa: BLE001
            print(f"branches={branches} wgrad={wgrad}: FAILED {str(e).splitlines()[0][:120]}")
            torch.cuda.synchronize()
R.Routeformer.overlap_branches = True
