"""Drop-in switch for a checkout of the reference: `routeformer_b200.compat.patch_reference()`.

The experiment driver imports the path's symbols from the reference package (experiments/full_comparison.py:25-46):
    from routeformer import Routeformer
    from routeformer.models import RouteformerConfig
    from routeformer.models.gps_backbone import GPSBackboneConfig, Informer
    from routeformer.losses.future_discounted_mse import FutureDiscountedLoss
    from routeformer.score import ade, fde
Everything else it imports (datasets, ablation backbones, schedulers, timm wrappers) is outside the accelerated path and stays
the reference's own code.  `patch_reference()` rebinds exactly those names, in the already-importable reference package, to the
CUDA implementations -- call it once before the experiment module is imported and no line of the experiment has to change:

    import routeformer_b200.compat as compat; compat.patch_reference()
    import full_comparison            # builds routeformer_b200 models, same constructor calls, same state_dict

Alternatively edit the five import lines.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List

#  module of the reference           name            -> attribute of routeformer_b200
_BINDINGS = {
    "routeformer": ["Routeformer"],
    "routeformer.models": ["Routeformer", "RouteformerConfig"],
    "routeformer.models.routeformer": ["Routeformer"],
    "routeformer.models.config": ["RouteformerConfig"],
    "routeformer.models.gps_backbone": ["GPSBackboneConfig", "Informer"],
    "routeformer.models.gps_backbone.config": ["GPSBackboneConfig"],
    "routeformer.models.video_backbone": ["VideoBackboneConfig", "VideoBackboneModule"],
    "routeformer.models.video_backbone.config": ["VideoBackboneConfig", "VideoBackboneModule"],
    "routeformer.models.cross_modal_transformer": ["PerceiveEncoder", "PerceiveDecoder"],
    "routeformer.losses.future_discounted_mse": ["FutureDiscountedLoss"],
    "routeformer.score": ["ade", "fde"],
    "routeformer.score.error": ["ade", "fde"],
}


def patch_reference(strict: bool = False) -> Dict[str, List[str]]:
    """Rebinds the hot path's public names inside the reference package.  Returns {module: [names rebound]}.
    Modules of the reference that cannot be imported in this environment are skipped unless `strict`."""
    import routeformer_b200 as R

    done: Dict[str, List[str]] = {}
    for mod_name, names in _BINDINGS.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                if strict:
                    raise
                continue
        for n in names:
            setattr(mod, n, getattr(R, n))
            done.setdefault(mod_name, []).append(n)
    return done
