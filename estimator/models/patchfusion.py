"""Drop-in for the reference module of the same path: `PatchFusion` registered under the same name."""
from estimator.registry import MODELS
from patchfusion_b200.model import PatchFusion as _PatchFusion

PatchFusion = MODELS.register_module(name='PatchFusion')(_PatchFusion)
