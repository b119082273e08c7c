"""Import-path shim: `from estimator.models.patchfusion import PatchFusion` resolves to the B200 implementation
(reference `estimator/models/patchfusion.py:55`).  Only the hot-path model is provided; the reference's trainers,
datasets and tools are out of scope."""
