"""Test infrastructure: imports the read-only reference tree (/root/reference) in THIS container only.

Used by oracle/make_golden.py to (i) pin oracle/pf_oracle.py against the real reference and (ii) emit the
small fixtures committed under tests/golden/.  Nothing here runs on the GPU box (the reference is absent there)
and nothing in patchfusion_b200/ imports it.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF, 'estimator'))


def _enter():
    for p in (os.path.join(REF, 'external'), REF, os.path.join(HERE, 'shims')):
        if p not in sys.path:
            sys.path.insert(0, p)


def ref_config(encoder='vits'):
    path = os.path.join(REF, 'configs/patchfusion_depthanything/depthanything_%s_patchfusion_u4k.py' % encoder)
    return runpy.run_path(path)['model']['config']


def build_reference(encoder='vits', cfg=None):
    """Instantiate the reference PatchFusion (HF-dict constructor branch, PF:70-78) on CPU."""
    _enter()
    cwd = os.getcwd()
    os.chdir(REF)  # DPT:140 uses a relative torch.hub path
    try:
        import warnings
        warnings.filterwarnings('ignore')
        from estimator.models.patchfusion import PatchFusion
        cfg = dict(cfg if cfg is not None else ref_config(encoder))
        model = PatchFusion(cfg).eval()
    finally:
        os.chdir(cwd)
    return model
