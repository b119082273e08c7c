"""The C-ABI library: builds for sm_100a without a GPU, exports every symbol include/pf_b200.h declares, and the
ctypes mirror of pf_gemm_desc has the header's layout.  No compute calls here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'include', 'pf_b200.h')


def _declared():
    src = re.sub(r'/\*.*?\*/', '', open(HDR).read(), flags=re.S)
    return sorted(set(re.findall(r'\b(pf_[A-Za-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_header_symbols():
    from patchfusion_b200 import build, lib
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.check_output(['nm', '-D', '--defined-only', path], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == declared
    h = lib.load()
    assert h.pf_version() >= 100


def test_sass_is_blackwell_native():
    from patchfusion_b200 import build
    path = build.build()
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass or 'UTCMMA' in sass or 'UTC' in sass, 'no tcgen05.mma in SASS'
    assert 'UTMALDG' in sass, 'no TMA loads in SASS'
    assert 'LDTM' in sass, 'no tcgen05.ld in SASS'


def test_gemm_desc_layout(tmp_path):
    import ctypes
    from patchfusion_b200.lib import GemmDesc
    fields = [f[0] for f in GemmDesc._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){' % HDR
    for f in fields:
        prog += 'printf("%s %%zu\\n", offsetof(pf_gemm_desc, %s));' % (f, f)
    prog += 'printf("sizeof %zu\\n", sizeof(pf_gemm_desc));return 0;}'
    c = tmp_path / 'off.c'
    c.write_text(prog)
    exe = tmp_path / 'off'
    subprocess.check_call(['gcc', str(c), '-o', str(exe)])
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        k, v = line.split()
        if k == 'sizeof':
            assert int(v) == ctypes.sizeof(GemmDesc)
        else:
            assert getattr(GemmDesc, k).offset == int(v), k


def test_no_fallback_when_library_missing(monkeypatch, tmp_path):
    from patchfusion_b200 import lib
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(lib.PFError):
        lib.load()


def test_ctypes_signatures_match_header_arity():
    """every `int pf_*(...)` prototype in include/pf_b200.h is bound in lib.SIGNATURES with the same arity"""
    from patchfusion_b200 import lib
    src = re.sub(r'/\*.*?\*/', '', open(HDR).read(), flags=re.S)
    protos = re.findall(r'\bint\s+(pf_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S)
    assert len(protos) >= 25
    from patchfusion_b200 import stage
    h = stage._bind()
    for name, args in protos:
        if name == 'pf_version':
            continue
        n = 0 if args.strip() == 'void' else len(args.split(','))
        if name in stage.STAGE_EXPORTS:          # stage-level entry points are bound in patchfusion_b200/stage.py
            assert n == len(getattr(h, name).argtypes), (name, n)
            continue
        assert name in lib.SIGNATURES, name
        assert n == len(lib.SIGNATURES[name]), (name, n, len(lib.SIGNATURES[name]))


def test_stage_struct_layouts(tmp_path):
    """every struct of the stage-level ABI (pf_layer ... pf_fusion, pf_map, pf_branch_out): sizeof and the offset of
    every field of the ctypes mirror equal the header's (checked with gcc)"""
    import ctypes
    from patchfusion_b200 import stage
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){' % HDR
    for cname, cls in stage.STRUCTS.items():
        prog += 'printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname)
        for f in cls._fields_:
            prog += 'printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0])
    prog += 'return 0;}'
    c = tmp_path / 'lay.c'
    c.write_text(prog)
    exe = tmp_path / 'lay'
    subprocess.check_call(['gcc', str(c), '-o', str(exe)])
    n = 0
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, field, v = line.split()
        cls = stage.STRUCTS[cname]
        if field == 'sizeof':
            assert ctypes.sizeof(cls) == int(v), cname
        else:
            assert getattr(cls, field).offset == int(v), (cname, field)
        n += 1
    assert n > 100


def test_tuning_options_match_header_and_validate():
    """PF_OPT_* of include/pf_b200.h == lib.OPT_*; pf_set_option is a pure host call: accepts every declared option,
    rejects unknown ones with a message (no GPU needed)"""
    from patchfusion_b200 import lib
    src = open(HDR).read()
    opts = dict(re.findall(r'#define\s+(PF_OPT_[A-Z_]+)\s+(\d+)', src))
    assert len(opts) >= 6
    for name, val in opts.items():
        assert getattr(lib, name[3:]) == int(val), name
    h = lib.load()
    for val in opts.values():
        assert h.pf_set_option(int(val), 1) == 0
    assert h.pf_set_option(99, 1) != 0 and b'unknown option' in h.pf_last_error()
    # restore the documented defaults for the other tests of this process
    for name, default in (('PF_OPT_TMA_EPILOGUE', 1), ('PF_OPT_HALO_MULTICAST', 1), ('PF_OPT_GEMM_MULTICAST', 1),
                          ('PF_OPT_FUSED_RESAMPLE', 0), ('PF_OPT_PDL', 0), ('PF_OPT_RESIZE_SEPARABLE', 0)):
        assert h.pf_set_option(int(opts[name]), default) == 0
