"""bench.py contract on CPU: the reference arm (oracle port on host cores) prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(900)


def test_reference_arm_json():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--encoder', 'vits',
                        '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'tiles/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'tiles/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and 'tile' in cb['sample']
    assert 'workload' in d['config'] and 'model' not in d['config']
