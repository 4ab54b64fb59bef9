"""The CPU arm of bench.py (`--impl reference`: the oracle port of the reference path on the host cores) runs without a
GPU; the driver launches it next to the GPU arm and divides the two lines, so the keys both lines share are checked
here on the real command (one step of the default workload, ~15 s)."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'], cwd=ROOT,
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['n_gpus'] == 1 and line['steps'] == 1
    assert line['metric'] == 'aggregate detection FPS on synthetic 640x480 streams' and line['unit'] == 'frames/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['ms_per_step'] > 0
    assert 'configs[2]' in line['config']['workload']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['unit'] == line['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
    base = line['cpu_baseline']
    assert base['kind'] == 'port' and base['cores'] >= 1 and base['value'] == line['value'] and base['sample']
    assert not line.get('gpu_launches')          # nothing of ours runs on that arm
