"""bench.py contract pieces that run without a GPU: the reference arm (C port of
the reference path on the host cores) prints ONE JSON line with the agreed keys;
under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
        'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
        'cpu_baseline', 'e2e'}


def _run(env_extra):
  env = dict(os.environ, **env_extra)
  return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                         '--steps', '3', '--warmup', '1', '--cpu-cells', '8'],
                        capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_json_line():
  p = _run({})
  assert p.returncode == 0, p.stderr[-2000:]
  lines = [l for l in p.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert KEYS <= set(d)
  assert d['impl'] == 'reference' and d['metric'] == 'atom-timesteps/s' and d['value'] > 0
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
  assert d['e2e']['value'] == d['value'] and 'workload' in d['config']


def test_reference_arm_other_ranks_stay_silent():
  p = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
  assert p.returncode == 0 and p.stdout.strip() == ''
