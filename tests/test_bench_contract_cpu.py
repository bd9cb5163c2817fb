"""bench.py's JSON contract, checked on the CPU (reference arm) with the smallest configuration."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, universal_newlines=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _run('--impl', 'reference', '--config', 'cfg1', '--steps', '2', '--warmup', '1')
    assert j['impl'] == 'reference' and j['metric'] == 'RRI sweeps/sec' and j['unit'] == 'sweeps/s'
    assert j['higher_is_better'] is True and j['n_gpus'] == 1 and j['steps'] == 2 and j['warmup'] == 1
    assert j['value'] > 0 and abs(j['ms_per_step'] * j['value'] - 1e3) < 1e-6 * 1e3
    assert j['vs_baseline'] is None and j['data'] == 'synthetic' and 'workload' in j['config']
    cb = j['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['value'] == j['value'] and 'sample' in cb
    # all host cores, whatever OMP_NUM_THREADS says (torch.distributed.run exports OMP_NUM_THREADS=1)
    assert cb['cores'] == len(os.sched_getaffinity(0)) == cb['host_cores']
    assert j['e2e'] == {'value': j['value'], 'unit': 'sweeps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_full_threads_under_torchrun_env():
    """rank 0 of a torchrun launch: OMP_NUM_THREADS=1 in the environment must not reduce the CPU arm to one thread;
    RRI_BENCH_PORT=1 selects the NumPy port even when the reference copy is present"""
    env = dict(os.environ, RANK='0', WORLD_SIZE='2', LOCAL_RANK='0', OMP_NUM_THREADS='1', RRI_BENCH_PORT='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--config', 'cfg1', '--steps', '1', '--warmup', '0'], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, universal_newlines=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][0])
    assert j['cpu_baseline']['cores'] == len(os.sched_getaffinity(0)) and j['cpu_baseline']['kind'] == 'port'


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--config', 'cfg1', '--steps', '1', '--warmup', '0'], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, universal_newlines=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''
