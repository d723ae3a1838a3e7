"""bench.py contract checks that run without a GPU: the reference arm (the reference's modules / the CPU oracle on the host cores) prints the
driver's JSON line, rank > 0 stays silent under a multi-process launch, and the product arm refuses to run
without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ['--steps', '1', '--warmup', '0', '--n-antigen', '9', '--num-t', '4']


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})}, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(['--impl', 'reference'] + SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'designed_cdr_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['data'] == 'synthetic' and d['dtype'] == 'f32'
    assert d['value'] > 0 and abs(d['ms_per_step'] * d['value'] - 1e3) < 1e-6 * 1e3
    assert d['config']['n_res'] == 239 and d['config']['num_t'] == 4 and 'workload' in d['config']
    cb = d['cpu_baseline']
    # the reference's own modules when oracle/_ref was materialised (oracle/build_ref.py), else the oracle port
    ref_built = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', '.complete'))
    assert cb['kind'] == ('reference' if ref_built else 'port') and cb['cores'] >= 1 and cb['value'] == d['value']
    assert 'ScoreNetwork' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_exit_silently():
    r = _run(['--impl', 'reference', '--gpus', '2'] + SMALL, env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(SMALL)
    assert r.returncode != 0 and not any(l.startswith('{') for l in r.stdout.splitlines())
