"""CPU: the lines bench.py prints keep the driver's contract.  The reference arm (the oracle port of the reference's
CPU algorithm) runs here on a tiny sample; the CUDA arm must refuse to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
	e = dict(os.environ)
	e.update(env or {})
	return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
	res = run('--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0', '--ref-scale', '0.002')
	assert res.returncode == 0, res.stderr[-2000:]
	line = json.loads(res.stdout.strip().splitlines()[-1])
	assert line['impl'] == 'reference' and line['metric'] == 'candidate associations/sec' and line['unit'] == 'associations/s'
	assert line['higher_is_better'] is True and line['value'] > 0 and line['n_gpus'] == 1
	# the unmodified reference where it is available (/root/reference mounted, or its pip-installed copy oracle/_ref), else the port
	from oracle import refrun
	assert line['cpu_baseline']['kind'] == ('reference' if refrun.package_root() else 'port')
	assert line['cpu_baseline']['cores'] == 1 and line['cpu_baseline']['value'] == line['value']
	assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
	assert 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
	res = run('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--ref-scale', '0.002', env={'RANK': '1', 'WORLD_SIZE': '2'})
	assert res.returncode == 0 and res.stdout.strip() == ''


def test_cuda_arm_refuses_without_a_device():
	import torch
	if torch.cuda.is_available():
		import pytest
		pytest.skip('a device is visible here')
	res = run('--gpus', '1', '--steps', '1', '--warmup', '0', '--no-cpu')
	assert res.returncode != 0 and 'no CPU fallback' in (res.stderr + res.stdout)
