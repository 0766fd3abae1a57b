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


def test_committed_gpu_bench_line_is_consistent():
	"""the bench line measured on B200 for the committed tree (profiles/r02_last_bench.json): every key of the driver's
	contract, and the derived numbers follow from the measured ones (value from ms_per_step, roofline.achieved from the
	algorithmic bytes and the kernel's time, frac from the measured peak); roofline.traffic is there because the ncu capture
	it comes from was taken from exactly the kernel sources in this tree"""
	import bench
	line = json.load(open(os.path.join(ROOT, 'profiles', 'r02_last_bench.json')))
	for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
			'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'gpu_launches', 'clocks'):
		assert k in line, k
	base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
	assert base['metric'].startswith(line['metric']) and line['dtype'] == 'f64' and line['warmup'] >= 3 and line['vs_baseline'] is None
	rows = line['config']['rows_per_gpu'] * line['n_gpus']
	assert abs(line['value'] - rows / (line['ms_per_step'] * 1e-3)) <= 1e-6 * line['value']
	r = line['roofline']
	assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and r['peak_kind'] == 'measured'
	assert r['algorithmic_bytes'] == 16 * 10**7 + 16 * line['config']['pairs_per_gpu']   # k_pairs
	assert abs(r['achieved'] - r['algorithmic_bytes'] / (r['kernel_ms'] * 1e-3) / 1e9) <= 1e-6 * r['achieved']
	assert abs(r['frac'] - r['achieved'] / r['peak']) <= 1e-9
	assert r['algorithmic_bytes'] <= r['traffic'] <= 1.2 * r['algorithmic_bytes']   # no wasted re-reads
	traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
	assert traffic['kernel_source_sha256'] == bench.kernel_source_hash()
	assert r['traffic'] == traffic['kernels']['k_pairs']['dram_bytes_per_launch']
	e = line['e2e']
	assert e['h2d_bytes_per_step'] == 8 * 3 * (10**5 + 10**7) and e['d2h_bytes_per_step'] == 96 * line['config']['rows_per_gpu']
	assert e['value'] < line['value'] and abs(e['value'] - rows / (e['ms_per_step'] * 1e-3)) <= 1e-6 * e['value']
	assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] == 1
	assert line['gpu_launches'] > 0 and not set(line['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
