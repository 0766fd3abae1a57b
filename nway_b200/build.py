"""Build libnwayb200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libnwayb200.so')
SOURCES = ['nwb_api.cu']
HEADERS = ['nwb_device.cuh', 'nwb_grid.cuh', 'nwb_grid_host.h', 'nwb_rows.cuh', 'nwb_kernels.cuh', os.path.join('..', '..', 'include', 'nwayb200.h')]

NVCC_FLAGS = [
	'-gencode', 'arch=compute_100a,code=sm_100a',
	'-O3', '-lineinfo', '-std=c++17',
	'--fmad=false',            # the reference is unfused numpy: a*b+c must round twice
	'-Xcompiler', '-fPIC', '-shared',
	'-Xptxas', '-v',
]


def needs_build():
	if not os.path.exists(LIB):
		return True
	t = os.path.getmtime(LIB)
	deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
	return any(os.path.getmtime(d) > t for d in deps)


def build_variant(out, defines):
	"""experiment builds: same sources with -D overrides into another .so (select it with $NWB_LIB)"""
	nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
	cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', out, '-lcudart']
	res = subprocess.run(cmd, capture_output=True, text=True)
	if res.returncode != 0:
		sys.stderr.write(res.stdout + res.stderr)
		raise RuntimeError('nvcc failed')
	return res.stdout + res.stderr


def build(force=False, verbose=False):
	if not force and not needs_build():
		return LIB
	nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
	cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB, '-lcudart']
	res = subprocess.run(cmd, capture_output=True, text=True)
	if res.returncode != 0 or (verbose and os.environ.get("NWB_BUILD_VERBOSE")):
		sys.stderr.write(res.stdout + res.stderr)
	if res.returncode != 0:
		raise RuntimeError('nvcc failed building libnwayb200.so')
	with open(os.path.join(HERE, 'build.log'), 'w') as f:
		f.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
	return LIB


if __name__ == '__main__':
	print(build(force=True, verbose=True))
