"""Build libnwayb200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libnwayb200.so')
SOURCES = ['nwb_api.cu']   # one translation unit; it includes every header of csrc/ and include/nwayb200.h

NVCC_FLAGS = [
	'-gencode', 'arch=compute_100a,code=sm_100a',
	'-O3', '-lineinfo', '-std=c++17',
	'--fmad=false',            # the reference is unfused numpy: a*b+c must round twice
	'-Xcompiler', '-fPIC', '-shared',
	'-Xptxas', '-v',
]


STAMP = LIB + '.srchash'   # what the library was built from (travels with it; file times do not always survive a copy)


def dependencies():
	"""every file the library is compiled from: all of csrc/ and the public header"""
	deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh', '.h'))]
	return deps + [os.path.join(HERE, '..', 'include', 'nwayb200.h')]


def source_hash():
	"""sha256 over the compiler flags and the contents of dependencies()"""
	import hashlib
	h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
	for d in dependencies():
		h.update(os.path.basename(d).encode())
		with open(d, 'rb') as f:
			h.update(f.read())
	return h.hexdigest()


def needs_build():
	"""stale = built from other sources or flags.  Decided by content where the stamp written by build() is there, by
	file times otherwise (a library built by hand)."""
	if not os.path.exists(LIB):
		return True
	if os.path.exists(STAMP):
		with open(STAMP) as f:
			return f.read().strip() != source_hash()
	t = os.path.getmtime(LIB)
	return any(os.path.getmtime(d) > t for d in dependencies() + [os.path.abspath(__file__)])


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
	stamp = source_hash()
	if os.path.exists(STAMP):
		os.remove(STAMP)
	res = subprocess.run(cmd, capture_output=True, text=True)
	if res.returncode != 0 or (verbose and os.environ.get("NWB_BUILD_VERBOSE")):
		sys.stderr.write(res.stdout + res.stderr)
	if res.returncode != 0:
		raise RuntimeError('nvcc failed building libnwayb200.so')
	with open(os.path.join(HERE, 'build.log'), 'w') as f:
		f.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
	with open(STAMP, 'w') as f:
		f.write(stamp + '\n')
	return LIB


if __name__ == '__main__':
	print(build(force=True, verbose=True))
