"""Recipe for oracle/_ref: the UNMODIFIED reference package (nway 4.7.1, /root/reference) installed with pip into
oracle/_ref/ so that it travels to the GPU box with the repository snapshot (oracle/_ref/ is git-ignored: it never
enters the history, and no reference source is copied into the tree by hand).

    python oracle/install_ref.py

TEST / BENCH INFRASTRUCTURE ONLY.  bench.py --impl reference and bench.py's cpu_baseline leg time the real
nwaylib.nway_match from this directory (oracle/refrun.py loads it behind the same three stub modules as in the build
container: astropy, healpy and matplotlib are not installed anywhere here); the product never imports it.
/root/reference is read-only and pip builds in the source tree, so the install runs from a scratch copy."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')
SOURCE = '/root/reference'


def installed():
	return os.path.isfile(os.path.join(TARGET, 'nwaylib', '__init__.py'))


def install(force=False):
	"""returns the directory, or None when there is nothing to install from (the GPU box: only the shipped copy exists)"""
	if installed() and not force:
		return TARGET
	if not os.path.isdir(os.path.join(SOURCE, 'nwaylib')):
		return None
	tmp = tempfile.mkdtemp(prefix='nwb_refsrc_')
	try:
		src = os.path.join(tmp, 'reference')
		shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns('.git', 'doc'))
		if os.path.isdir(TARGET):
			shutil.rmtree(TARGET)
		cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--quiet',
			'--find-links', '/opt/wheelhouse', '--target', TARGET, src]
		res = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
		if res.returncode != 0:
			sys.stderr.write(res.stdout + res.stderr)
			raise RuntimeError('pip install of the reference into oracle/_ref failed')
	finally:
		shutil.rmtree(tmp, ignore_errors=True)
	# the installed files must be the reference's, byte for byte
	for name in sorted(os.listdir(os.path.join(SOURCE, 'nwaylib'))):
		a, b = os.path.join(SOURCE, 'nwaylib', name), os.path.join(TARGET, 'nwaylib', name)
		if name.endswith('.py'):
			assert open(a, 'rb').read() == open(b, 'rb').read(), name
	return TARGET


if __name__ == '__main__':
	print(install(force=True))
