"""CPU: the C-ABI library loads and exports every symbol include/nwayb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
	text = open(os.path.join(ROOT, 'include', 'nwayb200.h')).read()
	text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
	return sorted(set(re.findall(r'\b(nwb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
	syms = header_symbols()
	for needed in ('nwb_create', 'nwb_destroy', 'nwb_set_catalogue', 'nwb_set_params', 'nwb_set_maghist',
			'nwb_match', 'nwb_finalize', 'nwb_truncate', 'nwb_fetch', 'nwb_last_error', 'nwb_timing', 'nwb_dist', 'nwb_log_bf'):
		assert needed in syms


def test_library_exports_every_declared_symbol():
	from nway_b200 import _lib, build
	build.build()
	lib = ctypes.CDLL(_lib.LIB_PATH)
	for name in header_symbols():
		assert hasattr(lib, name), 'libnwayb200.so does not export %s' % name
	# and the binding declares a prototype for each of them
	assert set(header_symbols()) == set(_lib.EXPORTS), set(header_symbols()) ^ set(_lib.EXPORTS)
	assert _lib.load().nwb_version() >= 100


def test_build_is_stale_by_content_not_by_file_time(monkeypatch):
	"""build() stamps the library with the hash of its sources and flags: a copy of the tree whose file times changed
	(the push to the GPU box) is not rebuilt, a changed source is"""
	from nway_b200 import build
	build.build()
	assert os.path.exists(build.STAMP) and not build.needs_build()
	assert len(build.dependencies()) >= 8 and all(os.path.exists(d) for d in build.dependencies())
	old = os.stat(build.LIB)
	os.utime(build.LIB, (old.st_atime, old.st_mtime - 10 * 365 * 86400))   # the library looks ten years older than its sources
	try:
		assert not build.needs_build()
	finally:
		os.utime(build.LIB, (old.st_atime, old.st_mtime))
	monkeypatch.setattr(build, 'NVCC_FLAGS', build.NVCC_FLAGS + ['-DSOMETHING_ELSE'])
	assert build.needs_build()


def test_no_cpu_fallback_without_device():
	"""the product path must fail loudly, not fall back, when no CUDA device is present"""
	import torch
	if torch.cuda.is_available():
		pytest.skip('a device is visible here')
	import nway_b200
	from nway_b200 import _lib
	from tests import cases
	with pytest.raises(_lib.NwbError):
		nway_b200.nway_match(cases.uniform_patch(1, (10, 100), (1.0, 0.5), 0.01), 5.0, 0.9, logger=nway_b200.NullOutputLogger())


def test_product_does_not_import_the_oracle():
	pkg = os.path.join(ROOT, 'nway_b200')
	for dirpath, _, files in os.walk(pkg):
		for f in files:
			if f.endswith(('.py', '.cu', '.cuh', '.h')):
				text = open(os.path.join(dirpath, f)).read()
				assert 'oracle' not in text.replace('no CPU fallback', ''), '%s mentions the oracle' % f
