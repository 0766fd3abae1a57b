import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
	"""gpu tests are skipped (not failed) when no device is visible, so that a bare `pytest tests/` works
	in the CPU container; the driver selects them with -m gpu on the GPU box."""
	try:
		import torch
		have = torch.cuda.is_available()
	except Exception:
		have = False
	if have:
		return
	skip = pytest.mark.skip(reason='no CUDA device visible')
	for item in items:
		if 'gpu' in item.keywords:
			item.add_marker(skip)
