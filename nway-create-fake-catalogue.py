#!/usr/bin/env python
"""nway-create-fake-catalogue.py -- fake catalogue for the false-association calibration, with the reference's arguments.
See nway_b200/calibrate_cli.py (fake_main) and nway_b200/calibrate.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nway_b200.calibrate_cli import fake_main as main  # noqa: E402

if __name__ == '__main__':
	sys.exit(main())
