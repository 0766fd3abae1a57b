#!/usr/bin/env python
"""nway-calibrate-cutoff.py -- p_any cut-off table for the false-association calibration, with the reference's arguments.
See nway_b200/calibrate_cli.py (cutoff_main) and nway_b200/calibrate.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nway_b200.calibrate_cli import cutoff_main as main  # noqa: E402

if __name__ == '__main__':
	sys.exit(main())
