#!/usr/bin/env python
"""nway-explain.py -- the associations of one primary source of a match table as text, with the reference's arguments.
See nway_b200/calibrate_cli.py (explain_main)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nway_b200.calibrate_cli import explain_main as main  # noqa: E402

if __name__ == '__main__':
	sys.exit(main())
