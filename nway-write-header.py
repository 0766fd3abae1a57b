#!/usr/bin/env python
"""nway-write-header.py -- name and sky area of a catalogue (EXTNAME, SKYAREA), set in place, with the reference's arguments.
See nway_b200/calibrate_cli.py (write_header_main)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nway_b200.calibrate_cli import write_header_main as main  # noqa: E402

if __name__ == '__main__':
	sys.exit(main())
