#!/usr/bin/env python
"""nway.py -- command-line entry point with the reference's arguments (reference: /root/reference/nway.py),
running the match path on the GPU.  See nway_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from nway_b200.cli import main  # noqa: E402

if __name__ == '__main__':
	sys.exit(main())
