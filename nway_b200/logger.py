"""Logger duck types of the reference (nwaylib/logger.py:28-53): log / warn / progress."""
import sys
import warnings


class _PassThroughBar(object):
	def __init__(self, *args, **kwargs):
		pass

	def __call__(self, it):
		return it

	def start(self):
		return self

	def increment(self):
		pass

	def finish(self):
		pass


class NullOutputLogger(object):
	def log(self, *msg):
		pass

	def warn(self, msg):
		warnings.warn(msg, stacklevel=3)

	def progress(self, *args, **kwargs):
		return _PassThroughBar()


class NormalLogger(NullOutputLogger):
	def log(self, *msg):
		sys.stderr.write('%s\n' % ' '.join(str(m) for m in msg))
