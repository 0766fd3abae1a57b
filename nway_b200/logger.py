"""The logger objects nway_match() accepts: anything with log(*msg), warn(msg) and progress(...) -- the duck type of the
reference's nwaylib/logger.py:28-53.  Progress bars are not drawn: the device path has no loop worth one."""
import sys
import warnings


class _Identity(object):
	"""stands in for a progress bar: wraps an iterable without touching it, ignores every bar method"""

	def __init__(self, *unused, **unused_kw):
		pass

	def __call__(self, iterable):
		return iterable

	def __getattr__(self, name):
		if name == 'start':
			return lambda *a, **k: self
		if name in ('increment', 'finish', 'update'):
			return lambda *a, **k: None
		raise AttributeError(name)


class _Logger(object):
	stream = None   # None: silent

	def log(self, *msg):
		if self.stream is not None:
			self.stream.write(' '.join(str(m) for m in msg) + '\n')

	def warn(self, msg):
		warnings.warn(msg, stacklevel=3)

	def progress(self, *args, **kwargs):
		return _Identity()


class NullOutputLogger(_Logger):
	"""warnings only"""


class NormalLogger(_Logger):
	"""messages to stderr, like the reference's default logger"""

	@property
	def stream(self):
		return sys.stderr
