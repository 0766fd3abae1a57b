"""Minimal FITS binary-table I/O for the nway.py command-line surface (no astropy in this image).

The reference reads its catalogues with astropy (`pyfits.open(f)[1]`, nway.py:174-181) and writes the match
table with `BinTableHDU.from_columns` + `HDUList.writeto` (nway.py:629-649, fastskymatch.py:345-363).  This
module covers exactly what that needs: a primary HDU without data followed by ONE BINTABLE extension with
columns of type L, B, I, J, K, E, D, C, M (scalar or fixed-length vectors such as 2E), bit arrays nX (carried as their
bytes) and fixed-width strings nA --
fixed-width big-endian rows in 2880-byte blocks of 80-character header cards.  TSCALn / TZEROn are applied on read as
astropy does: the unsigned-integer convention (TSCAL 1, TZERO 2^15 / 2^31 / 2^63; -128 for signed bytes) gives
uint16 / uint32 / uint64 / int8 columns, which are written back with the same keywords; any other scaling gives a
float64 column (format D).  Variable-length columns (P, Q descriptors) cannot be carried through without their heap:
they are left out and listed in Table.skipped instead of making the catalogue unreadable.

	table = read_table('COSMOS_XMM.fits')        # Table: .name, .header, .columns, .formats, .data
	write_table('out.fits', columns, extname='NWAYMATCH', primary_header=[...], table_header=[...])
"""
import datetime
from collections import OrderedDict

import numpy

BLOCK = 2880

# TFORM letter -> numpy big-endian dtype
_FORMATS = {'L': 'i1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8', 'E': '>f4', 'D': '>f8', 'C': '>c8', 'M': '>c16', 'X': 'u1'}


def _elements(rep, letter):
	"""array elements per row of a column: the repeat count, except for bit arrays (nX: n bits in (n + 7) // 8 bytes)"""
	return (rep + 7) // 8 if letter == 'X' else rep


# the unsigned-integer convention: TFORM letter -> (numpy dtype of the values, TZERO)
_UNSIGNED = {'I': ('u2', 2 ** 15), 'J': ('u4', 2 ** 31), 'K': ('u8', 2 ** 63), 'B': ('i1', -128)}


class Column(object):
	"""name + FITS format + array; the array is converted to the format's type on construction, as
	astropy's Column does (so that a later in-place change of the source array does not leak in).  An unsigned array
	(uint16 / uint32 / uint64; int8 for B) in an I / J / K / B column keeps its values and is written with the TZERO of
	the unsigned-integer convention (.zero)."""

	def __init__(self, name, format, array):
		self.name = name
		self.format = format
		self.zero = None
		rep, letter = split_format(format)
		src = numpy.asarray(array)
		if letter in _UNSIGNED and rep == 1 and src.dtype == numpy.dtype(_UNSIGNED[letter][0]):
			self.array = numpy.array(src)
			self.zero = _UNSIGNED[letter][1]
		elif letter in _FORMATS and _elements(rep, letter) > 1:
			self.array = numpy.array(src, dtype=numpy.dtype(_FORMATS[letter]).newbyteorder('='))
			if self.array.ndim != 2 or self.array.shape[1] != _elements(rep, letter):
				raise ValueError('column "%s" (%s) needs an array of shape (rows, %d)' % (name, format, _elements(rep, letter)))
		else:
			self.array = numpy.array(src, dtype=native_dtype(format))

	def __repr__(self):
		return 'Column(%r, %r, %d rows)' % (self.name, self.format, len(self.array))


class Table(object):
	"""one BINTABLE extension: name (EXTNAME), header (OrderedDict keyword -> value), column names / TFORMs,
	data (numpy structured array, native byte order)"""

	def __init__(self, name, header, columns, formats, data, skipped=()):
		self.name = name
		self.header = header
		self.columns = columns
		self.formats = formats
		self.data = data
		self.skipped = list(skipped)   # variable-length columns left out: (name, TFORM)

	def __len__(self):
		return len(self.data)


def split_format(fmt):
	"""'1E' -> (1, 'E'); '12A' -> (12, 'A')"""
	fmt = fmt.strip()
	k = 0
	while k < len(fmt) and fmt[k].isdigit():
		k += 1
	return (int(fmt[:k]) if k else 1), fmt[k:k + 1]


def native_dtype(fmt):
	rep, letter = split_format(fmt)
	if letter == 'A':
		return numpy.dtype('S%d' % rep)
	if letter == 'L':
		return numpy.dtype(bool)
	if letter not in _FORMATS:
		raise ValueError('unsupported FITS column format "%s"' % fmt)
	base = numpy.dtype(_FORMATS[letter]).newbyteorder('=')
	n = _elements(rep, letter)
	return base if n == 1 else numpy.dtype((base, (n,)))


def _disk_dtype(fmt):
	rep, letter = split_format(fmt)
	if letter == 'A':
		return numpy.dtype('S%d' % rep)
	if letter in ('P', 'Q'):   # variable-length array descriptor (length, heap offset): read over, not interpreted
		return numpy.dtype('V%d' % (rep * (8 if letter == 'P' else 16)))
	if letter not in _FORMATS:
		raise ValueError('unsupported FITS column format "%s"' % fmt)
	base = numpy.dtype(_FORMATS[letter])
	n = _elements(rep, letter)
	return base if n == 1 else numpy.dtype((base, (n,)))


def _parse_value(v):
	v = v.strip()
	if v.startswith("'"):
		end = 1
		out = ''
		while end < len(v):   # '' is an escaped quote
			if v[end] == "'":
				if end + 1 < len(v) and v[end + 1] == "'":
					out += "'"
					end += 2
					continue
				break
			out += v[end]
			end += 1
		return out.rstrip()
	v = v.split('/')[0].strip()
	if v == 'T':
		return True
	if v == 'F':
		return False
	try:
		return int(v)
	except ValueError:
		pass
	try:
		return float(v.replace('D', 'E'))
	except ValueError:
		return v


def _read_header(buf, pos):
	cards = OrderedDict()
	last = None
	while True:
		block = buf[pos:pos + BLOCK]
		if len(block) < BLOCK:
			raise ValueError('truncated FITS header')
		pos += BLOCK
		for i in range(36):
			c = block[i * 80:(i + 1) * 80].decode('ascii', 'replace')
			k = c[:8].strip()
			if k == 'END':
				return cards, pos
			if c[8:10] == '= ':
				cards[k] = _parse_value(c[10:])
				last = k
			elif k == 'CONTINUE' and last is not None and isinstance(cards[last], str):
				# long-string convention: 'first part&' / CONTINUE  'next part'
				prev = cards[last]
				cards[last] = (prev[:-1] if prev.endswith('&') else prev) + _parse_value(c[8:])
			elif k in ('COMMENT', 'HISTORY'):
				cards.setdefault(k, []).append(c[8:].rstrip())


def _data_size(cards):
	naxis = int(cards.get('NAXIS', 0))
	if naxis == 0:
		return 0
	size = abs(int(cards['BITPIX'])) // 8
	for i in range(1, naxis + 1):
		size *= int(cards['NAXIS%d' % i])
	return size + int(cards.get('PCOUNT', 0))


def _file_bytes(path):
	with open(path, 'rb') as f:
		buf = f.read()
	if buf[:2] == b'\x1f\x8b':   # gzip, whatever the file is called
		import gzip
		buf = gzip.decompress(buf)
	return buf


def read_table(path, ext=1):
	"""the BINTABLE in extension `ext` of a FITS file (gzip-compressed files are read as astropy reads them: transparently)"""
	buf = _file_bytes(path)
	pos = 0
	ihdu = 0
	while pos < len(buf):
		cards, pos = _read_header(buf, pos)
		size = _data_size(cards)
		if ihdu == ext:
			if str(cards.get('XTENSION', '')).strip() != 'BINTABLE':
				raise ValueError('%s: extension %d is not a binary table' % (path, ext))
			nf = int(cards['TFIELDS'])
			names = [str(cards['TTYPE%d' % i]) for i in range(1, nf + 1)]
			formats = [str(cards['TFORM%d' % i]) for i in range(1, nf + 1)]
			disk = numpy.dtype([(n, _disk_dtype(f)) for n, f in zip(names, formats)])
			if disk.itemsize != int(cards['NAXIS1']):
				raise ValueError('%s: row width %d does not match the column formats (%d)' % (path, int(cards['NAXIS1']), disk.itemsize))
			nrows = int(cards['NAXIS2'])
			raw = numpy.frombuffer(buf, dtype=disk, count=nrows, offset=pos)
			keep, out_formats, values, skipped = [], [], {}, []
			for i, (n, f) in enumerate(zip(names, formats)):
				rep, letter = split_format(f)
				if letter in ('P', 'Q'):
					skipped.append((n, f))
					continue
				scale, zero = cards.get('TSCAL%d' % (i + 1), 1), cards.get('TZERO%d' % (i + 1), 0)
				if letter == 'L':
					v = raw[n] == ord('T')
				elif letter != 'A' and (scale != 1 or zero != 0):
					if letter in _UNSIGNED and scale == 1 and zero == _UNSIGNED[letter][1]:
						# unsigned-integer convention: exact integer arithmetic, no detour through floating point
						udt = numpy.dtype(_UNSIGNED[letter][0])
						v = (raw[n].astype(numpy.dtype(_FORMATS[letter]).newbyteorder('=')).view(udt) ^ udt.type(1 << (8 * udt.itemsize - 1))) if letter != 'B' \
							else (raw[n].astype('u1') ^ numpy.uint8(128)).view('i1')
					else:
						v = raw[n].astype(numpy.float64) * scale + zero
						f = ('%d' % rep if rep != 1 else '') + 'D'
				else:
					v = raw[n].astype(native_dtype(f).base) if letter != 'A' else raw[n]
				keep.append(n)
				out_formats.append(f)
				values[n] = v
			native = numpy.dtype([(n, values[n].dtype, values[n].shape[1:]) for n in keep])
			data = numpy.empty(nrows, dtype=native)
			for n in keep:
				data[n] = values[n]
			names, formats = keep, out_formats
			return Table(str(cards.get('EXTNAME', '')), cards, names, formats, data, skipped)
		pos += (size + BLOCK - 1) // BLOCK * BLOCK
		ihdu += 1
	raise ValueError('%s: no extension %d' % (path, ext))


def _cards(key, value):
	"""one card, or several for a string too long for one (CONTINUE long-string convention)"""
	if isinstance(value, str) and len(value.replace("'", "''")) > 67:
		out = []
		rest = value
		first = True
		while rest:
			# keep escaped quotes together and leave room for the '&' marker
			piece = rest[:60]
			rest = rest[60:]
			text = piece.replace("'", "''") + ('&' if rest else '')
			out.append((('%-8s= ' % key[:8].upper()) if first else 'CONTINUE  ') + "'%s'" % text)
			first = False
		return [c[:80].ljust(80) for c in out]
	return [_card(key, value)]


_STRUCTURAL = ('XTENSION', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'PCOUNT', 'GCOUNT', 'TFIELDS', 'EXTNAME', 'COMMENT', 'HISTORY')


def extra_header(table):
	"""the (keyword, value) pairs of a table's header that are not structural (SKYAREA and friends), to carry over
	into a rewritten copy (nway-create-shifted-catalogue.py:88-90)"""
	out = []
	for k, v in table.header.items():
		if k in _STRUCTURAL or k[:5] in ('TTYPE', 'TFORM', 'TUNIT', 'TDISP', 'TNULL', 'TSCAL', 'TZERO', 'TDIM'):
			continue
		out.append((k, v))
	return out


def _card(key, value, comment=''):
	if isinstance(value, bool):
		v = '%20s' % ('T' if value else 'F')
	elif isinstance(value, (int, numpy.integer)):
		v = '%20d' % value
	elif isinstance(value, (float, numpy.floating)):
		v = '%20s' % repr(float(value)).upper()
	else:
		s = str(value).replace("'", "''")
		v = "'%-8s'" % s[:67]
	c = '%-8s= %s' % (key[:8].upper(), v)
	if comment:
		c += ' / ' + comment
	return c[:80].ljust(80)


def _comment_cards(text, key='COMMENT'):
	out = []
	text = str(text)
	for i in range(0, max(len(text), 1), 72):
		out.append(('%-8s%s' % (key, text[i:i + 72])).ljust(80))
	return out


def _finish(cards):
	cards = list(cards) + ['END'.ljust(80)]
	while len(cards) % 36:
		cards.append(' ' * 80)
	return ''.join(cards).encode('ascii', 'replace')


def write_table(path, columns, extname, primary_header=(), table_header=(), comments=()):
	"""columns: list of Column.  primary_header / table_header: sequences of (keyword, value) pairs; comments:
	COMMENT lines of the primary header (nway.py:644-645 stores its arguments there)."""
	nrows = len(columns[0].array) if columns else 0
	disk = numpy.dtype([(c.name, _disk_dtype(c.format)) for c in columns])
	rec = numpy.empty(nrows, dtype=disk)
	for c in columns:
		if len(c.array) != nrows:
			raise ValueError('column "%s" has %d rows, expected %d' % (c.name, len(c.array), nrows))
		if split_format(c.format)[1] == 'L':
			rec[c.name] = numpy.where(c.array, ord('T'), ord('F'))
		elif getattr(c, 'zero', None) is not None:   # unsigned-integer convention: stored value = value - TZERO (flip the top bit)
			a = c.array
			rec[c.name] = (a.view('u1') ^ numpy.uint8(128)) if a.dtype.itemsize == 1 else (a ^ a.dtype.type(1 << (8 * a.dtype.itemsize - 1))).view(a.dtype.str.replace('u', 'i'))
		else:
			rec[c.name] = c.array
	now = datetime.datetime.now().isoformat()
	now = now[:now.rfind('.')] if '.' in now else now
	prim = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0), _card('EXTEND', True)]
	prim.append(_card('DATE', now))
	for k, v in primary_header:
		prim += _cards(k, v)
	for text in comments:
		prim += _comment_cards(text)
	tab = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8), _card('NAXIS', 2),
		_card('NAXIS1', disk.itemsize), _card('NAXIS2', nrows), _card('PCOUNT', 0), _card('GCOUNT', 1), _card('TFIELDS', len(columns))]
	for i, c in enumerate(columns):
		tab.append(_card('TTYPE%d' % (i + 1), c.name))
		tab.append(_card('TFORM%d' % (i + 1), c.format))
		if getattr(c, 'zero', None) is not None:
			tab.append(_card('TSCAL%d' % (i + 1), 1))
			tab.append(_card('TZERO%d' % (i + 1), c.zero))
	tab.append(_card('EXTNAME', extname))
	for k, v in table_header:
		tab += _cards(k, v)
	payload = rec.tobytes()
	pad = (-len(payload)) % BLOCK
	with open(path, 'wb') as f:
		f.write(_finish(prim))
		f.write(_finish(tab))
		f.write(payload)
		f.write(b'\0' * pad)


def set_table_keywords(path, keywords, ext=1):
	"""Set header keywords of the table in extension `ext` IN PLACE, leaving every other card and the data bytes as they are
	(what `f[1].name = ...; f[1].header[...] = ...; f.writeto(path, overwrite=True)` amounts to for nway-write-header.py:19-31).
	keywords: sequence of (keyword, value).  Returns the previous values (None where the keyword was absent)."""
	with open(path, 'rb') as f:
		buf = f.read()
	pos, ihdu = 0, 0
	while pos < len(buf):
		start = pos
		cards, pos = _read_header(buf, pos)
		if ihdu == ext:
			break
		pos += (_data_size(cards) + BLOCK - 1) // BLOCK * BLOCK
		ihdu += 1
	else:
		raise ValueError('%s: no extension %d' % (path, ext))
	lines = [buf[k:k + 80].decode('ascii', 'replace') for k in range(start, pos, 80)]
	end = [k for k, c in enumerate(lines) if c[:8].strip() == 'END'][0]
	lines = lines[:end]
	previous = []
	for key, value in keywords:
		key = key[:8].upper()
		previous.append(cards.get(key))
		if key in cards and cards[key] == value:
			continue   # says so already: the card stays as it is
		new = _cards(key, value)
		at = [k for k, c in enumerate(lines) if c[:8].strip() == key and c[8:10] == '= ']
		if at:
			k = at[0]
			if len(new) == 1 and not isinstance(value, str) and ' / ' in lines[k][10:]:
				new = [_card(key, value, lines[k][10:].split(' / ', 1)[1].rstrip())]   # a changed number keeps the card's comment
			stop = k + 1
			while stop < len(lines) and lines[stop][:8].strip() == 'CONTINUE':   # a long string spans several cards
				stop += 1
			lines[k:stop] = new
		else:
			lines += new
	with open(path, 'wb') as f:
		f.write(buf[:start])
		f.write(_finish(lines))
		f.write(buf[pos:])
	return previous
