"""CPU tests of the command-line layer: FITS table I/O, and the oracle in cli_compat mode against the golden
outputs of the UNMODIFIED reference nway.py (tests/golden/ref_cli_*.npz, made by oracle/make_golden_cli.py)."""
import os

import numpy as np
import pytest

from tests import cases, cliparity


def test_fits_roundtrip(tmp_path):
	from nway_b200 import fitsio as F
	rng = np.random.default_rng(5)
	n = 1234
	cols = [F.Column('A_ID', 'J', np.arange(n)), F.Column('x', 'E', rng.normal(size=n)), F.Column('n', 'I', rng.integers(-99, 99, n)),
		F.Column('s', '7A', [('s%d' % i).encode() for i in range(n)]), F.Column('d', 'D', rng.normal(size=n) * 1e10),
		F.Column('k', 'K', 2**40 + np.arange(n)), F.Column('b', 'L', rng.integers(0, 2, n) > 0), F.Column('u', 'B', rng.integers(0, 255, n))]
	path = str(tmp_path / 't.fits')
	long_cmd = 'nway.py ' + ' '.join("--arg%d 'value %d'" % (i, i) for i in range(20))
	F.write_table(path, cols, 'NWAYMATCH', primary_header=[('METHOD', 'NWAY multi-way matching'), ('NWAYCMD', long_cmd), ('N', 3), ('X', 2.5)],
		table_header=[('SKYAREA', 2.0)], comments=['argument radius: 20', 'x' * 200])
	assert os.path.getsize(path) % 2880 == 0
	t = F.read_table(path)
	assert t.name == 'NWAYMATCH' and t.columns == [c.name for c in cols] and t.formats == [c.format for c in cols]
	assert t.header['SKYAREA'] == 2.0
	for c in cols:
		assert t.data[c.name].dtype == c.array.dtype and (t.data[c.name] == c.array).all(), c.name
	cards, _ = F._read_header(open(path, 'rb').read(), 0)
	assert cards['NWAYCMD'] == long_cmd and cards['N'] == 3 and cards['X'] == 2.5
	assert cards['COMMENT'][0].strip() == 'argument radius: 20'
	with pytest.raises(ValueError):
		F.read_table(path, ext=2)


def test_fits_gzip(tmp_path):
	"""a gzip-compressed catalogue (COSMOS.fits.gz) reads like the plain file"""
	import gzip
	from nway_b200 import fitsio as F
	plain, packed = str(tmp_path / 'c.fits'), str(tmp_path / 'c.fits.gz')
	F.write_table(plain, [F.Column('ID', 'J', np.arange(9)), F.Column('RA', 'D', np.arange(9) / 7.)], 'CAT', table_header=[('SKYAREA', 0.5)])
	with open(plain, 'rb') as f, gzip.open(packed, 'wb') as g:
		g.write(f.read())
	a, b = F.read_table(plain), F.read_table(packed)
	assert b.name == 'CAT' and b.header['SKYAREA'] == 0.5 and b.columns == a.columns and (b.data == a.data).all()


def test_fits_empty_and_bad_format(tmp_path):
	from nway_b200 import fitsio as F
	path = str(tmp_path / 'e.fits')
	F.write_table(path, [F.Column('a', 'D', np.zeros(0))], 'EMPTY')
	assert len(F.read_table(path)) == 0
	with pytest.raises(ValueError):
		F.Column('v', '3E', np.zeros(3))
	with pytest.raises(ValueError):
		F.write_table(path, [F.Column('a', 'D', np.zeros(2)), F.Column('b', 'D', np.zeros(3))], 'X')


def test_subset_fits_files(tmp_path):
	from nway_b200 import fitsio as F
	paths = cases.write_cosmos_subset_fits(str(tmp_path))
	t = F.read_table(paths['XMM'])
	assert t.name == 'XMM' and len(t) == 1797 and t.columns == ['ID', 'RA', 'DEC', 'pos_err'] and t.formats == ['J', 'D', 'D', 'E']
	assert t.header['SKYAREA'] == 2.0


@pytest.mark.parametrize('name', ['cli2', 'cli3_minprob', 'cli3_prefilter'])
def test_oracle_cli_mode_against_reference_cli(name, tmp_path):
	"""the oracle port with cli_compat=True == the real nway.py, bit for bit in the output formats"""
	got = cliparity.oracle_cli_table(name)
	cliparity.check_against_cli_digest(name, got, exact=True)


def test_fits_unsigned_scaled_and_vector_columns(tmp_path):
	"""TZERO / TSCAL as astropy applies them (unsigned-integer convention -> uint columns, written back with the same
	keywords; other scalings -> float64), fixed-length vector columns carried through, variable-length columns left out"""
	from nway_b200 import fitsio
	n = 5
	cols = [fitsio.Column('ID', 'K', np.array([0, 1, 2 ** 63, 2 ** 64 - 1, 7], dtype=np.uint64)),
		fitsio.Column('U2', 'I', np.array([0, 1, 32768, 65535, 9], dtype=np.uint16)),
		fitsio.Column('U4', 'J', np.array([0, 1, 2 ** 31, 2 ** 32 - 1, 9], dtype=np.uint32)),
		fitsio.Column('S1', 'B', np.array([-128, -1, 0, 127, 5], dtype=np.int8)),
		fitsio.Column('RA', 'D', np.arange(n) * 1.5), fitsio.Column('V', '2E', np.arange(2 * n).reshape(n, 2)),
		fitsio.Column('I', 'J', np.arange(n) - 2)]
	path = str(tmp_path / 't.fits')
	fitsio.write_table(path, cols, 'T', table_header=[('SKYAREA', 2.0)])
	t = fitsio.read_table(path)
	assert t.formats == ['K', 'I', 'J', 'B', 'D', '2E', 'J'] and t.header['TZERO1'] == 2 ** 63 and t.header['TZERO2'] == 32768
	for c in cols:
		assert t.data[c.name].dtype == c.array.dtype and np.array_equal(t.data[c.name], c.array), c.name
	# a generic scaling and a variable-length column, patched into the header / row layout by hand
	raw = bytearray(open(path, 'rb').read())
	hdr_end = raw.index(b'END' + b' ' * 77, 2880)
	card = ('%-8s= %20s' % ('TSCAL7', '0.5')).ljust(80).encode() + ('%-8s= %20s' % ('TZERO7', '10.0')).ljust(80).encode()
	assert raw[hdr_end + 80:hdr_end + 240] == b' ' * 160   # room for two more cards before the block ends
	raw[hdr_end:hdr_end + 240] = card + b'END'.ljust(80)
	open(path, 'wb').write(bytes(raw))
	t2 = fitsio.read_table(path)
	assert t2.formats[-1] == 'D' and np.array_equal(t2.data['I'], (np.arange(n) - 2) * 0.5 + 10.0)
	assert fitsio._disk_dtype('1PE(7)').itemsize == 8 and fitsio._disk_dtype('QD').itemsize == 16


def test_fits_bit_and_complex_columns(tmp_path):
	"""columns the match never looks at must not make a catalogue unreadable: bit arrays (nX, carried as their bytes) and
	complex numbers (C, M) go through the reader, the -99 fill of the merged table and the writer"""
	from nway_b200 import cli, fitsio
	n = 6
	rng = np.random.default_rng(2)
	cols = [fitsio.Column('ID', 'J', np.arange(n)), fitsio.Column('RA', 'D', rng.uniform(size=n)),
		fitsio.Column('FLAGS', '12X', rng.integers(0, 255, (n, 2))), fitsio.Column('ONEBIT', '3X', rng.integers(0, 255, n)),
		fitsio.Column('Z', 'C', rng.normal(size=n) + 1j * rng.normal(size=n)), fitsio.Column('W', '2M', rng.normal(size=(n, 2)) * (1 + 2j))]
	path = str(tmp_path / 'x.fits')
	fitsio.write_table(path, cols, 'T')
	t = fitsio.read_table(path)
	assert t.formats == ['J', 'D', '12X', '3X', 'C', '2M'] and t.header['NAXIS1'] == 4 + 8 + 2 + 1 + 8 + 32
	for c in cols:
		assert t.data[c.name].dtype == c.array.dtype and np.array_equal(t.data[c.name], c.array), c.name
	idx = {'T': np.array([2, -1, 0])}
	merged = {c.name: c.array for c in cli.merged_input_columns([t], ['T'], idx)}
	assert np.array_equal(merged['T_FLAGS'][0], cols[2].array[2]) and merged['T_Z'][1] == -99 and merged['T_RA'][2] == cols[1].array[0]
	with pytest.raises(ValueError):
		fitsio.Column('bad', '12X', np.zeros((n, 3)))


def test_write_header_sets_keywords_in_place(tmp_path, capsys):
	"""nway-write-header.py: EXTNAME and SKYAREA of a catalogue set in place -- every other card (comments included) and the
	data bytes stay; a header that outgrows its 2880-byte block moves the data; an unchanged value changes nothing"""
	from nway_b200 import calibrate_cli, fitsio
	n = 7
	path = str(tmp_path / 'cat.fits')
	fitsio.write_table(path, [fitsio.Column('ID', 'J', np.arange(n)), fitsio.Column('RA', 'D', np.arange(n) * 0.5)], 'OLDNAME',
		table_header=[('OBSERVER', "it's me"), ('EXPTIME', 12.5)])
	raw = open(path, 'rb').read()
	assert calibrate_cli.write_header_main([path, 'XMM', '2']) == 0
	assert capsys.readouterr().out.splitlines() == ['current OLDNAME SKYAREA: None', 'new     XMM SKYAREA: 2.0']
	t = fitsio.read_table(path)
	assert t.name == 'XMM' and t.header['SKYAREA'] == 2.0 and t.header['OBSERVER'] == "it's me" and t.header['EXPTIME'] == 12.5
	assert np.array_equal(t.data['RA'], np.arange(n) * 0.5) and len(open(path, 'rb').read()) == len(raw)
	once = open(path, 'rb').read()
	assert calibrate_cli.write_header_main([path, 'XMM', '2.0']) == 0 and open(path, 'rb').read() == once
	# a changed number keeps its comment; a long string takes CONTINUE cards; enough new cards push the data into the next block
	fitsio.set_table_keywords(path, [('XTENSION', 'BINTABLE')])   # unchanged: stays, with its comment
	assert b"XTENSION= 'BINTABLE' / binary table extension" in open(path, 'rb').read()
	long_text = 'a long description ' * 8
	old = fitsio.set_table_keywords(path, [('SKYAREA', 3.5), ('NAXIS2', n), ('NOTE', long_text)] + [('KEY%d' % k, k) for k in range(40)])
	assert old[:3] == [2.0, n, None]
	t = fitsio.read_table(path)
	assert t.header['SKYAREA'] == 3.5 and t.header['NOTE'] == long_text.rstrip() and t.header['KEY39'] == 39
	assert np.array_equal(t.data['ID'], np.arange(n)) and len(open(path, 'rb').read()) == len(raw) + 2880
	fitsio.set_table_keywords(path, [('NOTE', 'short')])
	assert fitsio.read_table(path).header['NOTE'] == 'short' and b'CONTINUE' not in open(path, 'rb').read()
	assert calibrate_cli.write_header_main([path]) == 1
	with pytest.raises(AssertionError):
		calibrate_cli.write_header_main([path, 'X_Y', '1'])
