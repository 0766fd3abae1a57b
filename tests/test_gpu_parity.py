"""GPU: the CUDA path (through the C ABI, via nway_b200.nway_match) against the oracle run live on the same
inputs, and against the committed outputs of the real reference."""
import os

import numpy as np
import pytest

from tests import cases, parity

pytestmark = pytest.mark.gpu


def run_cuda(tables, radius, completeness, **kw):
	import nway_b200
	return nway_b200.nway_match(tables, radius, completeness, logger=nway_b200.NullOutputLogger(),
		store_mag_hists=False, as_frame=False, **kw)


def report(name, lines):
	out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
	if os.path.isdir(out):
		with open(os.path.join(out, 'parity_report.txt'), 'a') as f:
			f.write('== %s\n%s\n' % (name, '\n'.join(lines)))


@pytest.mark.parametrize('name', list(cases.GOLDEN_CASES))
def test_cuda_matches_oracle_and_reference(name):
	from oracle import nway_oracle as O
	spec = cases.GOLDEN_CASES[name]
	kw = spec.get('kwargs', {})
	got = run_cuda(cases.build_case(name), spec['radius'], spec['completeness'], **kw)
	tables = cases.build_case(name)
	ref = O.nway_match(tables, spec['radius'], spec['completeness'], enumerator=spec.get('enumerator', 'reference'), **kw)
	cols = [c for c in ref if not c.startswith('_')]
	assert list(got.keys()) == cols, (list(got.keys()), cols)
	# every case at the north star's tolerance (tests/parity.py: 1e-10 relative), all-sky and off-equator included
	lines = parity.assert_tables_match(ref, got, columns=cols, context=name)
	report(name, lines)
	parity.check_against_digest(name, got, [t['name'] for t in tables])


@pytest.mark.parametrize('name', ['offeq2', 'offeq3'])
def test_flat_hash_switch(name):
	"""NWB_COMPAT_FLAT_HASH on (default): the rows of the UNMODIFIED reference on a flat-sky field away from the equator
	(committed golden, fastskymatch.py:94-101,123-133 leaves pairs out there); off: the complete enumeration, a strict
	superset with identical per-row columns; and the switch is inert where the reference hashes with HEALPix."""
	import nway_b200
	from oracle import nway_oracle as O
	spec = cases.GOLDEN_CASES[name]
	names = [t['name'] for t in cases.build_case(name)]
	on = run_cuda(cases.build_case(name), spec['radius'], spec['completeness'])
	assert nway_b200._lib.get_context().flat_hash_applied()
	parity.check_against_digest(name, on, names)
	off = run_cuda(cases.build_case(name), spec['radius'], spec['completeness'], flat_hash_compat=False)
	assert not nway_b200._lib.get_context().flat_hash_applied()
	full = O.nway_match(cases.build_case(name), spec['radius'], spec['completeness'], enumerator='complete')
	cols = [c for c in full if not c.startswith('_')]
	report(name + '/complete', parity.assert_tables_match(full, off, columns=cols, context=name + '/complete'))
	assert len(off[names[0]]) > len(on[names[0]])
	rows_off = {tuple(r): k for k, r in enumerate(np.stack([off[n] for n in names], axis=1).tolist())}
	pos = np.array([rows_off[tuple(r)] for r in np.stack([on[n] for n in names], axis=1).tolist()])   # KeyError = not a subset
	for c in cols:
		if c.startswith('Separation') or c.startswith('dist_') or c == 'ncat':
			assert np.array_equal(np.asarray(on[c]), np.asarray(off[c])[pos], equal_nan=True), c   # per-row columns: same bits
	# a whole-sky field: the reference uses its HEALPix hash, the switch changes nothing
	a = run_cuda(cases.build_case('allsky2'), 120, 0.9)
	assert not nway_b200._lib.get_context().flat_hash_applied()
	b = run_cuda(cases.build_case('allsky2'), 120, 0.9, flat_hash_compat=False)
	for c in a:
		assert np.array_equal(np.asarray(a[c]), np.asarray(b[c]), equal_nan=True), c


@pytest.mark.parametrize('mode', ['cli'])
def test_cli_unrelated_correction(mode):
	from oracle import nway_oracle as O
	for name, radius in (('syn3', 8), ('syn4', 6), ('cosmos3', 20)):
		got = run_cuda(cases.build_case(name), radius, 0.9, unrelated_mode=mode)
		ref = O.nway_match(cases.build_case(name), radius, 0.9, unrelated_mode=mode)
		cols = [c for c in ref if not c.startswith('_')]
		changed = (ref['dist_bayesfactor'] != ref['dist_bayesfactor_uncorrected']).sum()
		assert changed > 0
		report(name + '/cli', parity.assert_tables_match(ref, got, columns=cols, context=name + '/cli'))


def test_allsky_poles_and_wraparound():
	"""complete on the whole sphere: primaries at the poles, across ra = 0/360 and everywhere else; the oracle's
	KD-tree enumerator is the all-sky stand-in for the reference's HEALPix branch (healpy is un-vendored)."""
	from oracle import nway_oracle as O
	rng = np.random.default_rng(77)
	tables = cases.allsky(78, (3000, 200000, 150000), (5.0, 3.0, 4.0))
	# force difficult primaries
	p = tables[0]
	p['ra'][:6] = [0.0, 359.99999, 0.00001, 123.0, 250.0, 180.0]
	p['dec'][:6] = [0.0, 10.0, -45.0, 89.9999, -89.99995, 90.0]
	# secondaries clustered around them
	for t in tables[1:]:
		n = 400
		k = rng.integers(0, 6, n)
		rad = np.radians(rng.uniform(0, 0.12, n))
		ang = rng.uniform(0, 2 * np.pi, n)
		dec = p['dec'][k] + np.degrees(rad * np.cos(ang))
		over = dec > 90
		dec[over] = 180 - dec[over]
		under = dec < -90
		dec[under] = -180 - dec[under]
		cosd = np.maximum(np.cos(np.radians(p['dec'][k])), 1e-6)
		ra = (p['ra'][k] + np.degrees(rad * np.sin(ang)) / cosd + np.where(over | under, 180, 0)) % 360
		t['ra'][:n] = ra
		t['dec'][:n] = dec
	radius = 300.0
	got = run_cuda(tables, radius, 0.8)
	ref = O.nway_match(tables, radius, 0.8)
	cols = [c for c in ref if not c.startswith('_')]
	report('allsky', parity.assert_tables_match(ref, got, columns=cols, context='allsky'))
	assert (np.bincount(got['A'], minlength=6)[:6] > 1).all(), 'the pole / wrap primaries must have found counterparts'


def test_ragged_and_degenerate_inputs():
	from oracle import nway_oracle as O
	import nway_b200
	# a primary catalogue of one source; empty secondary lists for most; a secondary exactly on top of a primary
	tables = cases.uniform_patch(3, (1, 50), (1.0, 0.5), 0.01)
	tables[1]['ra'][0] = tables[0]['ra'][0]
	tables[1]['dec'][0] = tables[0]['dec'][0]
	got = run_cuda(tables, 10.0, 0.9)
	ref = O.nway_match(tables, 10.0, 0.9)
	parity.assert_tables_match(ref, got, columns=[c for c in ref if not c.startswith('_')], context='single primary')
	# no secondary anywhere near: every group is the lone no-counterpart row (flag 1, p_any 0)
	tables = cases.uniform_patch(4, (20, 30), (1.0, 0.5), 0.01)
	tables[1]['dec'] = tables[1]['dec'] + 5.0
	got = run_cuda(tables, 3.0, 0.9)
	assert len(got['A']) == 20 and (got['B'] == -1).all() and (got['match_flag'] == 1).all() and (got['prob_has_match'] == 0).all()
	# an empty secondary catalogue (the reference itself raises IndexError here, __init__.py:144-146 / SURVEY Q4;
	# we return the 2-catalogue row set with the third column absent everywhere)
	tables = cases.uniform_patch(4, (20, 30, 0), (1.0, 0.5, 0.5), 0.01)
	got = run_cuda(tables, 30.0, 0.9)
	two = run_cuda(cases.uniform_patch(4, (20, 30), (1.0, 0.5), 0.01), 30.0, 0.9)
	assert (got['C'] == -1).all() and np.isnan(got['Separation_A_C']).all() and np.isnan(got['Separation_B_C']).all()
	assert (got['A'] == two['A']).all() and (got['B'] == two['B']).all()
	assert np.array_equal(got['Separation_A_B'], two['Separation_A_B'], equal_nan=True)


@pytest.mark.parametrize('ncat', [2, 3])
def test_clustered_field_spills_and_big_groups(ncat):
	"""a few primaries sit in dense clumps of secondaries: many more matches than the per-primary slots sized from
	the mean density (spill path), and groups far beyond the shared-memory fast path of the 2-catalogue row kernel"""
	from oracle import nway_oracle as O
	rng = np.random.default_rng(5)
	counts = (400, 3000, 2500)[:ncat]
	tables = cases.uniform_patch(31, counts, (1.0, 0.4, 0.6)[:ncat], 0.5)
	for c in range(1, ncat):
		t = tables[c]
		k = 0
		for prim, m in ((3, 700 if ncat == 2 else 60), (77, 150 if ncat == 2 else 25), (200, 40)):
			ang = rng.uniform(0, 2 * np.pi, m)
			rad = 6.0 / 3600 * np.sqrt(rng.uniform(size=m))
			t['ra'][k:k + m] = tables[0]['ra'][prim] + rad * np.cos(ang)
			t['dec'][k:k + m] = tables[0]['dec'][prim] + rad * np.sin(ang)
			k += m
		# shuffle so the clump members are spread over the index range
		perm = rng.permutation(len(t['ra']))
		t['ra'], t['dec'] = t['ra'][perm], t['dec'][perm]
	got = run_cuda(tables, 5.0, 0.9)
	ref = O.nway_match(tables, 5.0, 0.9)
	report('clustered%d' % ncat, parity.assert_tables_match(ref, got, columns=[c for c in ref if not c.startswith('_')], context='clustered'))
	assert np.bincount(got['A']).max() > 129


def test_extreme_bayes_factors():
	"""positional errors far smaller than the separations: log BFs of -1e5, every exponential under/overflows; and
	the opposite, errors far larger than the radius"""
	from oracle import nway_oracle as O
	for sig, seed in (((0.01, 0.005), 61), ((300.0, 200.0), 62), ((1e-4, 1.0), 63)):
		tables = cases.uniform_patch(seed, (200, 20000), sig, 0.05)
		got = run_cuda(tables, 8.0, 0.9)
		ref = O.nway_match(tables, 8.0, 0.9)
		report('extreme%s' % (sig,), parity.assert_tables_match(ref, got, columns=[k for k in ref if not k.startswith('_')], context='extreme %s' % (sig,)))


def test_back_to_back_matches_reuse_and_invalidate_cached_state():
	"""the context keeps the grid geometry and the output capacity of the previous match and launches the row kernel
	speculatively; a second match with the same sizes but moved / denser catalogues must notice and redo"""
	from oracle import nway_oracle as O
	a = cases.uniform_patch(51, (300, 20000), (1.0, 0.3), 0.1)
	b = cases.uniform_patch(52, (300, 20000), (1.0, 0.3), 0.1, ra0=150.3, dec0=0.2)     # same shapes, elsewhere
	c = cases.uniform_patch(53, (300, 60000), (1.0, 0.3), 0.1, ra0=150.3, dec0=0.2)     # same primaries' box, 3x the rows
	for tables in (a, a, b, c, a):
		got = run_cuda(tables, 6.0, 0.9)
		ref = O.nway_match(tables, 6.0, 0.9)
		parity.assert_tables_match(ref, got, columns=[k for k in ref if not k.startswith('_')], context='back-to-back')


def test_sharded_primary_ranges_concatenate_to_the_full_table():
	"""groups never span shards, so per-shard tables concatenate to the single-device table."""
	tables = cases.build_case('syn3')
	full = run_cuda(tables, 8, 0.9)
	n0 = len(tables[0]['ra'])
	parts = []
	for first, count in ((0, 100), (100, 257), (357, n0 - 357)):
		parts.append(run_cuda(cases.build_case('syn3'), 8, 0.9, primary_range=(first, count)))
	for c in full:
		cat = np.concatenate([p[c] for p in parts])
		assert len(cat) == len(full[c])
		assert ((cat == full[c]) | (np.isnan(cat.astype(float)) & np.isnan(full[c].astype(float)))).all(), c


@pytest.mark.parametrize('ncat,mode', [(2, 'api'), (3, 'api'), (3, 'cli'), (4, 'cli')])
def test_elliptical_errors(ncat, mode):
	"""the CLI's elliptical Bayes factor (nway.py:346-354 / bayesdistance.py:207-240); the oracle's restatement is
	pinned to the real log_bf_elliptical / convert_from_ellipse (tests/golden/kat.npz); the tangent-plane offsets
	(astropy in the reference) are the closed form of SURVEY.md A.6 on both sides"""
	import nway_b200
	from oracle import nway_oracle as O
	def build():
		rng = np.random.default_rng(17)
		tables = cases.uniform_patch(19, (150, 4000, 3000, 2500)[:ncat], (1.0, 0.4, 0.6, 0.5)[:ncat], 0.05, dec0=35.0)
		for k, t in enumerate(tables):
			n = len(t['ra'])
			if k == 2:
				continue   # one catalogue keeps a circular error column
			a = rng.uniform(0.5, 2.0, n)
			t['error'] = nway_b200.ellipse_error(a, rng.uniform(0.3, 1.0, n) * a, rng.uniform(0, 180, n))
		return tables
	got = run_cuda(build(), 6.0, 0.9, unrelated_mode=mode)
	ref = O.nway_match(build(), 6.0, 0.9, unrelated_mode=mode)
	# at the north star's tolerance: the sines and cosines inside the tangent-plane offsets carry the reference's bits, so
	# the cancelling latitude term cos d1 sin d2 - sin d1 cos d2 cos dl is reproduced exactly
	report('elliptical%d/%s' % (ncat, mode), parity.assert_tables_match(ref, got, columns=[c for c in ref if not c.startswith('_')],
		context='elliptical'))
	# and the reference's own consistency property (tests/bayesdistance_test.py:149-203): circular errors through
	# the elliptical code give the circular answer to ~7 decimals
	circ = cases.uniform_patch(19, (150, 4000, 3000)[:min(ncat, 3)], (1.0, 0.4, 0.6)[:min(ncat, 3)], 0.05)
	a = run_cuda(circ, 6.0, 0.9)
	for t in circ:
		t['error'] = (t['error'], t['error'].copy(), np.zeros(len(t['ra'])))
	b = run_cuda(circ, 6.0, 0.9)
	assert np.array_equal(a['A'], b['A']) and np.abs(a['dist_bayesfactor'] - b['dist_bayesfactor']).max() < 1e-6


def test_nccl_sharded_match_equals_single_device():
	"""one process per GPU over NCCL (2 ranks when the box has 2 GPUs, else 1): gathered table == single-device table"""
	import subprocess
	import sys
	import torch
	world = min(2, torch.cuda.device_count())
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
		'--master-port', '29631', os.path.join(root, 'tests', 'run_sharded.py')]
	res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
	assert res.returncode == 0 and 'SHARDED_OK' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_elementwise_surface():
	from nway_b200 import fastskymatch, bayesdistance
	k = parity.load_golden('kat.npz')
	d = fastskymatch.dist((k['dist_ra1'], k['dist_dec1']), (k['dist_ra2'], k['dist_dec2']))
	ref = k['dist_out']
	# the Vincenty form carries an ABSOLUTE error of a few ulp of sin/cos(lat) (~1e-16 rad = 6e-15 deg), whatever
	# the separation; relative 1e-9 on top
	ok = np.abs(d - ref) <= 1e-9 * np.abs(ref) + 5e-14
	assert ok.all(), (d[~ok][:5], ref[~ok][:5])
	for n in (1, 2, 3, 4):
		s, p = k['logbf%d_s' % n], k['logbf%d_p' % n]
		out = bayesdistance.log_bf([[p[i][j] for j in range(n)] for i in range(n)], list(s))
		assert np.allclose(out, k['logbf%d_out' % n], rtol=1e-12, atol=1e-11)
	out = bayesdistance.posterior(k['post_prior'], k['post_logbf'])
	assert np.allclose(out, k['post_out'], rtol=1e-10, atol=0)


def _low_level_context(tables, radius, completeness):
	import nway_b200
	from nway_b200 import _lib
	ctx = _lib.Context(0)
	_load_tables(ctx, tables, radius, completeness)
	return ctx


def _load_tables(ctx, tables, radius, completeness):
	import nway_b200
	from nway_b200 import _lib
	for c, t in enumerate(tables):
		ctx.set_catalogue(c, len(tables), t['ra'], t['dec'], t['error'], t['area'])
	tab = nway_b200._scalar_tables(tables, completeness, nway_b200.NullOutputLogger())
	ctx.set_params(radius, tab['pc'], 0.5, _lib.UNRELATED_API)
	ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])


def _all_columns(ctx, nrows):
	from nway_b200 import _lib
	sels = [_lib.COL_IDX, _lib.COL_IDX + 1, _lib.COL_SEP, _lib.COL_SEPMAX, _lib.COL_NCAT, _lib.COL_LOGBF_UNCORR, _lib.COL_LOGBF,
		_lib.COL_DIST_POST, _lib.COL_P_SINGLE, _lib.COL_MATCH_FLAG, _lib.COL_P_ANY, _lib.COL_P_I]
	out = [ctx.fetch(sel, nrows, dtype=np.int64) for sel in sels]   # bit patterns: NaNs compare equal too
	ctx.sync()
	return out


def test_async_match_equals_sync_match_and_recovers_from_failed_speculation():
	"""nwb_match_async / nwb_match_wait (no host round trip at the end of a match): same table bit for bit as nwb_match,
	also when the blindly enqueued pipeline does not hold (catalogue replaced by a shifted / a three times denser one)"""
	from nway_b200 import _lib
	a = cases.uniform_patch(51, (300, 20000), (1.0, 0.3), 0.1)
	b = cases.uniform_patch(52, (300, 20000), (1.0, 0.3), 0.1, ra0=150.3, dec0=0.2)
	c = cases.uniform_patch(53, (300, 60000), (1.0, 0.3), 0.1, ra0=150.3, dec0=0.2)
	ctx = _low_level_context(a, 6.0, 0.9)
	with pytest.raises(_lib.NwbError):
		ctx.match_wait()                      # nothing in flight, nothing matched
	ctx.match_async()                         # first match of the context: runs synchronously inside
	n0 = ctx.match_wait()
	want = _all_columns(ctx, n0)
	ref_ctx = _low_level_context(a, 6.0, 0.9)
	assert ref_ctx.match() == n0
	for x, y in zip(want, _all_columns(ref_ctx, n0)):
		assert (x == y).all()
	for _ in range(3):                        # now truly asynchronous, back to back
		ctx.match_async()
	with pytest.raises(_lib.NwbError):
		ctx.fetch(_lib.COL_P_ANY, n0)         # results may be read only after match_wait
	assert ctx.match_wait() == n0
	assert ctx.match_wait() == n0             # idempotent
	for x, y in zip(want, _all_columns(ctx, n0)):
		assert (x == y).all()
	for tables in (b, c, a):                  # moved primaries (stale grid geometry), then a table that outgrows the buffers
		_load_tables(ctx, tables, 6.0, 0.9)
		ctx.match_async()
		n = ctx.match_wait()
		ref_ctx = _low_level_context(tables, 6.0, 0.9)
		assert ref_ctx.match() == n
		for x, y in zip(_all_columns(ctx, n), _all_columns(ref_ctx, n)):
			assert (x == y).all()


def _table_columns(ctx, nrows, ncat):
	from nway_b200 import _lib
	npairs = ncat * (ncat - 1) // 2
	sels = [_lib.COL_IDX + c for c in range(ncat)] + [_lib.COL_SEP + k for k in range(npairs)] + [_lib.COL_SEPMAX, _lib.COL_NCAT,
		_lib.COL_LOGBF_UNCORR, _lib.COL_LOGBF, _lib.COL_DIST_POST, _lib.COL_P_SINGLE, _lib.COL_MATCH_FLAG, _lib.COL_P_ANY, _lib.COL_P_I]
	out = [ctx.fetch(sel, nrows, dtype=np.int64) for sel in sels]
	ctx.sync()
	return out


@pytest.mark.parametrize('ncat,mode', [(3, 'api'), (4, 'cli')])
def test_speculative_pipeline_for_three_and_more_catalogues(ncat, mode):
	"""N >= 3: the second match of a context enqueues lists, separation scratch, row count, rows and normalisation behind
	device-side gates with one synchronisation at the end (nwb_api.cu, k_spec_gate).  Same table bit for bit as the
	stage-by-stage path of a fresh context -- also when a gate stays shut: catalogues replaced by denser ones (more list
	entries and rows than the buffers of the previous match hold), by clustered ones (primaries that need the
	warp-per-primary kernels where only the thread-per-primary ones were launched), and back."""
	from nway_b200 import _lib
	sizes = (400, 9000, 7000, 6000)[:ncat]
	sig = (1.0, 0.3, 0.5, 0.4)[:ncat]
	a = cases.uniform_patch(61, sizes, sig, 0.3)                                   # sparse: at most a few matches per primary
	b = cases.uniform_patch(62, (400,) + tuple(6 * n for n in sizes[1:]), sig, 0.3)   # six times denser: buffers outgrown
	c = cases.uniform_patch(63, sizes, sig, 0.06 if ncat == 4 else 0.03)           # crowded: big groups
	unrelated = _lib.UNRELATED_CLI if mode == 'cli' else _lib.UNRELATED_API

	def load(ctx, tables):
		_load_tables(ctx, tables, 6.0, 0.9)
		import nway_b200
		tab = nway_b200._scalar_tables(tables, 0.9, nway_b200.NullOutputLogger())
		ctx.set_params(6.0, tab['pc'], 0.5, unrelated)
		ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])

	ctx = _lib.Context(0)
	for tables in (a, a, a, b, b, c, a, c, c):
		load(ctx, tables)
		n = ctx.match()
		ref_ctx = _lib.Context(0)
		load(ref_ctx, tables)
		assert ref_ctx.match() == n
		for x, y in zip(_table_columns(ctx, n, ncat), _table_columns(ref_ctx, n, ncat)):
			assert (x == y).all()
		ref_ctx.close()
	ctx.close()


def test_score_rows_of_a_supplied_candidate_list():
	"""nwb_score_rows: the caller's index tuples (here the oracle's complete enumeration plus tuples the radius filter
	would drop) scored on the device -- separations, Separation_max, ncat, log Bayes factor, dist_post -- against the
	oracle's arithmetic on the same tuples (nwaylib/__init__.py:123-196, 220-259)"""
	import nway_b200
	from nway_b200 import _lib
	from oracle import nway_oracle as O
	tables = cases.uniform_patch(31, (300, 6000, 5000), (1.0, 0.3, 0.5), 0.06, dec0=20.0)
	radius, pc = 7.0, 0.9
	run_cuda([dict(t) for t in tables], radius, pc)   # leaves catalogues, parameters and tables set on the context
	radec = [(t['ra'], t['dec']) for t in tables]
	idx = O.crossproduct_complete(radec, 2.5 * radius / 3600)   # a wider net: most of these fail the radius filter
	assert len(idx) > 5000
	got = _lib.get_context().score_rows(idx)
	n = len(tables)
	sepmax = np.zeros(len(idx))
	k = 0
	seps = {}
	for a in range(n):
		for b in range(a + 1, n):
			ia, ib = idx[:, a], idx[:, b]
			col = O.dist((radec[a][0][ia], radec[a][1][ia]), (radec[b][0][ib], radec[b][1][ib])) * 60 * 60
			col[(ia == -1) | (ib == -1)] = np.nan
			seps[(a, b)] = col
			with np.errstate(invalid='ignore'):
				sepmax = np.where(np.isnan(col), sepmax, np.maximum(col, sepmax))
			assert parity.column_error('Separation', col, got['sep'][k])[0]
			k += 1
	assert parity.column_error('Separation_max', sepmax, got['sepmax'])[0]
	assert np.array_equal(got['ncat'], (idx > -1).sum(axis=1))
	mt = dict(idx=idx, sep=seps, sepmax=sepmax, ncat=(idx > -1).sum(axis=1), errors=[np.asarray(t['error'], dtype=float)[idx[:, c]] for c, t in enumerate(tables)])
	nu, nu_plus = O.source_densities(tables)
	prior, lbf = O.single_log_bf(mt, nu, nu_plus, O.completeness_vector(pc, n))
	ok, dabs, drel, worst = parity.column_error('dist_bayesfactor', lbf, got['log_bf'])
	assert ok, (dabs, drel, worst)
	big = lbf > -300   # dist_post underflows to 0 below; relative comparison where it is a number
	ok, dabs, drel, worst = parity.column_error('dist_post', O.posterior(prior, lbf)[big], got['dist_post'][big])
	assert ok, (dabs, drel, worst)
