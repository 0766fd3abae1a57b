"""CPU: a whole two-catalogue match computed on the host from the device's own per-element source
(tests/emu/rows_emu.cpp: grid, registration, pre-tests, exact separation, rows2_write, the fused one-exponential group
normalisation -- nwb_grid.cuh / nwb_rows.cuh / nwb_device.cuh built with g++), against the oracle with the parity
metric of the GPU tests: identical row set and order, integer columns exact, floats within 1e-10.  The GPU tests do the
same through the real kernels; this one runs where there is no GPU and covers everything but the kernels'
orchestration (queues, slots, the in-group sort)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import nway_oracle as O
from tests import cases, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ctypes.c_void_p


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
	out = str(tmp_path_factory.mktemp('emu') / 'rows_emu.so')
	res = subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-w', '-I', os.path.join(ROOT, 'tests', 'emu'),
		'-o', out, os.path.join(ROOT, 'tests', 'emu', 'rows_emu.cpp')], capture_output=True, text=True)
	assert res.returncode == 0, res.stderr[-3000:]
	lib = ctypes.CDLL(out)
	lib.nwb_emu_match2.restype = ctypes.c_longlong
	return lib


def host_match(emu, tables, radius, completeness):
	import nway_b200
	tab = nway_b200._scalar_tables(tables, completeness, nway_b200.NullOutputLogger())   # the host scalars the library gets too
	a = [np.ascontiguousarray(tables[0][k], dtype=np.float64) for k in ('ra', 'dec', 'error')]
	b = [np.ascontiguousarray(tables[1][k], dtype=np.float64) for k in ('ra', 'dec', 'error')]
	from nway_b200 import magnitudeweights
	mags = []
	for magvals, maghist in zip(tables[1].get('mags', []), tables[1].get('maghists', [])):
		m = np.array(magvals, dtype=np.float64)
		m[m == -99] = np.nan                                       # nway_b200.nway_match does the same (__init__.py:318-319)
		lo, hi, hs, ha = maghist
		e, w, bv = magnitudeweights.step_tables(np.array(list(lo) + [hi[-1]]), hs, ha)
		mags.append([np.ascontiguousarray(x, dtype=np.float64) for x in (m, e, w, bv)])
	nmag = len(mags)
	PP = P * max(nmag, 1)
	nb = (ctypes.c_int * max(nmag, 1))(*[len(x[2]) for x in mags])
	cap = 64
	while True:
		cols = [np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap), np.zeros(cap), np.zeros(cap, dtype=np.int64),
			np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, dtype=np.int64), np.zeros(cap), np.zeros(cap)]
		bias_cols = [np.zeros(cap) for _ in range(nmag)]
		norm, prior, l10p = [np.ascontiguousarray(tab[k], dtype=np.float64) for k in ('norm', 'prior', 'log10prior')]
		R = emu.nwb_emu_match2(len(a[0]), P(a[0].ctypes.data), P(a[1].ctypes.data), P(a[2].ctypes.data),
			len(b[0]), P(b[0].ctypes.data), P(b[1].ctypes.data), P(b[2].ctypes.data), ctypes.c_double(radius),
			P(norm.ctypes.data), ctypes.c_double(tab['log10e']), P(prior.ctypes.data), P(l10p.ctypes.data), ctypes.c_double(0.5),
			nmag, PP(*[x[0].ctypes.data for x in mags]), nb, PP(*[x[1].ctypes.data for x in mags]), PP(*[x[2].ctypes.data for x in mags]),
			PP(*[x[3].ctypes.data for x in mags]), PP(*[c.ctypes.data for c in bias_cols]),
			ctypes.c_longlong(cap), *[P(c.ctypes.data) for c in cols])
		if R >= 0:
			break
		cap = -R - 1 + 16
	na, nb_name = tables[0]['name'], tables[1]['name']
	names = [na, nb_name, 'Separation_%s_%s' % (na, nb_name), 'Separation_max', 'ncat', 'dist_bayesfactor_uncorrected', 'dist_bayesfactor',
		'dist_post', 'p_single', 'match_flag', 'prob_has_match', 'prob_this_match']
	out = {n: c[:R] for n, c in zip(names, cols)}
	for magname, c in zip(tables[1].get('magnames', []), bias_cols):
		out['bias_%s_%s' % (nb_name, magname)] = c[:R]
	return out


@pytest.mark.parametrize('name', ['syn2', 'syn2_sparse', 'syn2_maghist', 'cosmos2', 'allsky2'])
def test_host_build_of_the_device_source_matches_the_oracle(emu, name):
	spec = cases.GOLDEN_CASES[name]
	got = host_match(emu, cases.build_case(name), spec['radius'], spec['completeness'])
	ref = O.nway_match(cases.build_case(name), spec['radius'], spec['completeness'])
	cols = [c for c in ref if not c.startswith('_')]
	assert sorted(got.keys()) == sorted(cols)
	parity.assert_tables_match(ref, got, columns=cols, context='host emulation / ' + name)
	parity.check_against_digest(name, got, [t['name'] for t in cases.build_case(name)])   # and the reference's own output


def test_bench_workload_sample(emu):
	"""the C3 generator of bench.py at 1/50 of the area: the configuration the throughput is quoted on"""
	tables = cases.config_c3(scale=0.02, seed=20260301)
	got = host_match(emu, tables, 5.0, 0.9)
	ref = O.nway_match(cases.config_c3(scale=0.02, seed=20260301), 5.0, 0.9)
	parity.assert_tables_match(ref, got, columns=[c for c in ref if not c.startswith('_')], context='host emulation / C3 sample')
	assert len(got['A']) > 100000


# ---- three and four catalogues (tests/emu/rowsn_emu.cpp) ---------------------------------------------------

@pytest.fixture(scope='module')
def emun(tmp_path_factory):
	out = str(tmp_path_factory.mktemp('emu') / 'rowsn_emu.so')
	res = subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-w', '-I', os.path.join(ROOT, 'tests', 'emu'),
		'-o', out, os.path.join(ROOT, 'tests', 'emu', 'rowsn_emu.cpp')], capture_output=True, text=True)
	assert res.returncode == 0, res.stderr[-3000:]
	lib = ctypes.CDLL(out)
	lib.nwb_emu_matchn.restype = ctypes.c_longlong
	return lib


def host_match_n(emun, tables, radius, completeness, ratio_secondary=0.5, flat_hash=False):
	import nway_b200
	from nway_b200 import magnitudeweights
	nc = len(tables)
	npair = nc * (nc - 1) // 2
	tab = nway_b200._scalar_tables(tables, completeness, nway_b200.NullOutputLogger())
	keep = []   # arrays that must stay alive during the call

	def arr(x, dtype=np.float64):
		a = np.ascontiguousarray(x, dtype=dtype)
		keep.append(a)
		return a

	def pp(arrays):
		return (P * max(len(arrays), 1))(*[a.ctypes.data for a in arrays])

	ra, dec, err = [[arr(t[k]) for t in tables] for k in ('ra', 'dec', 'error')]
	n = (ctypes.c_int * nc)(*[len(t['ra']) for t in tables])
	mag_cat, mags, edges, weights, biases, names_bias = [], [], [], [], [], []
	for c, t in enumerate(tables):
		for magvals, maghist, magname in zip(t.get('mags', []), t.get('maghists', []), t.get('magnames', [])):
			m = np.array(magvals, dtype=np.float64)
			m[m == -99] = np.nan
			lo, hi, hs, ha = maghist
			e, w, bv = magnitudeweights.step_tables(np.array(list(lo) + [hi[-1]]), hs, ha)
			mag_cat.append(c); mags.append(arr(m)); edges.append(arr(e)); weights.append(arr(w)); biases.append(arr(bv))
			names_bias.append('bias_%s_%s' % (t['name'], magname))
	nmag = len(mags)
	cap = 1024
	while True:
		idx = [np.zeros(cap, dtype=np.int64) for _ in range(nc)]
		seps = [np.zeros(cap) for _ in range(npair)]
		bias_cols = [np.zeros(cap) for _ in range(nmag)]
		sepmax, lbf_u, lbf, post, ps, pany, pi = [np.zeros(cap) for _ in range(7)]
		ncat, flag = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
		norm, prior, l10p = arr(tab['norm']), arr(tab['prior']), arr(tab['log10prior'])
		R = emun.nwb_emu_matchn(nc, n, pp(ra), pp(dec), pp(err), ctypes.c_double(radius), P(norm.ctypes.data), ctypes.c_double(tab['log10e']),
			P(prior.ctypes.data), P(l10p.ctypes.data), ctypes.c_double(ratio_secondary), nmag, (ctypes.c_int * max(nmag, 1))(*mag_cat), pp(mags),
			(ctypes.c_int * max(nmag, 1))(*[len(w) for w in weights]), pp(edges), pp(weights), pp(biases), pp(bias_cols), int(flat_hash), ctypes.c_longlong(cap),
			pp(idx), pp(seps), P(sepmax.ctypes.data), P(ncat.ctypes.data), P(lbf_u.ctypes.data), P(lbf.ctypes.data), P(post.ctypes.data),
			P(ps.ctypes.data), P(flag.ctypes.data), P(pany.ctypes.data), P(pi.ctypes.data))
		if R >= 0:
			break
		cap = -R - 1 + 16
	names = [t['name'] for t in tables]
	out = {}
	for c in range(nc):
		out[names[c]] = idx[c][:R]
	k = 0
	for a in range(nc):
		for b in range(a + 1, nc):
			out['Separation_%s_%s' % (names[a], names[b])] = seps[k][:R]
			k += 1
	out.update(Separation_max=sepmax[:R], ncat=ncat[:R], dist_bayesfactor_uncorrected=lbf_u[:R], dist_bayesfactor=lbf[:R], dist_post=post[:R])
	for name, col in zip(names_bias, bias_cols):
		out[name] = col[:R]
	out.update(p_single=ps[:R], match_flag=flag[:R], prob_has_match=pany[:R], prob_this_match=pi[:R])
	return out


@pytest.mark.parametrize('name', ['syn2', 'syn3', 'syn3_pcvec', 'syn4', 'cosmos3', 'allsky3'])
def test_host_build_of_the_device_source_matches_the_oracle_n_catalogues(emun, name):
	spec = cases.GOLDEN_CASES[name]
	kw = spec.get('kwargs', {})
	got = host_match_n(emun, cases.build_case(name), spec['radius'], spec['completeness'], ratio_secondary=kw.get('prob_ratio_secondary', 0.5))
	ref = O.nway_match(cases.build_case(name), spec['radius'], spec['completeness'], **kw)
	cols = [c for c in ref if not c.startswith('_')]
	assert sorted(got.keys()) == sorted(cols)
	parity.assert_tables_match(ref, got, columns=cols, context='host emulation / ' + name)
	parity.check_against_digest(name, got, [t['name'] for t in cases.build_case(name)])


@pytest.mark.parametrize('seed', [1003, 1019, 1071, 1038])
def test_flat_hash_switch_reproduces_the_reference_rows(emun, seed):
	"""the planned NWB_COMPAT_FLAT_HASH predicate (flat_hash_cell / flat_hash_same_bucket, nwb_grid.cuh) applied inside the
	emulated row enumeration: the table equals the oracle's with the reference's own flat-sky hash -- fewer rows than
	the complete search at mid-latitudes (1003, 1019, 1071), the same near the equator"""
	from tests.test_gpu_fuzz import random_case
	tables, radius, pc, kw, kind = random_case(seed)
	assert not kw and O.flat_sky_applicable([(t['ra'], t['dec']) for t in tables], radius / 3600)
	got = host_match_n(emun, [dict(t) for t in tables], radius, pc, flat_hash=True)
	ref = O.nway_match([dict(t) for t in tables], radius, pc, enumerator='refhash')
	cols = [c for c in ref if not c.startswith('_')]
	parity.assert_tables_match(ref, got, columns=cols, context='flat-hash switch / seed %d' % seed, rtol=3e-9)
	full = host_match_n(emun, [dict(t) for t in tables], radius, pc)
	assert len(full[tables[0]['name']]) >= len(got[tables[0]['name']])
