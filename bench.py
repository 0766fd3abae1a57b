#!/usr/bin/env python
"""bench.py -- candidate associations / second of the nway match-probability path on B200.

  python bench.py --gpus N --steps K --warmup W                 this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   the reference's own CPU path: the unmodified
                                                                nwaylib.nway_match from oracle/_ref (the oracle's port
                                                                only where that copy is missing), on a bounded sample

A "step" is one full pass of the hot path (grid -> stream secondaries -> lists -> rows + group normalisation)
over the workload.  N = 1 workload: BASELINE.json configs[2] ("C3": 1e5 x 1e7 uniform on 1 deg^2, r = 5 arcsec,
circular errors, fp64) -- the configuration BASELINE.md quotes the 1e9 associations/s target on; configs[1]
(COSMOS 3-catalogue) needs the reference's FITS files, has 1797 primaries and is launch-latency bound, so it is
a parity-test case (tests/golden), not a bench line.  N > 1: weak scaling -- every rank matches its own block
of 1e5 primaries against the (replicated) 1e7 secondaries; the only exchange of `value` is an all-gather of the per-rank
row counts (what is needed to place each shard in the global table); `with_table_allgather` is the same step with the
whole table reassembled on every rank inside the timed region (peer-memory stores, nway_b200.parallel.TableGather).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RADIUS = 5.0
COMPLETENESS = 0.9
N_PRIMARY = 100000
N_SECONDARY = 10000000


def make_workload(world, scale=1.0, seed=20260301):
	"""C3 generator of SURVEY.md 8d; `world` blocks of primaries (weak scaling), secondaries shared."""
	rng = np.random.default_rng(seed)
	side = np.sqrt(scale)
	n0, n1 = int(round(N_PRIMARY * scale)), int(round(N_SECONDARY * scale))
	pra = 150 + side * rng.uniform(size=n0)
	pdec = -side / 2 + side * rng.uniform(size=n0)
	sra = 150 + side * rng.uniform(size=n1)
	sdec = -side / 2 + side * rng.uniform(size=n1)
	for r in range(1, world):   # further primary blocks, one per extra rank
		rr = np.random.default_rng(seed + 1000 * r)
		pra = np.concatenate((pra, 150 + side * rr.uniform(size=n0)))
		pdec = np.concatenate((pdec, -side / 2 + side * rr.uniform(size=n0)))
	prim = dict(name='A', ra=pra, dec=pdec, error=1.0 * np.ones(len(pra)), area=scale, mags=[], magnames=[], maghists=[])
	sec = dict(name='B', ra=sra, dec=sdec, error=0.2 * np.ones(n1), area=scale, mags=[], magnames=[], maghists=[])
	return [prim, sec], n0


class ClockSampler(object):
	"""nvidia-smi clocks / throttle reasons while the measured loops run (B200_PROFILING.md): one background
	`nvidia-smi -lms 20` process, started before the timed region and stopped after the e2e loop."""
	FIELDS = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

	def __init__(self, device):
		self.device = device
		self.proc = None
		self.lines = []

	def start(self):
		try:
			self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.FIELDS,
				'--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
		except Exception:
			self.proc = None

	def stop(self):
		if self.proc is None:
			return
		try:
			self.proc.terminate()
			out, _ = self.proc.communicate(timeout=5)
			self.lines = [l for l in out.splitlines() if l.strip()]
		except Exception:
			pass

	def summary(self):
		samples = [[x.strip() for x in l.split(',')] for l in self.lines]
		samples = [s for s in samples if len(s) >= 7]
		if not samples:
			return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
		num = lambda v: float(v) if v.replace('.', '', 1).isdigit() else None
		sm = sorted(x for x in (num(s[0]) for s in samples) if x is not None)
		power = [x for x in (num(s[2]) for s in samples) if x is not None]
		reasons = set()
		for s in samples:
			for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
				if v.lower().startswith('active'):
					reasons.add(name)
		return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': num(samples[0][1]), 'power_w_max': max(power) if power else None,
			'reasons': sorted(reasons), 'samples': len(samples), 'window': 'timed loop + e2e loop, nvidia-smi -lms 20'}


def bind_near_gpu(local):
	"""run this rank's host threads (and, by first touch, its pinned buffers) on the CPUs NVML reports as local to the
	GPU: with one rank per GPU the host <-> device copies of the e2e loop otherwise cross the socket interconnect for
	half of the ranks.  Best effort: returns a description, or None when nothing was changed."""
	try:
		import pynvml
		import torch
		pynvml.nvmlInit()
		try:
			pr = torch.cuda.get_device_properties(local)
			h = pynvml.nvmlDeviceGetHandleByPciBusId(('%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
		except Exception:
			h = pynvml.nvmlDeviceGetHandleByIndex(local)
		ncpu = os.cpu_count() or 1
		mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
		near = set(64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1)
		allowed = os.sched_getaffinity(0)
		target = near & allowed
		if not target or target == allowed:
			return None
		os.sched_setaffinity(0, target)
		return '%d of %d host cpus (NVML affinity of the GPU)' % (len(target), len(allowed))
	except Exception:
		return None


def kernel_source_hash():
	"""sha256 over the CUDA sources of the library: ties profiles/traffic.json to the code it was captured from"""
	import hashlib
	h = hashlib.sha256()
	d = os.path.join(ROOT, 'nway_b200', 'csrc')
	for name in sorted(os.listdir(d)):
		if name.endswith(('.cu', '.cuh', '.h')):
			h.update(name.encode())
			h.update(open(os.path.join(d, name), 'rb').read())
	return h.hexdigest()


def hbm_peak():
	path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
	if os.path.exists(path):
		try:
			return float(json.load(open(path))['hbm_gbs']), 'measured'
		except Exception:
			pass
	return 6650.0, 'fallback'


WORKLOAD = 'C3 (BASELINE.json configs[2]): synthetic 2-cat, 100000 primaries per GPU x 10000000 secondaries uniform on 1 deg^2, r=5 arcsec, circular errors'


def cpu_reference(scale, steps=1, warmup=0):
	"""The reference's own CPU path on a bounded sample of the workload (same surface densities, `scale` of the area):
	the UNMODIFIED nwaylib.nway_match from oracle/_ref (pip-installed copy of the reference, oracle/install_ref.py;
	kind 'reference') -- or, where that copy is missing, the oracle's port of the same algorithm (flat-sky hash +
	cartesian product + numpy scoring; kind 'port').  Single-threaded: the reference has no parallel path
	(/root/reference/TODO:5-14; joblib is only its disk cache, bypassed here so that every step does the work)."""
	tables, _ = make_workload(1, scale=scale)
	from oracle import refrun
	kind = 'reference' if refrun.package_root() is not None else 'port'
	if kind == 'reference':
		refrun.load_reference()
		run = lambda: len(refrun.run_reference([dict(t) for t in tables], RADIUS, COMPLETENESS))
	else:
		from oracle import nway_oracle as O
		run = lambda: len(O.nway_match(tables, RADIUS, COMPLETENESS, enumerator='refhash')['A'])
	times, rows = [], 0
	for k in range(warmup + steps):
		t0 = time.perf_counter()
		rows = run()
		dt = time.perf_counter() - t0
		if k >= warmup:
			times.append(dt)
	return rows, times, kind


def cpu_sample_text(kind, scale, rows, sec):
	what = ('the unmodified nwaylib.nway_match (nway 4.7.1, oracle/_ref), joblib cache bypassed' if kind == 'reference'
		else 'oracle port of the reference algorithm (flat-sky hash + product + numpy scoring)')
	return 'C3 at %.4g of the area (%d x %d sources, same densities): %d rows per step in %.2f s; %s; single-threaded like the reference' % (
		scale, round(N_PRIMARY * scale), round(N_SECONDARY * scale), rows, sec, what)


def run_reference(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	scale = args.ref_scale
	if scale is None:
		# bounded sample: size it so that the warm-up + K timed steps end within about three minutes on whatever host this
		# is (the algorithm is linear in the area at fixed densities; one small probe step measures this host's speed)
		probe = 0.004
		cpu_reference(probe)   # imports, first-call costs
		_, t, _ = cpu_reference(probe)
		per_unit = t[0] / probe
		scale = max(0.004, min(0.05, 150.0 / (per_unit * (args.steps + min(args.warmup, 1)))))
	rows, times, kind = cpu_reference(scale, steps=args.steps, warmup=min(args.warmup, 1))
	sec = sum(times) / len(times)
	value = rows / sec
	line = {
		'impl': 'reference', 'metric': 'candidate associations/sec', 'value': value, 'unit': 'associations/s',
		'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
		'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
		'config': {'workload': WORKLOAD, 'radius_arcsec': RADIUS, 'prior_completeness': COMPLETENESS,
			'sample': 'each step = the workload at %.4g of its area (same densities), see cpu_baseline.sample' % scale},
		'cpu_baseline': {'value': value, 'unit': 'associations/s', 'cores': 1, 'kind': kind, 'sample': cpu_sample_text(kind, scale, rows, sec)},
		'e2e': {'value': value, 'unit': 'associations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0,
	}
	print(json.dumps(line))


def run_b200(args):
	import torch
	import torch.distributed as dist
	import nway_b200
	from nway_b200 import _lib

	rank = int(os.environ.get('RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	local = int(os.environ.get('LOCAL_RANK', '0'))
	if not torch.cuda.is_available():
		raise SystemExit('bench.py: no CUDA device -- there is no CPU fallback for the product path')
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	affinity = bind_near_gpu(local)
	if world > 1:
		if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':   # its banner goes to stdout, which carries the ONE JSON line
			os.environ['NCCL_DEBUG'] = 'WARN'
		dist.init_process_group('nccl', device_id=dev)

	tables, n0 = make_workload(world, scale=args.scale)
	ctx = _lib.Context(local)
	# one explicit stream for the match kernels AND the collectives (torch's default stream has the handle 0, which the
	# library reads as "use your own stream"): everything of a step is ordered on it
	stream = torch.cuda.Stream(device=dev)
	torch.cuda.set_stream(stream)
	ctx.set_stream(stream.cuda_stream)

	# ---- inputs resident in HBM ------------------------------------------------------------------------
	dev_arrays = []
	for c, t in enumerate(tables):
		ra = torch.from_numpy(np.ascontiguousarray(t['ra'])).to(dev)
		dec = torch.from_numpy(np.ascontiguousarray(t['dec'])).to(dev)
		err = torch.from_numpy(np.ascontiguousarray(t['error'], dtype=np.float64)).to(dev)
		dev_arrays.append((ra, dec, err))
		ctx.set_catalogue_device(c, len(tables), len(t['ra']), ra.data_ptr(), dec.data_ptr(), err.data_ptr(), t['area'])
	tab = nway_b200._scalar_tables(tables, COMPLETENESS, nway_b200.NullOutputLogger())
	ctx.set_params(RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
	ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
	# the row set of the reference: on this flat-sky field nwaylib hashes on (ra, dec) cells (fastskymatch.py:94-101); the
	# device applies the same bucket predicate to every match (at |dec| <= 0.5 deg it removes nothing, tests/test_gpu_fullsize.py)
	ctx.set_compat(_lib.COMPAT_FLAT_HASH)
	ctx.set_primary_range(rank * n0, n0)
	# the one exchange of the sharded path: per-rank row counts (what is needed to place each shard in the global table),
	# gathered by NCCL straight from device memory.  It runs on a side stream, double-buffered, so that the next
	# match does not queue behind the collective's latency; torch.cuda.synchronize() at the end of the timed region
	# waits for the last one.
	counts = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(2)]
	mine = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(2)]
	side = torch.cuda.Stream(device=dev) if world > 1 else None
	copied = [torch.cuda.Event() for _ in range(2)]
	gathered = [torch.cuda.Event() for _ in range(2)]

	state = {'view': None, 'k': 0}

	def step():
		# nwb_match_async: the whole match is enqueued without a host round trip (the row count stays on the device), so
		# consecutive steps run back to back; match_wait() below collects the last one and checks its status words
		ctx.match_async(fuse_final=True)
		if world > 1:
			b = state['k'] & 1
			state['k'] += 1
			stream.wait_event(gathered[b])          # the collective that last used this buffer is done
			mine[b].copy_(state['view'])            # 8 bytes, before the next match overwrites the device word
			copied[b].record(stream)
			side.wait_event(copied[b])
			with torch.cuda.stream(side):
				dist.all_gather_into_tensor(counts[b], mine[b])
				gathered[b].record(side)

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	rows = ctx.match(fuse_final=True)   # the first match of a context chooses the grid geometry and sizes the buffers
	if world > 1:
		state['view'] = torch.as_tensor(_lib.DeviceView(ctx.nrows_device_ptr(), 1), device=dev)
	for _ in range(max(args.warmup, 3)):
		step()
	rows = ctx.match_wait()
	barrier()
	sampler = ClockSampler(local)
	if rank == 0:
		sampler.start()
	acc = {k: 0.0 for k in _lib.STAGE_NAMES}
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	barrier()
	e0.record(stream)
	nsampled = 0
	for it in range(args.steps):
		step()
		if it % 32 == 31 or it == args.steps - 1:   # per-stage event times of that step: needs the step collected (one host sync)
			rows = ctx.match_wait()
			for k, v in ctx.timings().items():
				acc[k] += v
			nsampled += 1
	e1.record(stream)
	barrier()
	ms_step = e0.elapsed_time(e1) / args.steps
	launches = ctx.launch_count() * args.steps
	stats = ctx.stats()

	t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
	r = torch.tensor([rows], dtype=torch.int64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		dist.all_reduce(r, op=dist.ReduceOp.SUM)
	ms_step_max = float(t.item())
	total_rows = int(r.item())
	value = total_rows / (ms_step_max * 1e-3)
	if world > 1:
		assert int(counts[(state['k'] - 1) & 1].sum().item()) == total_rows, 'the gathered row counts do not add up'

	# ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ------------------------------
	names = [tables[0]['name'], tables[1]['name']]
	host_in = []
	for c, tb in enumerate(tables):
		if c == 0:
			sl = slice(rank * n0, (rank + 1) * n0)
		else:
			sl = slice(None)
		arrs = [torch.from_numpy(np.ascontiguousarray(np.asarray(tb[k], dtype=np.float64)[sl])).pin_memory() for k in ('ra', 'dec', 'error')]
		host_in.append(arrs)
	colsel = [_lib.COL_IDX, _lib.COL_IDX + 1, _lib.COL_SEP, _lib.COL_SEPMAX, _lib.COL_NCAT, _lib.COL_LOGBF_UNCORR, _lib.COL_LOGBF,
		_lib.COL_DIST_POST, _lib.COL_P_SINGLE, _lib.COL_MATCH_FLAG, _lib.COL_P_ANY, _lib.COL_P_I]
	host_out = [torch.empty(rows + 1024, dtype=torch.float64).pin_memory() for _ in colsel]
	ctx2 = ctx
	ctx2.set_primary_range(0, n0)
	h2d = sum(a.numel() * 8 for arrs in host_in for a in arrs)
	n1_all = host_in[1][0].numel()
	d2h = rows * 8 * len(colsel)

	def e2e_step():
		for c, arrs in enumerate(host_in):
			n = arrs[0].numel()
			ctx2.check(ctx2.lib.nwb_set_catalogue(ctx2.h, c, 2, n, arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(),
				_lib.ERR_CIRCULAR, None, 0, float(tables[c]['area']), 0))
		ctx2.set_params(RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
		# the densities belong to the whole catalogue (world blocks of primaries): keep the tables of the resident run
		ctx2.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
		nr = ctx2.match(fuse_final=True)
		for sel, buf in zip(colsel, host_out):
			ctx2.check(ctx2.lib.nwb_fetch(ctx2.h, sel, buf.data_ptr()))
		ctx2.sync()
		return nr

	e2e_step()
	e2e_step()
	barrier()
	t0 = time.perf_counter()
	for _ in range(3):
		nr = e2e_step()
	barrier()
	single_call_ms = (time.perf_counter() - t0) / 3 * 1e3   # one call at a time: H2D, match, D2H strictly one after the other
	assert nr == rows

	# Throughput: two contexts in flight.  PCIe is full duplex, but within ONE step the table can only travel to the host
	# after the catalogues have travelled to the device; with a second context (own stream, own buffers) the next step's
	# H2D + match run while the previous step's table is still on its way back.  Every step still copies all of its
	# inputs from pinned host memory and all of its result columns back, in order, on its own stream.
	ctx.set_stream(None)   # back to the context's own non-blocking stream
	ctx_b = _lib.Context(local)
	ctx_b.set_compat(_lib.COMPAT_FLAT_HASH)
	lanes = [(ctx, host_out), (ctx_b, [torch.empty(rows + 1024, dtype=torch.float64).pin_memory() for _ in colsel])]
	# N > 1: the secondary catalogue is the same on every rank.  It crosses PCIe ONCE per step (rank 0's pinned copy) and
	# reaches the other GPUs by an NCCL broadcast over NVLink, instead of N uploads competing for the host's memory system;
	# every rank still uploads its own primaries and downloads its own table.
	lane_streams = [torch.cuda.Stream(device=dev) for _ in lanes] if world > 1 else None
	sec_dev = [torch.empty((3, n1_all), dtype=torch.float64, device=dev) for _ in lanes] if world > 1 else None
	if world > 1:
		for (c, _), ls in zip(lanes, lane_streams):
			c.set_stream(ls.cuda_stream)

	def e2e_begin(c, out, lane=0):
		c.sync()   # this lane's previous step is complete (its host buffers may be overwritten)
		for k, arrs in enumerate(host_in):
			if k == 1 and world > 1:
				with torch.cuda.stream(lane_streams[lane]):
					if rank == 0:
						for j in range(3):
							sec_dev[lane][j].copy_(arrs[j], non_blocking=True)
					dist.broadcast(sec_dev[lane], src=0)
				c.set_catalogue_device(1, 2, n1_all, sec_dev[lane][0].data_ptr(), sec_dev[lane][1].data_ptr(), sec_dev[lane][2].data_ptr(), float(tables[1]['area']))
				continue
			c.check(c.lib.nwb_set_catalogue(c.h, k, 2, arrs[0].numel(), arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(),
				_lib.ERR_CIRCULAR, None, 0, float(tables[k]['area']), 0))
		c.set_params(RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
		c.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
		c.set_primary_range(0, n0)
		c.match_async(fuse_final=True)
		nr = c.match_wait()
		for sel, buf in zip(colsel, out):
			c.check(c.lib.nwb_fetch(c.h, sel, buf.data_ptr()))   # asynchronous: collected by the lane's next sync()
		return nr

	e2e_steps = max(4, min(args.steps, 12))
	for k in range(4):
		e2e_begin(*lanes[k % 2], lane=k % 2)
	for c, _ in lanes:
		c.sync()
	barrier()
	t0 = time.perf_counter()
	for k in range(e2e_steps):
		nr = e2e_begin(*lanes[k % 2], lane=k % 2)
		assert nr == rows
	for c, _ in lanes:
		c.sync()
	barrier()
	e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
	if rank == 0:
		sampler.stop()
	t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	e2e_ms_max = float(t.item())
	e2e_value = total_rows / (e2e_ms_max * 1e-3)
	for _, out in lanes:   # both lanes must have delivered the same table
		for k in (1, 2, 10, 11):   # bit patterns (the index columns are int64 in these 8-byte buffers; -1 would read as NaN)
			assert torch.equal(out[k][:rows].view(torch.int64), host_out[k][:rows].view(torch.int64)), 'lanes disagree in column %d' % k
	ctx_b.close()
	# a cheap sanity check of what came back (the parity tests do the real checking)
	p_any = host_out[10][:rows].numpy()
	assert np.isfinite(p_any).all() and (p_any >= -1e-12).all() and (p_any <= 1 + 1e-12).all()

	# ---- (N > 1) the same step WITH the reassembly of the output table: match, all-gather of the row counts, one
	# unpadded all-gather-v of every rank's shard (12 columns) straight from the context's column allocation into
	# its place in the gathered table (nway_b200.parallel.allgather_table: one NCCL group over NVLink) ----------------
	table_gather = None
	if world > 1:
		from nway_b200 import parallel
		torch.cuda.set_stream(stream)
		ctx.set_stream(stream.cuda_stream)
		for c, tb in enumerate(tables):   # back to the resident catalogues (the e2e loop left this rank's host copies in the context)
			ra, dec, err = dev_arrays[c]
			ctx.set_catalogue_device(c, len(tables), len(tb['ra']), ra.data_ptr(), dec.data_ptr(), err.data_ptr(), tb['area'])
		ctx.set_params(RADIUS, tab['pc'], 0.5, _lib.UNRELATED_API)
		ctx.set_tables(tab['norm'], tab['log10e'], tab['prior'], tab['log10prior'], tab['sub_log10prior'])
		ctx.set_primary_range(rank * n0, n0)
		recv_bytes = (total_rows - rows) * 8 * len(colsel)

		def check_gathered(gtab, cnts):
			# the gathered table: every shard in rank order, primaries ascending, identical on all ranks
			assert gtab.shape[1] == sum(cnts) == total_rows
			prim = gtab[0]
			assert bool((prim[1:] >= prim[:-1]).all()) and int(prim[0]) == 0 and int(prim[-1]) == world * n0 - 1
			chk = torch.stack([gtab[k].sum() for k in range(gtab.shape[0])]).to(torch.float64)   # bit patterns of all columns
			mx, mn = chk.clone(), chk.clone()
			dist.all_reduce(mx, op=dist.ReduceOp.MAX)
			dist.all_reduce(mn, op=dist.ReduceOp.MIN)
			assert bool((mx == mn).all()), 'the ranks hold different gathered tables'
			return chk

		def time_gather(step):
			for _ in range(3):
				step()
			barrier()
			g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			gsteps = max(3, min(args.steps, 20))
			g0.record(stream)
			for _ in range(gsteps):
				gtab, cnts = step()
			g1.record(stream)
			barrier()
			tg = torch.tensor([g0.elapsed_time(g1) / gsteps], dtype=torch.float64, device=dev)
			dist.all_reduce(tg, op=dist.ReduceOp.MAX)
			gms = float(tg.item())
			return gms, gsteps, check_gathered(gtab, cnts)

		def line_of(gms, gsteps, how):
			return {'value': total_rows / (gms * 1e-3), 'unit': 'associations/s', 'ms_per_step': gms, 'steps': gsteps,
				'allgather_ms': gms - ms_step_max, 'received_bytes_per_gpu': recv_bytes,
				'receive_GBs_per_gpu': recv_bytes / (max(gms - ms_step_max, 1e-6) * 1e-3) / 1e9, 'how': how}

		# (a) the product path: the library's peer-memory gather (nwb_gather_*: every GPU stores its shard into every rank's
		# table over NVLink), SM stores and, for comparison, the copy engines
		variants = {}
		sums = {}
		for name, engine in (('peer_memory_sm', 0), ('peer_memory_copy_engines', 1)):
			tg_ = parallel.TableGather(None, local, stream=stream, engine=engine)
			try:
				tg_.setup(ctx, total_rows + 4096, len(colsel))   # raises on every rank or on none
			except RuntimeError as e:
				variants[name] = {'unavailable': str(e)[:300]}
				continue

			def peer_step():
				ctx.match_async(fuse_final=True)
				nr = ctx.match_wait()
				return tg_(ctx)

			gms, gsteps, sums[name] = time_gather(peer_step)
			variants[name] = line_of(gms, gsteps, 'every step: match, NCCL all-gather of the row counts, nwb_gather_push (%s: each GPU writes its 12 columns into their final place in all %d tables over NVLink peer memory), one NCCL barrier; all ranks end with the whole table in HBM'
				% ('one kernel of 16-byte stores' if engine == 0 else 'cudaMemcpyAsync on one stream per destination', world))
			tg_.close(ctx)
		# (b) the same through torch.distributed / NCCL (parallel.allgather_table: one packed message per peer + unpacking)
		gt = torch.empty((len(colsel), total_rows), dtype=torch.int64, device=dev)

		def nccl_step():
			ctx.match_async(fuse_final=True)
			nr = ctx.match_wait()
			cnts = parallel.exchange_counts(nr, None, dev)
			parallel.allgather_table(ctx.table_view(), cnts, out=gt)
			return gt, cnts

		gms, gsteps, sums['nccl'] = time_gather(nccl_step)
		variants['nccl_packed'] = line_of(gms, gsteps, 'every step: match, NCCL all-gather of the row counts, one exact-size message per peer in one ncclGroup, unpacked into the columns on arrival')
		del gt
		for name in sums:
			assert torch.equal(sums[name], sums['nccl']), 'the gather variants disagree: ' + name
		table_gather = dict(variants['peer_memory_sm'] if 'ms_per_step' in variants['peer_memory_sm'] else variants['nccl_packed'])
		table_gather['variants'] = {k: ({'ms_per_step': v['ms_per_step'], 'allgather_ms': v['allgather_ms'], 'receive_GBs_per_gpu': v['receive_GBs_per_gpu']} if 'ms_per_step' in v else v) for k, v in variants.items()}

	if rank != 0:
		if world > 1:
			dist.destroy_process_group()
		return

	# ---- roofline of the dominant kernel ---------------------------------------------------------------
	peak, peak_kind = hbm_peak()
	k_pairs_ms = acc['k_pairs'] / nsampled
	k_rows_ms = acc['k_rows'] / nsampled
	n1 = len(tables[1]['ra'])
	pairs = stats['pairs_kept']
	ncols = 12
	bytes_pairs = n1 * 16 + pairs * 16                 # (ra, dec) of every secondary read once + pair records written
	bytes_rows = rows * 8 * ncols + pairs * 12 + n0 * 24  # output columns + sorted lists read + primary (err, offsets)
	if k_rows_ms >= k_pairs_ms:
		kname, kms, kbytes = 'k_rows2<fused>', k_rows_ms, bytes_rows
	else:
		kname, kms, kbytes = 'k_pairs', k_pairs_ms, bytes_pairs
	achieved = kbytes / (kms * 1e-3) / 1e9
	b_alg = sum(len(tb['ra']) for tb in tables[1:]) * 24 + n0 * 24 + rows * 8 * ncols   # B_in + R * B_row
	roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
		'peak_kind': peak_kind, 'traffic': None, 'kernel_ms': kms, 'algorithmic_bytes': kbytes,
		'other_kernel': {'k_pairs_ms': k_pairs_ms, 'k_pairs_GBs': bytes_pairs / (k_pairs_ms * 1e-3) / 1e9,
			'k_rows_ms': k_rows_ms, 'k_rows_GBs': bytes_rows / (k_rows_ms * 1e-3) / 1e9},
		'pipeline': {'B_alg_bytes': b_alg, 'GBs': b_alg / (ms_step_max * 1e-3) / 1e9, 'frac': b_alg / (ms_step_max * 1e-3) / 1e9 / peak}}
	ncu_traffic = os.path.join(ROOT, 'profiles', 'traffic.json')
	if os.path.exists(ncu_traffic) and args.scale == 1.0:
		try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel, from the committed ncu capture --
			# only if that capture was taken from the kernel sources this library was built from
			cap = json.load(open(ncu_traffic))
			prof = cap['kernels'][kname.split('<')[0]]
			if cap.get('kernel_source_sha256') == kernel_source_hash():
				roofline['traffic'] = prof['dram_bytes_per_launch']
				if 'limiter' in prof:   # what the ncu capture says actually bounds the kernel (it is not HBM): informational
					roofline['limiter'] = prof['limiter']
			else:
				roofline['traffic_note'] = 'profiles/traffic.json was captured from other kernel sources (%s...): not reported' % str(cap.get('kernel_source_sha256'))[:12]
		except Exception:
			pass
	if world == 1 and args.scale == 1.0:
		try:
			# the memory-system skeleton of k_pairs on the same grid and catalogue: every global-memory access of the kernel,
			# none of its arithmetic (nwb_bench_skeleton) -- what the access pattern alone costs
			ctx.set_compat(_lib.COMPAT_FLAT_HASH)
			ctx.match(fuse_final=True)
			by_blocks = {}
			for b in (2, 3, 4, 5):   # resident blocks per SM: the access stream is fastest when fewer warps than fit are in flight
				ctx.match(fuse_final=True)
				by_blocks[b] = ctx.bench_skeleton(1, 5, b)
			sk = min(by_blocks.values())
			roofline['k_pairs_memory_skeleton'] = {'ms': sk, 'GBs': bytes_pairs / (sk * 1e-3) / 1e9, 'k_pairs_over_skeleton': k_pairs_ms / sk,
				'ms_by_resident_blocks_per_sm': by_blocks, 'what': 'the same global-memory accesses, queues and atomics as k_pairs without its arithmetic; best of 2..5 resident blocks per SM'}
		except Exception as e:
			roofline['k_pairs_memory_skeleton'] = {'error': str(e)[:200]}

	cpu = None
	if world == 1 and not args.no_cpu:
		crow, ctimes, ckind = cpu_reference(args.cpu_scale)
		cpu = {'value': crow / ctimes[0], 'unit': 'associations/s', 'cores': 1, 'kind': ckind,
			'sample': cpu_sample_text(ckind, args.cpu_scale, crow, ctimes[0])}

	line = {
		'metric': 'candidate associations/sec', 'value': value, 'unit': 'associations/s', 'n_gpus': world,
		'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step_max, 'higher_is_better': True,
		'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
		'config': {'workload': WORKLOAD if args.scale == 1.0 else WORKLOAD + ' -- at %.3g of the area (%d x %d)' % (args.scale, n0, n1),
			'rows_per_gpu': rows, 'pairs_per_gpu': pairs, 'radius_arcsec': RADIUS, 'prior_completeness': COMPLETENESS,
			'parallelism': 'primary rows sharded, %d rank(s); secondaries replicated; exchange = NCCL all-gather of row counts (side stream, overlaps the next match)' % world,
			'stepping': 'nwb_match_async back to back, nwb_match_wait every 32 steps and at the end', 'cpu_affinity': affinity,
			'l2': 'inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2; no explicit flush' % (h2d / 1e6, d2h / 1e6),
			'stage_ms': {k: acc[k] / nsampled for k in acc}, 'grid': stats,
			'compat': 'NWB_COMPAT_FLAT_HASH (the reference\'s flat-sky row set)'},
		'roofline': roofline,
		'cpu_baseline': cpu,
		'with_table_allgather': table_gather,
		'e2e': {'value': e2e_value, 'unit': 'associations/s', 'h2d_bytes_per_step': h2d if world == 1 else h2d + (world - 1) * (h2d - 3 * 8 * n1_all), 'd2h_bytes_per_step': d2h * world, 'ms_per_step': e2e_ms_max,
			'bytes_are': 'whole job (all ranks): host -> device and device -> host bytes per step' + ('' if world == 1 else '; the secondary catalogue crosses PCIe once (rank 0) and is broadcast over NVLink (%d bytes per step)' % (3 * 8 * n1_all)),
			'how': 'host pinned buffers through the C ABI, every step: H2D of all catalogue columns, match, D2H of all 12 result columns; two contexts in flight (the H2D + match of one step overlap the D2H of the previous one)',
			'single_call_ms': single_call_ms},
		'gpu_launches': launches,
		'clocks': sampler.summary(),
	}
	print(json.dumps(line))
	if world > 1:
		dist.destroy_process_group()


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=200)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
	ap.add_argument('--scale', type=float, default=1.0, help='fraction of the C3 area (same densities); 1.0 = the named workload')
	ap.add_argument('--cpu-scale', type=float, default=0.05, help='sample of the workload the CPU baseline is timed on')
	ap.add_argument('--ref-scale', type=float, default=None, help='sample per step of --impl reference (fraction of the C3 area; default: sized from --steps so that the run ends within about two minutes)')
	ap.add_argument('--no-cpu', action='store_true')
	args = ap.parse_args()
	if args.impl == 'reference':
		run_reference(args)
	else:
		run_b200(args)


if __name__ == '__main__':
	main()
