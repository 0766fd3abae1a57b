// Kernels of the nway match path (sm_100a, fp64 CUDA cores -- there is no dense contraction on this path).
//
// Pipeline (one shard of primaries per context):
//   k_prim_prep     per primary: lon, sin/cos(lat), search box           -> P_* arrays; with a known grid
//                   geometry also the count pass: cellcnt[cell] += 1 for every cell the box overlaps (K0)
//   k_cell_headers  per cell: count, start of its overflow segment (block scan, no device-wide scan) (K0)
//   k_prim_cells    count (grid geometry just chosen) / fill: the first three primaries of a cell go
//                   straight into its 32-byte record, later ones into the overflow segment        (K0)
//   k_pairs         stream a secondary catalogue once: cell -> box test -> per-warp candidate queue ->
//                   dense exact separations -> (secondary, sep) straight into the primary's slot (K1)
//   k_spill_*       the rare primaries with more matches than slots: overflow records -> segments (K1)
//   k_sort_lists    N >= 3: per primary rank-sort by secondary index into compact lists          (K1)
//   k_count_rows    N >= 3: secondary-secondary separations + number of valid tuples             (K2)
//   k_rows          per primary: enumerate tuples in lexicographic order, score, write columns;
//                   optionally fused with the group normalisation                                (K2[+K3])
//   k_correct_cli   nway.py:366-421                                                              (K2b)
//   k_final         bias lookup, p_single, per-primary log-sum-exp, p_any, p_i, match_flag       (K3)
#pragma once
#include "nwb_device.cuh"
#include "nwb_grid.cuh"
#include "nwb_rows.cuh"

namespace nwb {

// box margins: rb (deg) is the search radius inflated by 1e-9 relative + 1e-12, so that rounding in the
// box test can never reject a pair the exact formula would accept.
// COUNT: the grid geometry is already known (re-used from the previous match, verified afterwards on the device), so
// the primary is counted into its cells right here -- one kernel and one pass over the primaries less.
template <bool COUNT>
__global__ void k_prim_prep(int np, long long first, const double *__restrict__ ra, const double *__restrict__ dec,
	double rb, PrimArrays P, unsigned long long *__restrict__ red /* [6], zeroed */,
	Grid G, double rb_ins, double dra_eps, int *__restrict__ cellcnt, double tau_max, FlatHash flat,
	CellRec *cells, OverflowItem *__restrict__ worklist, int *__restrict__ worklist_n, long long worklist_cap)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	double v[6] = {1e300, -1e300, 1e300, -1e300, 1e300, -1e300};   // dec min/max, A lo/hi, B lo/hi
	if (i < np) {
		double r = ra[first + i], d = dec[first + i];
		double lat = deg2rad_ref(d);
		double sl, cl;
		sincos_ref(lat, &sl, &cl);
		PrimRec pr;
		pr.lon = deg2rad_ref(r); pr.slat = sl; pr.clat = cl;
		pr.ij = flat.err > 0.0 ? flat_hash_pack(r, d, flat) : 0ll;
		P.rec[i] = pr;
		double rn = wrap360(r);
		const double dra = search_box_dra(d, rb);
		P.ra_n[i] = rn;
		P.dec[i] = d;
		P.dra[i] = dra;
		{
			// cos(dec) as the fp32 flat pre-test of the overflow entries wants it (struct Entry): rounded down, 0 for primaries
			// too close to a pole for the flat metric -- decided here, once per primary, not in the (divergent) fill pass
			const double tau = (rb_ins / 180 * NWB_PI) * tan(fmin(fabs(d), 89.9999) / 180 * NWB_PI);
			P.clat[i] = (tau > tau_max || dra >= 180.0) ? 0.0 : (double) __double2float_rd(cl);
		}
		double rn_b = wrap360(rn + 180.0);
		v[0] = d; v[1] = d;
		v[2] = rn - dra; v[3] = rn + dra;
		v[4] = rn_b - dra; v[5] = rn_b + dra;
		if (COUNT) prim_register<REG_COUNT_INLINE>(G, i, d, rn, dra, cl, rb_ins, dra_eps, 0, 1, cellcnt, cells, nullptr, worklist, worklist_n, worklist_cap);
	}
	// block reduce, then one atomicMax per quantity on an order-preserving integer image of the double
	// (minima are stored negated), so no second kernel is needed
	__shared__ double sm[6][32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 6; k++) {
		double x = (k & 1) ? v[k] : -v[k];
		for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(NWB_FULL, x, o));
		if (lane == 0) sm[k][w] = x;
	}
	__syncthreads();
	if (w == 0) {
		int nw = blockDim.x >> 5;
#pragma unroll
		for (int k = 0; k < 6; k++) {
			double x = lane < nw ? sm[k][lane] : -1e300;
			for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(NWB_FULL, x, o));
			if (lane == 0) {
				unsigned long long u = (unsigned long long) __double_as_longlong(x);
				u ^= (u >> 63) ? ~0ull : 0x8000000000000000ull;
				atomicMax(red + k, u);
			}
		}
	}
}

// four threads per primary, one declination band each (see prim_register).  FILL returns without writing when the
// overflow segments do not fit the entry buffer; the host sees the total and retries.
#ifndef NWB_FILL_MINBLOCKS
#define NWB_FILL_MINBLOCKS 3
#endif
template <bool FILL>
__global__ void __launch_bounds__(256, NWB_FILL_MINBLOCKS)
k_prim_cells(int np, Grid G, PrimArrays P, double rb_ins, double dra_eps,
	int *__restrict__ cellcnt, CellRec *cells, Entry *__restrict__ entries, const int *__restrict__ etotal, long long entries_cap)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int i = t >> 2, bslot = t & 3;
	if (i >= np) return;
	if (FILL && (long long) etotal[0] > entries_cap) return;
	prim_register<FILL ? REG_FILL : REG_COUNT>(G, i, P.dec[i], P.ra_n[i], P.dra[i], FILL ? P.clat[i] : 0.0, rb_ins, dra_eps, bslot, 4, cellcnt, cells, entries);
}

// the entries beyond a cell's third, noted by k_prim_prep<COUNT> (prim_register<REG_COUNT_INLINE>): 16-byte fp32 entries
// into the overflow segments k_cell_headers handed out
__global__ void __launch_bounds__(256)
k_fill_overflow(Grid G, PrimArrays P, const OverflowItem *__restrict__ worklist, const int *__restrict__ worklist_n, long long worklist_cap,
	const CellRec *__restrict__ cells, Entry *__restrict__ entries, const int *__restrict__ etotal, long long entries_cap)
{
	if ((long long) etotal[0] > entries_cap) return;   // the host retries with a bigger buffer
	const long long n = min((long long) worklist_n[0], worklist_cap);
	for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) {
		const int4 it = __ldg(reinterpret_cast<const int4 *>(worklist + k));
		const int i = it.x;
		double x = P.ra_n[i] - G.ra_org_n;
		if (x < 0.0) x += 360.0;
		Entry en;
		en.x = (float) x;
		en.y = (float) (P.dec[i] - G.dec_lo);
		en.clat = (float) P.clat[i];
		en.p = i;
		const int est = (int) (cells[it.y].q[0] >> 32);
		*reinterpret_cast<int4 *>(entries + est + it.z) = *reinterpret_cast<const int4 *>(&en);
	}
}

// One header per cell: q[0] = count | (start of the cell's overflow segment - 3) << 32, so that entry k >= 3 of the cell
// is entries[start + k]; the segments are handed out block by block (block scan + one atomicAdd per block -- no
// device-wide prefix sum).  Also the occupancy bitmap (sparse primaries) and the totals the host checks:
// totals[0] = overflow entries (must fit the entry buffer), totals[1] = registrations.
__global__ void __launch_bounds__(256)
k_cell_headers(long long ncells, const int *__restrict__ cellcnt, CellRec *__restrict__ cells, unsigned *__restrict__ bits,
	int *__restrict__ totals /* zeroed */)
{
	__shared__ int wneed[8], wcnt[8];
	__shared__ int base;
	const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int cnt = c < ncells ? cellcnt[c] : 0;
	if (bits) {   // a warp covers 32 consecutive cells = one word of the bitmap
		unsigned word = __ballot_sync(NWB_FULL, cnt > 0);
		if (lane == 0 && c < ncells) bits[c >> 5] = word;
	}
	const int need = max(cnt - 3, 0);
	int incl = need;
	for (int o = 1; o < 32; o <<= 1) {
		int y = __shfl_up_sync(NWB_FULL, incl, o);
		if (lane >= o) incl += y;
	}
	const int csum = __reduce_add_sync(NWB_FULL, cnt);
	if (lane == 31) wneed[w] = incl;
	if (lane == 0) wcnt[w] = csum;
	__syncthreads();
	if (threadIdx.x == 0) {
		int s = 0, sc = 0;
		for (int k = 0; k < 8; k++) { int t = wneed[k]; wneed[k] = s; s += t; sc += wcnt[k]; }
		base = s ? atomicAdd(&totals[0], s) : 0;
		if (sc) atomicAdd(&totals[1], sc);
	}
	__syncthreads();
	if (c < ncells) {
		const int start = base + wneed[w] + incl - need - 3;
		cells[c].q[0] = (unsigned long long) (unsigned) cnt | ((unsigned long long) (unsigned) start << 32);
	}
}

// ---------------------------------------------------------------------------------------------------------
// K1: stream one secondary catalogue
// ---------------------------------------------------------------------------------------------------------
// One thread per secondary source, coalesced streaming loads of (ra, dec): 16 algorithmic bytes per source, read
// exactly once.  Phase A (cheap, divergent): cell lookup in the L2-resident grid and the box test against the
// primaries registered in that cell; survivors go into a per-warp shared-memory queue.  Phase B (expensive,
// dense): whenever the queue holds 32 candidates every lane evaluates one exact separation, so the fp64 pipe
// runs full warps instead of the 1-5 active lanes a per-thread loop would give.  A match takes the next slot of
// its primary (atomicAdd on the per-primary counter) and is written there directly -- no append buffer, no
// scatter pass.
#ifndef NWB_K1_WARPS
#define NWB_K1_WARPS 8
#endif
constexpr int K1_WARPS = NWB_K1_WARPS;
constexpr int K1_ICAP = 64;                      // work items: 31 left over + 32 new (tested before the next 32)
constexpr int K1_QCAP = 64;                      // candidates: 31 left over + 32 new (flushed before the next 32)
constexpr int K1_SBANDS = 1024;                  // bands cached in shared memory (16 KB)

// resident blocks per SM the register budget is set for: 3 (<= 80 registers; the kernel wants 76-80 with the reference's
// sin / cos and the flat-hash predicate in the exact stage).  At 4 (64 registers) ptxas spills loop-carried values of the
// streaming loop: measured 291-302 us against 228-236 us on the benchmark (profiles/r02_kpairs_variants.txt)
#ifndef NWB_K1_MINBLOCKS
#define NWB_K1_MINBLOCKS 3
#endif
// the other instantiation (occupancy bitmap: sparse primaries, e.g. an all-sky match): measured on C5 (3e8 sources, 97 %
// of which end at their bitmap bit) 3.55 ms at 3 blocks per SM, 6.1 ms at 4 (64 registers: the streaming loop's values
// spill) -- and 2.5 ms in round 1, whose exact stage needed fewer registers.  Long sparse streams therefore run as two
// kernels (k_filter + k_pairs over the survivors, nwb_api.cu); this instantiation handles the survivors and the
// short catalogues.
#ifndef NWB_K1_MINBLOCKS_SPARSE
#define NWB_K1_MINBLOCKS_SPARSE 3
#endif


struct K1Smem {   // per warp
	int2 item_es[K1_ICAP];      // (entry index, secondary index)
	double2 item_rd[K1_ICAP];   // (ra, dec) of the secondary
	int2 cand_sp[K1_QCAP];      // (secondary, primary) that passed the pre-test
	double2 cand_rd[K1_QCAP];   // (ra, dec) of the secondary
};

struct K1Args {
	PrimArrays P;
	double radius;
	Slot16 *base;
	int C;
	int *cnt;
	SpillRec *spill;
	unsigned long long spill_cap;
	unsigned long long *spill_count;
	FlatHash flat;     // NWB_COMPAT_FLAT_HASH (k_pairs<.., FLAT = true>): the reference's bucket size in degrees
	// shard mode (k_pairs<.., SCAT = true>, nwb_shard_*): this rank streams a slice of the catalogue against ALL primaries
	// and writes every match straight into the pair store of the rank that owns the primary -- peer memory over NVLink
	int s_base;                    // catalogue index of the first streamed source
	int x_block;                   // primaries per rank: owner = p / x_block, index there = p % x_block
	char *const *x_peers;          // [world] base of every rank's exchange buffer (device array; own rank: local)
	long long x_cnt_off, x_slot_off, x_spill_off, x_spillcnt_off;   // byte offsets of this catalogue's areas in the exchange buffer
	// sparse primaries, two-kernel stream (k_filter + k_pairs over the survivors): surv_mode 1 = take the sources from the
	// survivor list (do nothing if it overflowed), 2 = the direct stream as the fallback (do nothing unless it overflowed)
	int surv_mode;
	const int *surv;               // survivors: catalogue indices ...
	const double2 *surv_rd;        // ... and their (ra, dec): one segment of surv_segcap records per block of k_filter
	const int *surv_cnt;           // [surv_nseg] records in every segment, [surv_nseg] = 1 if a segment overflowed
	int surv_nseg, surv_segcap;
};

// exact fp64 separation for `count` (<= 32) queued candidates, one per lane; a match takes the next slot of its
// primary
#ifndef NWB_FLUSH_NOINLINE
#define NWB_FLUSH_NOINLINE 0
#endif
#if NWB_FLUSH_NOINLINE
#define NWB_FLUSH_ATTR __noinline__
#else
#define NWB_FLUSH_ATTR __forceinline__
#endif
template <bool FLAT, bool SKEL, bool SCAT>
__device__ NWB_FLUSH_ATTR void k1_flush(const K1Smem &M, int lo, int count, int lane, const K1Args &A)
{
	if (lane < count) {
		const int2 c = M.cand_sp[lo + lane];
		const double2 rd = M.cand_rd[lo + lane];
		const int s = c.x, p = c.y;
		bool same = true;
		if (FLAT) {
			// NWB_COMPAT_FLAT_HASH: did the reference's hash bring these two together at all (fastskymatch.py:125-132)?  Decided
			// first, from the primary's cell word alone -- one predicate lives through the arithmetic below
			const unsigned long long wp = (unsigned long long) __ldg(&A.P.rec[p].ij);
			const int di = flat_hash_cell_fast(rd.x, A.flat) - (int) (wp >> 32);
			const int dj = flat_hash_cell_fast(rd.y, A.flat) - (int) (unsigned) wp;
			same = (unsigned) (di + 1) <= 2u && (unsigned) (dj + 1) <= 2u;   // cells below 2^30: the differences cannot wrap
		}
		const Sector32 pr = ldg_sector(A.P.rec + p);   // one sector, one request
		double sep;
		bool keep;
		if (SKEL) {
			// the memory-system skeleton (nwb_bench_skeleton): every access of the real kernel, none of its arithmetic;
			// ~3 % of the candidates are dropped at random, the share the exact test rejects
			sep = __longlong_as_double((long long) pr.q[0]) + rd.y;
			keep = (((unsigned) s * 2654435761u) ^ (unsigned) p) % 32u != 0u;
		} else {
			double slat2, clat2;
			sincos_ref(deg2rad_ref(rd.y), &slat2, &clat2);
			double lon2 = deg2rad_ref(rd.x);
			sep = sep_arcsec_ref(__longlong_as_double(pr.q[0]), __longlong_as_double(pr.q[1]), __longlong_as_double(pr.q[2]),
				lon2, slat2, clat2);
			keep = sep < A.radius && same;
		}
		if (keep) {
			if (SCAT) {
				// the pair store of the primary's owner, in that rank's exchange buffer: one system-scope atomic and one
				// 16-byte store over NVLink (the own rank's share stays local)
				const int r = p / A.x_block, pl = p - r * A.x_block;
				char *pb = reinterpret_cast<char *>(__ldg(reinterpret_cast<const unsigned long long *>(A.x_peers + r)));
				const int slot = atomicAdd_system(reinterpret_cast<int *>(pb + A.x_cnt_off) + pl, 1);
				if (slot < A.C) {
					int4 v;
					v.x = s + A.s_base; v.y = 0; v.z = __double2loint(sep); v.w = __double2hiint(sep);
					*reinterpret_cast<int4 *>(reinterpret_cast<Slot16 *>(pb + A.x_slot_off) + (size_t) pl * A.C + slot) = v;
				} else {
					const unsigned long long pos = atomicAdd_system(reinterpret_cast<unsigned long long *>(pb + A.x_spillcnt_off), 1ull);
					if (pos < A.spill_cap) {
						SpillRec rec;
						rec.p = pl; rec.slot = slot; rec.s = s + A.s_base; rec.pad = 0; rec.sep = sep;
						reinterpret_cast<SpillRec *>(pb + A.x_spill_off)[pos] = rec;
					}
				}
			} else {
				int slot = atomicAdd(&A.cnt[p], 1);
				if (slot < A.C) {
					int4 v;
					v.x = s; v.y = 0; v.z = __double2loint(sep); v.w = __double2hiint(sep);
					*reinterpret_cast<int4 *>(A.base + (size_t) p * A.C + slot) = v;
				} else {
					unsigned long long pos = atomicAdd(A.spill_count, 1ull);
					if (pos < A.spill_cap) {
						SpillRec rec;
						rec.p = p; rec.slot = slot; rec.s = s; rec.pad = 0; rec.sep = sep;
						A.spill[pos] = rec;
					}
				}
			}
		}
	}
}

// queue the lanes with pass == true as candidates; run the exact stage when 32 are there
template <bool FLAT, bool SKEL, bool SCAT>
__device__ __forceinline__ void k1_enqueue(K1Smem &M, bool pass, int s, int p, double r, double d, int lane, int &qn,
	const K1Args &A)
{
	unsigned m = __ballot_sync(NWB_FULL, pass);
	if (m) {
		if (pass) {
			int q = qn + __popc(m & ((1u << lane) - 1));
			M.cand_sp[q] = make_int2(s, p);
			M.cand_rd[q] = make_double2(r, d);
		}
		qn += __popc(m);
		__syncwarp();
		if (qn >= 32) {
			qn -= 32;
			k1_flush<FLAT, SKEL, SCAT>(M, qn, 32, lane, A);
			__syncwarp();
		}
	}
}

// fp32 pre-test of `count` (<= 32) work items (entries beyond the third of a cell), one per lane; the survivors are
// appended to the candidate queue, which must hold fewer than 32 on entry
__device__ __forceinline__ void k1_items(K1Smem &M, int lo, int count, int lane, int &qn, const Grid &G,
	const Entry *__restrict__ entries)
{
	bool pass = false;
	int s = 0, p = 0;
	double r = 0, d = 0;
	if (lane < count) {
		const int2 es = M.item_es[lo + lane];
		const double2 rd = M.item_rd[lo + lane];
		const int4 ev = __ldg(reinterpret_cast<const int4 *>(entries + es.x));
		r = rd.x; d = rd.y;
		double x = wrap360(r) - G.ra_org_n;
		if (x < 0.0) x += 360.0;
		pass = k1_pretest(G, (float) x, (float) (d - G.dec_lo), __int_as_float(ev.x), __int_as_float(ev.y), __int_as_float(ev.z));
		s = es.y;
		p = ev.w;
	}
	const unsigned m = __ballot_sync(NWB_FULL, pass);
	if (pass) {
		const int q = qn + __popc(m & ((1u << lane) - 1));
		M.cand_sp[q] = make_int2(s, p);
		M.cand_rd[q] = make_double2(r, d);
	}
	qn += __popc(m);
}

// Sparse primaries (an all-sky match: a few per cent of the grid cells hold a primary), first of two kernels: the pure
// stream.  One thread per source: coordinates -> cell of a COARSE REGULAR bitmap -> one bit; the sources whose bit is set
// are appended to a survivor list, which k_pairs then works through (starting with the bitmap of the grid proper).  No queues, no exact stage: the kernel streams the catalogue
// (BASELINE.json configs[4]: 3e8 sources, 97 % of which end at their bit).
// The bit is what the stream pays for.  A gather from global memory costs one L1 wavefront per distinct sector -- 32
// per warp and look-up for sources in random order, at one wavefront per cycle and SM that alone is 1.07 ms for 3e8
// sources, an L1 hit or not.  So the bitmap lives in SHARED memory: one block of 1024 threads per SM holds a copy of all
// of it (<= 224 KB: 1.8 M cells -- for the all-sky case 9 arcmin wide, 6 % of them occupied by 1e5 primaries), and the
// look-up is a shared-memory load whose 32 random banks collide three or four deep.  Four sources per thread and
// iteration, each one's successor loaded into its registers while it is processed: 64 KB of coordinates in
// flight per SM, enough for the HBM latency-bandwidth product.  Every block appends its survivors (index, ra, dec) to
// a SEGMENT OF ITS OWN in global memory: the position comes from a shared-memory counter (one atomicAdd per warp and
// batch), so there is no global atomic, no staging buffer and no barrier in the loop; k_pairs walks all segments and
// skips their unused tails.  A segment that overflows (a far denser patch of sky than the average) raises a flag and
// the direct stream takes over.
constexpr int KF_U = 4;           // sources per thread and iteration (eight: spills at the 64 registers 1024 threads leave)
constexpr int KF_THREADS = 1024;
constexpr int KF_SMEM_WORDS = 56 * 1024;   // 224 KB of bitmap

// Cell functions of the regular bitmap: ROUND(coordinate * scale), by adding 1.5 * 2^52 and reading the low word -- one
// fixed-latency DADD where floor() is a conversion through the slow pipe; any monotone function serves, as long as the
// registration and the look-up share it.  inv_w2 and band2_scale are (nr2 - 1) / ra_span and (nb2 - 1) / nbands, so a
// coordinate inside the grid (0 <= x <= ra_span, 0 <= t < nbands) lands inside without clamping; the registration clamps.
__device__ __forceinline__ int kf_round(double v) { return __double2loint(v + 6755399441055744.0); }

__device__ __forceinline__ int kf_racell2(const Grid &G, double x)
{
	const int i = kf_round(fmin(fmax(x, 0.0), 360.0) * G.inv_w2);
	return i >= G.nr2 ? G.nr2 - 1 : (i < 0 ? 0 : i);
}

__device__ __forceinline__ int kf_band2(const Grid &G, double t /* k1_band_coord */)
{
	const int b = kf_round(fmin(fmax(t, 0.0), (double) G.nbands) * G.band2_scale);
	return b >= G.nb2 ? G.nb2 - 1 : (b < 0 ? 0 : b);
}

// The look-up in the block's copy of the regular bitmap.  (Following it with the grid's own bitmap for the few that pass --
// two dependent gathers that some lane of nearly every warp then waits for -- cost more than it saved: 2.49 ms instead of
// 1.89 ms for the stream of configs[4]; k_pairs makes that test anyway, with full warps of survivors.)
__device__ __forceinline__ bool kf_occupied(const Grid &G, const unsigned *__restrict__ sbits, double r, double d, double nbands_d)
{
	const double t = (d - G.dec_lo) * G.inv_h;
	double x = r - G.ra_org_n;   // ra in [0, 360), as nearly always: one subtraction; anything else takes the general route
	if (x < 0.0) x += 360.0;
	if (!(r >= 0.0 && r < 360.0)) x = k1_ra_coord(G, r);
	const bool inside = t >= 0.0 && t < nbands_d && (G.full_circle || x <= G.ra_span);
	const int cell2 = kf_round(t * G.band2_scale) * G.nr2 + kf_round(x * G.inv_w2);
	return inside && (sbits[cell2 >> 5] >> (cell2 & 31) & 1u) != 0u;
}

// Registration of the primaries in the regular bitmap: every cell the (slightly inflated) search box overlaps, strip by
// strip -- the same box, margins and wrap rules as prim_register, on the regular cells; strips and cells come from the
// same monotone functions the look-up uses, so a source inside the box finds a set bit.
__global__ void k_prim_bits2(int np, Grid G, PrimArrays P, double rb_ins, double dra_eps, unsigned *__restrict__ bits2)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= np) return;
	const double d = P.dec[i], rn = P.ra_n[i], di = P.dra[i] + dra_eps;
	const int b0 = kf_band2(G, k1_band_coord(G, d - rb_ins)), b1 = kf_band2(G, k1_band_coord(G, d + rb_ins));
	const int n = G.nr2;
	int i0, cnt;
	if (G.full_circle) {
		const double cellw = G.ra_span / n;
		if (2 * di + 2 * cellw >= 360.0) { i0 = 0; cnt = n; }
		else {
			i0 = kf_racell2(G, wrap360(rn - di - G.ra_org));
			const int i1 = kf_racell2(G, wrap360(rn + di - G.ra_org));
			cnt = (i1 - i0 + n) % n + 1;
		}
	} else {
		const double x0 = wrap360(rn - G.ra_org) - di, x1 = wrap360(rn - G.ra_org) + di;
		i0 = kf_racell2(G, fmax(x0, 0.0));
		cnt = kf_racell2(G, fmin(x1, G.ra_span)) - i0 + 1;
	}
	for (int b = b0; b <= b1; b++)
		for (int k = 0; k < cnt; k++) {
			const int cell = b * n + (i0 + k) % n;
			atomicOr(bits2 + (cell >> 5), 1u << (cell & 31));
		}
}

// one batch of the stream: KF_U sources per thread, each replaced in its registers by the source of the next batch as soon
// as it has been read; FULL = every index of this batch and the next is inside the catalogue
template <bool FULL>
__device__ __forceinline__ void kf_batch(const Grid &G, const unsigned *__restrict__ sbits, int n, int base, int next,
	const double *__restrict__ ra, const double *__restrict__ dec, double *rr, double *dd,
	double nbands_d, int *nblk, int *__restrict__ surv, double2 *__restrict__ surv_rd, int segcap)
{
#pragma unroll
	for (int u = 0; u < KF_U; u++) {
		const int i = base + u * KF_THREADS + (int) threadIdx.x;
		const int j = next + u * KF_THREADS + (int) threadIdx.x;
		const double r = rr[u], d = dd[u];
		if (FULL || j < n) { rr[u] = ra[j]; dd[u] = dec[j]; }
		if ((FULL || i < n) && kf_occupied(G, sbits, r, d, nbands_d)) {
			// a few per cent of the sources get here: one shared-memory atomic each (the compiler aggregates it per warp)
			const int pos = atomicAdd(nblk, 1);
			if (pos < segcap) {
				surv[pos] = i;
				surv_rd[pos] = make_double2(r, d);
			}
		}
	}
}

// n + one wave of sources < 2^31 (the host's guarantee for every stream kernel): 32-bit indices throughout
__global__ void __launch_bounds__(KF_THREADS, 1)
k_filter(int n, const double *__restrict__ ra, const double *__restrict__ dec, Grid G, int *__restrict__ surv,
	double2 *__restrict__ surv_rd, int *__restrict__ surv_cnt /* [gridDim.x + 1] */, int segcap)
{
	extern __shared__ unsigned kf_sbits[];   // nb2 * nr2 / 32 words
	__shared__ int nblk;
	{
		const int words = G.nb2 * (G.nr2 >> 5);
		const uint4 *src = reinterpret_cast<const uint4 *>(G.bits2);   // 16-byte aligned, padded to a multiple of four words
		uint4 *dst = reinterpret_cast<uint4 *>(kf_sbits);
		for (int w = threadIdx.x; w < (words + 3) / 4; w += KF_THREADS) dst[w] = __ldg(src + w);
	}
	if (threadIdx.x == 0) nblk = 0;
	__syncthreads();
	constexpr int chunk = KF_THREADS * KF_U;                   // sources per block and iteration
	const int stride = (int) gridDim.x * chunk;
	const double nbands_d = (double) G.nbands;
	surv += (long long) blockIdx.x * segcap;
	surv_rd += (long long) blockIdx.x * segcap;
	int base = (int) blockIdx.x * chunk;
	double rr[KF_U], dd[KF_U];
#pragma unroll
	for (int u = 0; u < KF_U; u++) {
		const int i = base + u * KF_THREADS + (int) threadIdx.x;
		rr[u] = 0; dd[u] = 0;
		if (i < n) { rr[u] = ra[i]; dd[u] = dec[i]; }
	}
	for (; base < n; base += stride) {
		const int next = base + stride;
		if (next + chunk <= n) kf_batch<true>(G, kf_sbits, n, base, next, ra, dec, rr, dd, nbands_d, &nblk, surv, surv_rd, segcap);
		else kf_batch<false>(G, kf_sbits, n, base, next, ra, dec, rr, dd, nbands_d, &nblk, surv, surv_rd, segcap);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		surv_cnt[blockIdx.x] = min(nblk, segcap);
		if (nblk > segcap) surv_cnt[gridDim.x] = 1;
	}
}

// One thread per secondary source, coalesced loads of (ra, dec): 16 algorithmic bytes per source, the next batch
// prefetched while the current one is processed.  The kernel lives on L2 -> SM sector traffic (every lookup is a random
// 32-byte sector), so the data is laid out to need few of them:
//   1. per source ONE sector: the cell record = number of primaries registered in the cell, and the first three of
//      them inline; they are pre-tested (fp32, flat metric) on the spot.  Further primaries of the cell become work
//      items in a per-warp shared-memory list and are pre-tested 32 at a time (one 16-byte entry each).
//   2. survivors (candidates) carry the source's (ra, dec) with them; whenever 32 are queued every lane
//      evaluates one exact fp64 separation in the reference's arithmetic against ONE sector of primary data.
//   3. a match takes the next slot of its primary (atomicAdd on the per-primary counter) and is written there
//      directly -- no append buffer, no scatter pass.
// The two dense stages (work items -> candidates, candidates -> matches) run at ONE place of the loop, after everything
// of the batch has been queued: nothing of the batch is live across them, and the code of the exact stage exists once.
// DENSE: the band table fits shared memory and there is no occupancy bitmap (the streaming configuration of the
// benchmark).  FLAT: NWB_COMPAT_FLAT_HASH is in force.  SKEL: the memory-system skeleton (nwb_bench_skeleton).
// SCAT: shard mode -- matches go to the owner of the primary over peer memory (K1Args::x_*).
template <bool DENSE, bool FLAT, bool SKEL, bool SCAT>
__global__ void __launch_bounds__(K1_WARPS * 32, DENSE ? NWB_K1_MINBLOCKS : NWB_K1_MINBLOCKS_SPARSE)
k_pairs(int n, const double *__restrict__ ra, const double *__restrict__ dec, Grid G,
	const int *__restrict__ etotal, const CellRec *__restrict__ cells, const Entry *__restrict__ entries,
	long long entries_cap, K1Args A)
{
	__shared__ K1Smem smem[K1_WARPS];
	__shared__ int4 sbands[K1_SBANDS];   // the band table, when it is small enough (one L2 round trip less)
	__shared__ float skx[K1_SBANDS];
	if ((long long) etotal[0] > entries_cap) return;   // the cell lists were not written: the host retries
	const bool bands_in_smem = DENSE || G.nbands <= K1_SBANDS;
	if (bands_in_smem) {
		for (int b = threadIdx.x; b < G.nbands; b += blockDim.x) {
			sbands[b] = __ldg(reinterpret_cast<const int4 *>(G.bands + b));
			skx[b] = __ldg(G.kx + b);
		}
		__syncthreads();
	}
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1;
	K1Smem &M = smem[threadIdx.x >> 5];
	int nit = 0, qn = 0;   // warp-uniform fill levels of the two lists; both below 32 at the top of the loop
	// 32-bit indices: the host guarantees n + (one wave of threads) < 2^31 (secondary indices are ints in the slots anyway)
	// sparse primaries: the sources may come from the survivor list of k_filter (the ~3 % whose grid cell holds a primary)
	const bool indirect = !DENSE && A.surv_mode == 1;
	if (!DENSE && A.surv_mode != 0) {
		const bool overflowed = A.surv_cnt[A.surv_nseg] != 0;
		if (indirect ? overflowed : !overflowed) return;
	}
	// Survivor list: the warps are dealt out to the segments (warp w -> segment w % nseg, every (w / nseg)-th batch of it)
	// and walk only the filled part of theirs -- from here on n, stride and j0 are this WARP's: its segment's fill level, the
	// step between its batches, its position in the segment
	int stride = gridDim.x * blockDim.x;
	int j0 = blockIdx.x * blockDim.x + threadIdx.x;   // position in the stream (or in the warp's segment of the survivor list)
	long long slot0 = 0;
	if (indirect) {
		const int gw = j0 >> 5, nw = stride >> 5;
		const int seg = gw % A.surv_nseg, sub = gw / A.surv_nseg;
		const int nsub = max(nw / A.surv_nseg, 1);   // the host launches at least one warp per segment
		n = sub < nsub ? __ldg(A.surv_cnt + seg) : 0;
		slot0 = (long long) seg * A.surv_segcap;
		j0 = sub * 32 + lane;
		stride = nsub * 32;
	}
	const int nround = (n + 31) / 32 * 32;
	const double nbands_d = (double) G.nbands;
	int i_nxt = j0;
	double r_nxt = 0, d_nxt = 0;
	bool live_nxt = j0 < n;
	if (j0 < n) {
		if (indirect) {
			i_nxt = __ldg(A.surv + slot0 + j0);
			const double2 rd = __ldg(A.surv_rd + slot0 + j0);
			r_nxt = rd.x; d_nxt = rd.y;
		} else {
			r_nxt = ra[j0]; d_nxt = dec[j0];
		}
	}
	for (; j0 < nround; j0 += stride) {
		const bool last = j0 + stride >= nround;   // this warp's final batch: the lists are drained completely
		const double r = r_nxt, d = d_nxt;
		const int i = indirect ? i_nxt : j0;        // the source's index in its catalogue
		const bool live = live_nxt;
		{
			const int j = j0 + stride;   // software prefetch of the next batch: hides the DRAM latency
			live_nxt = j < n;
			if (j < n) {
				if (indirect) {
					i_nxt = __ldg(A.surv + slot0 + j);
					const double2 rd = __ldg(A.surv_rd + slot0 + j);
					r_nxt = rd.x; d_nxt = rd.y;
				} else {
					r_nxt = ra[j]; d_nxt = dec[j];
				}
			}
		}
		int ecnt = 0, estart = 0;
		unsigned long long e0 = 0, e1 = 0, e2 = 0;
		float xr = 0.f, yr = 0.f, kx = 0.f;
		if (live) {
			const double t = k1_band_coord(G, d);
			if (t >= 0.0 && t < nbands_d) {
				const double x = k1_ra_coord(G, r);
				if (G.full_circle || x <= G.ra_span) {
					const int b = __double2int_rd(t);
					BandRec B;
					if (bands_in_smem) {
						const int4 v = sbands[b];
						B.base = v.x; B.nra = v.y; B.inv_w = __hiloint2double(v.w, v.z);
						kx = skx[b];
					} else {
						B = load_band(G, b);
						kx = __ldg(G.kx + b);
					}
					int ic;
					const double xcells = k1_ra_cell(B, x, ic);
					const int cell = B.base + ic;
					if (DENSE || !G.bits || (__ldg(G.bits + (cell >> 5)) >> (cell & 31) & 1u)) {
						const Sector32 cr = ldg_sector(cells + cell);   // one sector, one request
						ecnt = (int) (unsigned) cr.q[0];
						estart = (int) (cr.q[0] >> 32);
						e0 = cr.q[1]; e1 = cr.q[2]; e2 = cr.q[3];
					}
					xr = (float) (xcells - (double) ic);
					yr = (float) (t - (double) b);
				}
			}
		}
		// the inline entries: up to three fp32 pre-tests on the spot
		const int ninl = min(ecnt, 3);
		const bool any = __any_sync(NWB_FULL, ninl > 0);
		if (!any && !last) continue;   // sparse primaries: most batches end here
		int maxc = 0;
		if (any) {
			// the three pre-tests first (independent: they overlap), then ONE update of the candidate queue
			const bool p0 = ninl > 0 && k1_pretest_packed(G, xr, yr, kx, (unsigned) e0);
			const bool p1 = ninl > 1 && k1_pretest_packed(G, xr, yr, kx, (unsigned) e1);
			const bool p2 = ninl > 2 && k1_pretest_packed(G, xr, yr, kx, (unsigned) e2);
			const unsigned m0 = __ballot_sync(NWB_FULL, p0), m1 = __ballot_sync(NWB_FULL, p1), m2 = __ballot_sync(NWB_FULL, p2);
			const int n0 = __popc(m0), n01 = n0 + __popc(m1), n012 = n01 + __popc(m2);
			if (qn + n012 <= K1_QCAP) {
				if (p0) { const int q = qn + __popc(m0 & lt); M.cand_sp[q] = make_int2(i, (int) (e0 >> 32)); M.cand_rd[q] = make_double2(r, d); }
				if (p1) { const int q = qn + n0 + __popc(m1 & lt); M.cand_sp[q] = make_int2(i, (int) (e1 >> 32)); M.cand_rd[q] = make_double2(r, d); }
				if (p2) { const int q = qn + n01 + __popc(m2 & lt); M.cand_sp[q] = make_int2(i, (int) (e2 >> 32)); M.cand_rd[q] = make_double2(r, d); }
				qn += n012;
			} else {   // more than the queue holds at once (rare): one entry rank at a time
#pragma unroll 1
				for (int k = 0; k < 3; k++) {
					const bool pk = k == 0 ? p0 : (k == 1 ? p1 : p2);
					const unsigned long long ek = k == 0 ? e0 : (k == 1 ? e1 : e2);
					k1_enqueue<FLAT, SKEL, SCAT>(M, pk, i, (int) (ek >> 32), r, d, lane, qn, A);
				}
			}
			maxc = __reduce_max_sync(NWB_FULL, ecnt);
		}
		// crowded cells (> 3 primaries): the entries beyond the third become work items; then the two dense stages
		int k = 3;
		for (;;) {
			while (k < maxc && nit < 32) {
				const bool has = k < ecnt;
				const unsigned m = __ballot_sync(NWB_FULL, has);
				if (has) {
					const int q = nit + __popc(m & lt);
					M.item_es[q] = make_int2(estart + k, i);
					M.item_rd[q] = make_double2(r, d);
				}
				nit += __popc(m);
				k++;
			}
			const bool done = k >= maxc, fin = last && done;
			__syncwarp();
			for (;;) {
				if (qn >= 32 || (fin && nit == 0 && qn > 0)) {   // candidates -> matches, 32 at a time (fewer only when draining)
					const int take = min(qn, 32);
					qn -= take;
					k1_flush<FLAT, SKEL, SCAT>(M, qn, take, lane, A);
					__syncwarp();
				} else if (nit >= 32 || (fin && nit > 0)) {      // work items -> candidates (fewer than 32 are queued here)
					const int take = min(nit, 32);
					nit -= take;
					k1_items(M, nit, take, lane, qn, G, entries);
					__syncwarp();
				} else {
					break;
				}
			}
			if (done) break;
		}
	}
}

// the scalars the host needs after K1, gathered for one small copy: [0] cell entries, [c] spill records of
// catalogue c, [8] total rows (N == 2)
struct BoundsKey { unsigned long long k[6]; };

__global__ void k_collect_status(int ncat, const int *__restrict__ entries_total,
	const unsigned long long *__restrict__ spill_count, const long long *__restrict__ total_rows,
	const unsigned long long *__restrict__ bounds, BoundsKey expected, long long *__restrict__ out,
	long long *__restrict__ host /* the same words in mapped pinned host memory: no copy-engine transfer to wait for */)
{
	int t = threadIdx.x;
	long long v = 0;
	bool mine = false;
	if (t == 0) { v = entries_total[0]; mine = true; }     // overflow entries of the cell lists
	if (t == 10) { v = entries_total[1]; mine = true; }    // registrations (primary, cell)
	if (t >= 1 && t < ncat) { v = (long long) spill_count[t]; mine = true; }
	if (t == 8) { v = total_rows ? *total_rows : 0; mine = true; }
	if (t == 9) {   // [9] != 0: the primaries' bounding box is not the one the (re-used) grid geometry was built for
		for (int k = 0; k < 6; k++) v |= (long long) (bounds[k] != expected.k[k]);
		mine = true;
	}
	if (mine) { out[t] = v; host[t] = v; }
}

// Two catalogues, a few thousand primaries: row offsets AND the status words in ONE single-block launch.
// row_off[p] = sum over q < p of (cnt[q] + 1) -- a primary's matches plus its no-counterpart row -- row_off[np] = R;
// then the words of k_collect_status.  Replaces cub's two scan kernels + k_collect_status (three dependent launches) for
// the small matches whose time is launch latency (the COSMOS configurations: 1797 primaries).  Measured on the
// benchmark's 1e5 primaries one block takes 78 us against cub's 21: each 4096-element tile is a dependent
// load -> scan -> store round trip of ~3 us on one SM, so larger matches keep the device-wide scan.
constexpr int RO_THREADS = 1024;

__global__ void __launch_bounds__(RO_THREADS)
k_rowoff_status(int np, const int *__restrict__ cnt, long long *__restrict__ row_off,
	int ncat, const int *__restrict__ entries_total, const unsigned long long *__restrict__ spill_count,
	const unsigned long long *__restrict__ bounds, BoundsKey expected, long long *__restrict__ out, long long *__restrict__ host)
{
	__shared__ long long wsum[32];
	__shared__ long long carry_s;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	if (t == 0) carry_s = 0;
	__syncthreads();
	for (int base = 0; base < np; base += RO_THREADS * 4) {
		const int i = base + 4 * t;
		int4 c = make_int4(0, 0, 0, 0);
		if (i + 3 < np) c = __ldg(reinterpret_cast<const int4 *>(cnt + i));   // cnt is 16-byte aligned (a multiple of 4 ints into its block)
		else {
			if (i < np) c.x = cnt[i];
			if (i + 1 < np) c.y = cnt[i + 1];
			if (i + 2 < np) c.z = cnt[i + 2];
		}
		const int v0 = i < np ? c.x + 1 : 0, v1 = i + 1 < np ? c.y + 1 : 0, v2 = i + 2 < np ? c.z + 1 : 0, v3 = i + 3 < np ? c.w + 1 : 0;
		const long long tot = (long long) v0 + v1 + v2 + v3;
		long long incl = tot;
		for (int o = 1; o < 32; o <<= 1) {
			const long long y = __shfl_up_sync(NWB_FULL, incl, o);
			if (lane >= o) incl += y;
		}
		if (lane == 31) wsum[w] = incl;
		__syncthreads();
		if (w == 0) {
			long long x = wsum[lane];
			for (int o = 1; o < 32; o <<= 1) {
				const long long y = __shfl_up_sync(NWB_FULL, x, o);
				if (lane >= o) x += y;
			}
			wsum[lane] = x;   // inclusive over warps
		}
		__syncthreads();
		const long long carry = carry_s;
		const long long excl = carry + (w > 0 ? wsum[w - 1] : 0) + incl - tot;
		if (i < np) row_off[i] = excl;
		if (i + 1 < np) row_off[i + 1] = excl + v0;
		if (i + 2 < np) row_off[i + 2] = excl + v0 + v1;
		if (i + 3 < np) row_off[i + 3] = excl + v0 + v1 + v2;
		__syncthreads();
		if (t == 0) carry_s = carry + wsum[31];
		__syncthreads();
	}
	const long long total = carry_s;
	if (t == 0) row_off[np] = total;
	long long v = 0;
	bool mine = false;
	if (t == 0) { v = entries_total[0]; mine = true; }     // overflow entries of the cell lists
	if (t == 10) { v = entries_total[1]; mine = true; }    // registrations (primary, cell)
	if (t >= 1 && t < ncat) { v = (long long) spill_count[t]; mine = true; }
	if (t == 8) { v = total; mine = true; }
	if (t == 9) {   // [9] != 0: the primaries' bounding box is not the one the (re-used) grid geometry was built for
		for (int k = 0; k < 6; k++) v |= (long long) (bounds[k] != expected.k[k]);
		mine = true;
	}
	if (mine) { out[t] = v; host[t] = v; }
}

// N >= 3 / elliptical, speculative pipeline (no host round trip between the stages): after each prefix sum one thread
// checks that what the sum says fits the buffers the previous match left -- list entries per catalogue, the secondary-
// secondary separation scratch, the output table -- and that the launch decisions taken from the previous match still
// cover this one (warp-per-primary kernels present where some primary needs them); the next stage's kernels read the
// word and do nothing if it is 0.  The totals also go to mapped host memory: the host reads them after its single
// synchronisation at the end and, if a gate stayed shut, redoes the match stage by stage.
struct GateArgs {
	int level, ncat;
	const long long *seg_total[MAXC];   // level 1: list entries of catalogue c (segment offsets [np])
	long long cap_list[MAXC];
	const int *maxcnt;                  // [MAXC]: largest match count of a primary
	int big_sort[MAXC];                 // the warp-per-primary sort is launched for catalogue c
	const long long *mat_total;         // level 2
	long long cap_mat;
	const unsigned long long *max_tuples;
	int any_big;                        // the warp-per-primary count / row kernels are launched
	const long long *rows_total;        // level 3
	long long cap_rows;
	int *gates;                         // [4]
	long long *host;                    // mapped: [16 + c] list entries, [40 ..] maxcnt words, [24] mat, [26] tuples, [25] rows, [60 + level] gate
};

__global__ void k_spec_gate(GateArgs A, int small_n, int small_t)
{
	if (threadIdx.x != 0) return;
	bool ok = true;
	if (A.level == 1) {
		for (int c = 1; c < A.ncat; c++) {
			const long long n = *A.seg_total[c];
			A.host[16 + c] = n;
			ok = ok && n <= A.cap_list[c] && (A.maxcnt[c] <= small_n || A.big_sort[c]);
		}
		for (int k = 0; k < 4; k++) A.host[40 + k] = reinterpret_cast<const long long *>(A.maxcnt)[k];
	} else if (A.level == 2) {
		const long long m = *A.mat_total;
		const unsigned long long t = *A.max_tuples;
		A.host[24] = m;
		A.host[26] = (long long) t;
		ok = A.gates[1] != 0 && m <= A.cap_mat && (t <= (unsigned long long) small_t || A.any_big);
	} else {
		const long long r = *A.rows_total;
		A.host[25] = r;
		ok = A.gates[2] != 0 && r <= A.cap_rows;
	}
	A.gates[A.level] = ok ? 1 : 0;
	A.host[60 + A.level] = ok ? 1 : 0;
}

// a few 8-byte words from device memory into mapped pinned host memory.  Small read-backs go this way rather than through
// cudaMemcpyAsync: a copy would queue on the device-to-host copy engine behind whatever bulk transfer another context
// has in flight there (two contexts overlapping their H2D / D2H), a store from a kernel does not.
__global__ void k_words_to_host(const long long *__restrict__ src, long long *__restrict__ dst, int n)
{
	if ((int) threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

// zero-fill (16-byte units) plus a few extra words elsewhere -- one launch instead of two memsets
__global__ void k_zero(int4 *__restrict__ a, long long n16, unsigned long long *__restrict__ b, int nb)
{
	const long long stride = (long long) gridDim.x * blockDim.x;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) a[i] = make_int4(0, 0, 0, 0);
	if (blockIdx.x == 0 && (int) threadIdx.x < nb) b[threadIdx.x] = 0ull;
}

// overflow handling (only launched when some primary had more matches than slots)
__global__ void k_spill_sizes(int np, const int *__restrict__ cnt, int C, int *__restrict__ sizes)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < np) sizes[p] = max(cnt[p] - C, 0);
}

__global__ void k_spill_scatter(long long n, const SpillRec *__restrict__ recs, int C,
	const long long *__restrict__ spill_off, Slot16 *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	SpillRec r = recs[i];
	Slot16 v;
	v.s = r.s; v.pad = 0; v.sep = r.sep;
	out[spill_off[r.p] + (r.slot - C)] = v;
}

// N >= 3: one warp per primary rank-sorts its matches by secondary index (the reference's sorted() of every
// bucket list, fastskymatch.py:181) into compact lists, and stores lon, sin/cos(lat) of each secondary for the
// secondary-secondary separations.
__global__ void k_sort_lists(int np, PairStore S, const long long *__restrict__ seg_off, int *__restrict__ L_s,
	double *__restrict__ L_sep, const double *__restrict__ ra, const double *__restrict__ dec,
	double *__restrict__ L_lon, double *__restrict__ L_slat, double *__restrict__ L_clat, int small_n,
	long long *__restrict__ L_ij, FlatHash flat, const int *gate = nullptr)
{
	if (!gate_open(gate)) return;
	int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int p = warp; p < np; p += nwarps) {
		long long lo = seg_off[p];
		int n = S.cnt[p];
		if (n <= small_n) continue;   // k_sort_lists_small's
		for (int e = lane; e < n; e += 32) {
			Slot16 me = store_get(S, p, e);
			int rank = 0;
			for (int f = 0; f < n; f++) rank += store_get(S, p, f).s < me.s;
			L_s[lo + rank] = me.s;
			L_sep[lo + rank] = me.sep;
			double sl, cl;
			sincos_ref(deg2rad_ref(dec[me.s]), &sl, &cl);
			L_lon[lo + rank] = deg2rad_ref(ra[me.s]);
			L_slat[lo + rank] = sl;
			L_clat[lo + rank] = cl;
			if (flat.err > 0.0) L_ij[lo + rank] = flat_hash_pack(ra[me.s], dec[me.s], flat);
		}
	}
}

// the same for primaries with at most SMALL_N matches, one THREAD per primary (sparse all-sky matches: most primaries
// have 0 or 1): insertion sort in registers
constexpr int SMALL_N = 4;

__global__ void k_sort_lists_small(int np, PairStore S, const long long *__restrict__ seg_off, int *__restrict__ L_s,
	double *__restrict__ L_sep, const double *__restrict__ ra, const double *__restrict__ dec,
	double *__restrict__ L_lon, double *__restrict__ L_slat, double *__restrict__ L_clat,
	long long *__restrict__ L_ij, FlatHash flat, const int *gate = nullptr)
{
	if (!gate_open(gate)) return;
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= np) return;
	int n = S.cnt[p];
	if (n == 0 || n > SMALL_N) return;
	long long lo = seg_off[p];
	Slot16 x[SMALL_N];
#pragma unroll
	for (int e = 0; e < SMALL_N; e++)
		if (e < n) x[e] = store_get(S, p, e);
#pragma unroll
	for (int a = 1; a < SMALL_N; a++) {
#pragma unroll
		for (int b = a; b >= 1; b--) {
			if (b < n && a < n && x[b].s < x[b - 1].s) { Slot16 t = x[b]; x[b] = x[b - 1]; x[b - 1] = t; }
		}
	}
#pragma unroll
	for (int e = 0; e < SMALL_N; e++) {
		if (e >= n) break;
		L_s[lo + e] = x[e].s;
		L_sep[lo + e] = x[e].sep;
		double sl, cl;
		sincos_ref(deg2rad_ref(dec[x[e].s]), &sl, &cl);
		L_lon[lo + e] = deg2rad_ref(ra[x[e].s]);
		L_slat[lo + e] = sl;
		L_clat[lo + e] = cl;
		if (flat.err > 0.0) L_ij[lo + e] = flat_hash_pack(ra[x[e].s], dec[x[e].s], flat);
	}
}

template <int NC>
__device__ __forceinline__ long long mat_block_offset(const int *nl, int c, int d)
{
	// blocks in order (1,2),(1,3),...,(2,3),...; nl[k] = list length of catalogue k (k >= 1)
	long long off = 0;
	for (int a = 1; a < NC; a++)
		for (int b = a + 1; b < NC; b++) {
			if (a == c && b == d) return off;
			off += (long long) nl[a] * nl[b];
		}
	return off;
}

// sizes of the secondary-secondary separation scratch per primary
// (+ the largest number of candidate tuples of any primary, so that the host can skip the warp-per-primary kernels
// when every primary is handled by the one-thread-per-primary ones)
__global__ void k_mat_sizes(int np, int ncat, Lists L, long long *__restrict__ sizes, unsigned long long *__restrict__ max_tuples,
	const int *gate = nullptr)
{
	if (!gate_open(gate)) return;
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long ntup = 0;
	if (p < np) {
		long long tot = 0;
		ntup = 1;
		for (int a = 1; a < ncat; a++) {
			unsigned long long na = (unsigned long long) (L.off[a][p + 1] - L.off[a][p]);
			ntup = ntup > (1ull << 40) ? ntup : ntup * (na + 1);
			for (int b = a + 1; b < ncat; b++)
				tot += (long long) na * (L.off[b][p + 1] - L.off[b][p]);
		}
		sizes[p] = tot;
	}
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long y = __shfl_xor_sync(NWB_FULL, ntup, o);
		ntup = y > ntup ? y : ntup;
	}
	if ((threadIdx.x & 31) == 0 && ntup > 1) atomicMax(max_tuples, ntup);
}

// largest element of an int array (match counts per primary)
__global__ void k_max_int(long long n, const int *__restrict__ x, int *__restrict__ out /* zeroed */)
{
	int m = 0;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) m = max(m, x[i]);
	m = __reduce_max_sync(NWB_FULL, m);
	if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// N >= 3: one warp per primary computes every secondary-secondary separation once and counts the tuples
// that survive the pairwise radius filter (__init__.py:166,180).
template <int NC>
__global__ void k_count_rows(RowParams R, long long *__restrict__ rows)
{
	if (!gate_open(R.gate)) return;
	int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int p = warp; p < R.np; p += nwarps) {
		int nl[NC];
		long long lo[NC];
		long long T = 1;
		nl[0] = 0; lo[0] = 0;
#pragma unroll
		for (int c = 1; c < NC; c++) {
			lo[c] = R.L.off[c][p];
			nl[c] = (int) (R.L.off[c][p + 1] - lo[c]);
			T *= nl[c] + 1;
		}
		if (T <= R.small_t) continue;   // k_count_rows_small's
		const long long wp = R.flat.err > 0.0 ? flat_hash_pack(R.ra[0][R.first + p], R.dec[0][R.first + p], R.flat) : 0ll;
		double *mat = R.mat + R.mat_off[p];
		long long boff = 0;
#pragma unroll
		for (int a = 1; a < NC; a++) {
#pragma unroll
			for (int b = a + 1; b < NC; b++) {
				int na = nl[a], nb = nl[b];
				for (int e = lane; e < na * nb; e += 32) {
					int j = e / nb, k = e - j * nb;
					long long ja = lo[a] + j, kb = lo[b] + k;
					mat[boff + e] = sep_arcsec_ref(R.L.lon[a][ja], R.L.slat[a][ja], R.L.clat[a][ja],
						R.L.lon[b][kb], R.L.slat[b][kb], R.L.clat[b][kb]);
				}
				boff += (long long) na * nb;
			}
		}
		__syncwarp();
		long long count = 0;
		for (long long t = lane; t < T; t += 32) {
			int dg[NC];
			long long rem = t;
#pragma unroll
			for (int c = NC - 1; c >= 1; c--) {
				dg[c] = (int) (rem % (nl[c] + 1));
				rem /= (nl[c] + 1);
			}
			bool ok = true;
			long long bo = 0;
#pragma unroll
			for (int a = 1; a < NC; a++) {
#pragma unroll
				for (int b = a + 1; b < NC; b++) {
					if (dg[a] > 0 && dg[b] > 0) {
						double s = mat[bo + (long long) (dg[a] - 1) * nl[b] + (dg[b] - 1)];
						ok = ok && (s < R.pair_radius[pair_index(a, b, NC)]);
					}
					bo += (long long) nl[a] * nl[b];
				}
			}
			if (R.flat.err > 0.0) ok = ok && tuple_in_one_bucket<NC>(R, wp, dg, lo);
			count += ok;
		}
		count = warp_sum_ll(count);
		if (lane == 0) rows[p] = count;
	}
}

// the same for primaries with at most R.small_t candidate tuples, one THREAD per primary
template <int NC>
__global__ void k_count_rows_small(RowParams R, long long *__restrict__ rows)
{
	if (!gate_open(R.gate)) return;
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= R.np) return;
	int nl[NC];
	long long lo[NC];
	int T = 1;
	nl[0] = 0; lo[0] = 0;
#pragma unroll
	for (int c = 1; c < NC; c++) {
		lo[c] = R.L.off[c][p];
		nl[c] = (int) (R.L.off[c][p + 1] - lo[c]);
		T = T <= 8 ? T * (nl[c] + 1) : T;
	}
	if (T > R.small_t) return;
	if (T == 1) { rows[p] = 1; return; }
	const long long wp = R.flat.err > 0.0 ? flat_hash_pack(R.ra[0][R.first + p], R.dec[0][R.first + p], R.flat) : 0ll;
	double *mat = R.mat + R.mat_off[p];
	long long boff = 0;
#pragma unroll
	for (int a = 1; a < NC; a++) {
#pragma unroll
		for (int b = a + 1; b < NC; b++) {
			int na = nl[a], nb = nl[b];
			for (int e = 0; e < na * nb; e++) {
				int j = e / nb, k = e - j * nb;
				long long ja = lo[a] + j, kb = lo[b] + k;
				mat[boff + e] = sep_arcsec_ref(R.L.lon[a][ja], R.L.slat[a][ja], R.L.clat[a][ja],
					R.L.lon[b][kb], R.L.slat[b][kb], R.L.clat[b][kb]);
			}
			boff += (long long) na * nb;
		}
	}
	int count = 0;
	for (int t = 0; t < T; t++) {
		int dg[NC];
		int rem = t;
#pragma unroll
		for (int c = NC - 1; c >= 1; c--) {
			dg[c] = rem % (nl[c] + 1);
			rem /= (nl[c] + 1);
		}
		bool ok = true;
		long long bo = 0;
#pragma unroll
		for (int a = 1; a < NC; a++) {
#pragma unroll
			for (int b = a + 1; b < NC; b++) {
				if (dg[a] > 0 && dg[b] > 0) {
					double s = mat[bo + (long long) (dg[a] - 1) * nl[b] + (dg[b] - 1)];
					ok = ok && (s < R.pair_radius[pair_index(a, b, NC)]);
				}
				bo += (long long) nl[a] * nl[b];
			}
		}
		if (R.flat.err > 0.0) ok = ok && tuple_in_one_bucket<NC>(R, wp, dg, lo);
		count += ok;
	}
	rows[p] = count;
}

// N == 2: rows per primary = matches + 1
// ---------------------------------------------------------------------------------------------------------
// per-row finalisation pieces shared by k_rows<FUSE> and k_final
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// elliptical positional errors (CLI only in the reference): bayesdistance.py:164-240, fastskymatch.py:50-74
// ---------------------------------------------------------------------------------------------------------
// For the catalogues in `present` (bit c) with sources sidx[c]: the circularised errors (-> sig) and, for every
// present pair a < b, the separation rescaled by the ratio of circular to directional error (-> sep), ready for
// log_bf_ref.  bayesdistance.py:207-240; offsets measured in the frame of the later catalogue's source as the CLI
// does (fastskymatch.py:299-312 with nway.py's make_separation_table_matrix).
__device__ __forceinline__ void ell_prepare(const RowParams &R, unsigned present, const long long *sidx, double *sig, double *sep)
{
	const int nc = R.ncat;
	double sx[MAXC], sy[MAXC], rho[MAXC], ra[MAXC], de[MAXC];
	for (int c = 0; c < nc; c++)
		if (present >> c & 1u) {
			long long i = sidx[c];
			sx[c] = R.err[c][i]; sy[c] = R.err[c][R.n[c] + i]; rho[c] = R.err[c][2 * R.n[c] + i];
			ra[c] = R.ra[c][i]; de[c] = R.dec[c][i];
			sig[c] = sqrt((sx[c] * sx[c] + sy[c] * sy[c]) / 2);
		}
	for (int a = 0; a < nc; a++)
		for (int b = a + 1; b < nc; b++)
			if ((present >> a & 1u) && (present >> b & 1u)) {
				double vx, vy;
				offsets_ref(ra[b], de[b], ra[a], de[a], vx, vy);
				if (R.sep_f32) { vx = (double) (float) vx; vy = (double) (float) vy; }   // the CLI's 'E' columns
				sep[pair_index(a, b, nc)] = ell_rescaled_sep(vx, vy, sx[a], sy[a], rho[a], sx[b], sy[b], rho[b], sig[a], sig[b]);
			}
}

// Group normalisation over rows [r0, r0+n) whose log_post_weight v is stored in C.p_i (one warp).
// nwaylib/__init__.py:423-457.
__device__ __forceinline__ void group_normalise(const RowParams &R, long long r0, long long n, int lane)
{
	double *v = R.C.p_i;
	double m_all = -INFINITY, m_rest = -INFINITY;
	for (long long k = lane; k < n; k += 32) {
		double x = v[r0 + k];
		m_all = fmax(m_all, x);
		if (k > 0) m_rest = fmax(m_rest, x);
	}
	m_all = warp_max(m_all);
	m_rest = warp_max(m_rest);
	double s_all = 0.0, s_rest = 0.0;
	for (long long k = lane; k < n; k += 32) {
		double x = v[r0 + k];
		s_all += exp10(x - m_all);
		if (k > 0) s_rest += exp10(x - m_rest);
	}
	s_all = warp_sum(s_all);
	s_rest = warp_sum(s_rest);
	double bfsum = log10(s_all) + m_all;
	double bfsum1 = n > 1 ? log10(s_rest) + m_rest : 0.0;
	double v0 = v[r0];
	double p_any = 1 - exp10(v0 - bfsum);
	__syncwarp();
	double best = 0.0;   // p_i[0] = 0 always takes part in the max
	for (long long k = lane; k < n; k += 32) {
		double pi = k == 0 ? 0.0 : exp10(v[r0 + k] - bfsum1);
		v[r0 + k] = pi;
		R.C.p_any[r0 + k] = p_any;
		best = fmax(best, pi);
	}
	best = warp_max(best);
	for (long long k = lane; k < n; k += 32) {
		double pi = v[r0 + k];   // written by this same lane
		R.C.flag[r0 + k] = (pi == best) ? 1 : (pi > R.ratio_secondary * best ? 2 : 0);
	}
}

// ---------------------------------------------------------------------------------------------------------
// K2: rows
// ---------------------------------------------------------------------------------------------------------
template <int NC, bool FUSE>
__global__ void __launch_bounds__(256)
k_rows(RowParams R)
{
	if (!gate_open(R.gate)) return;
	const int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	const ConstTables *__restrict__ T = R.T;
	for (int p = warp; p < R.np; p += nwarps) {
		int nl[NC];
		long long lo[NC];
		long long ntup = 1;
		nl[0] = 0; lo[0] = 0;
#pragma unroll
		for (int c = 1; c < NC; c++) {
			lo[c] = R.L.off[c][p];
			nl[c] = (int) (R.L.off[c][p + 1] - lo[c]);
			ntup *= nl[c] + 1;
		}
		if (ntup <= R.small_t) continue;   // k_rows_small's
		const long long rbase = R.row_off[p];
		const double *mat = NC > 2 ? R.mat + R.mat_off[p] : nullptr;
		const long long gp = R.first + p;
		const double sig0 = R.err[0][gp];
		const long long wp = R.flat.err > 0.0 ? flat_hash_pack(R.ra[0][gp], R.dec[0][gp], R.flat) : 0ll;
		long long written = 0;
		for (long long t0 = 0; t0 < ntup; t0 += 32) {
			long long t = t0 + lane;
			bool ok = t < ntup;
			int dg[NC];
			long long sidx[NC];
			double sep[NC * (NC - 1) / 2];
			double sig[NC];
			unsigned present = 1u;
			sidx[0] = gp;
			sig[0] = sig0;
			if (ok) {
				long long rem = t;
#pragma unroll
				for (int c = NC - 1; c >= 1; c--) {
					dg[c] = (int) (rem % (nl[c] + 1));
					rem /= (nl[c] + 1);
				}
#pragma unroll
				for (int c = 1; c < NC; c++) {
					if (dg[c] > 0) {
						long long e = lo[c] + dg[c] - 1;
						sidx[c] = R.L.s[c][e];
						sep[pair_index(0, c, NC)] = R.L.sep[c][e];
						present |= 1u << c;
					} else {
						sidx[c] = -1;
						sep[pair_index(0, c, NC)] = nan("");
					}
				}
				if (NC > 2) {
					long long bo = 0;
#pragma unroll
					for (int a = 1; a < NC; a++) {
#pragma unroll
						for (int b = a + 1; b < NC; b++) {
							double s = nan("");
							if (dg[a] > 0 && dg[b] > 0) {
								s = mat[bo + (long long) (dg[a] - 1) * nl[b] + (dg[b] - 1)];
								ok = ok && (s < R.pair_radius[pair_index(a, b, NC)]);
							}
							sep[pair_index(a, b, NC)] = s;
							bo += (long long) nl[a] * nl[b];
						}
					}
				}
				if (R.flat.err > 0.0) ok = ok && tuple_in_one_bucket<NC>(R, wp, dg, lo);
			}
			unsigned m = __ballot_sync(NWB_FULL, ok);
			if (ok) {
				long long row = rbase + written + __popc(m & ((1u << lane) - 1));
				double smax = 0.0;
				if (R.sep_f32) {
#pragma unroll
					for (int k = 0; k < NC * (NC - 1) / 2; k++) sep[k] = (double) (float) sep[k];   // NaN stays NaN
				}
#pragma unroll
				for (int c = 0; c < NC; c++) R.C.idx[c][row] = sidx[c];
#pragma unroll
				for (int k = 0; k < NC * (NC - 1) / 2; k++) {
					R.C.sep[k][row] = sep[k];
					if (sep[k] > smax) smax = sep[k];   // NaN compares false: nanmax
				}
				R.C.sepmax[row] = smax;
				R.C.ncat[row] = __popc(present);
				double lbf;
				if (R.ell) {
					double esig[MAXC], esep[MAXP];
					ell_prepare(R, present, sidx, esig, esep);
					lbf = log_bf_ref<0>(T, NC, present, esig, esep);
				} else {
#pragma unroll
					for (int c = 1; c < NC; c++)
						if (present >> c & 1u) sig[c] = R.err[c][sidx[c]];
					lbf = log_bf_ref<NC>(T, NC, present, sig, sep, R.sep_f32 != 0);
				}
				unsigned smask = present >> 1;
				double prior = T->prior[smask], l10p = T->log10prior[smask];
				R.C.lbf_u[row] = lbf;
				R.C.lbf[row] = lbf;
				R.C.dist_post[row] = posterior_ref(prior, l10p, lbf);
				if (FUSE) {
					double total = lbf + row_bias(R, row, sidx);
					R.C.p_single[row] = posterior_ref(prior, l10p, total);
					R.C.p_i[row] = total + l10p;
				}
			}
			written += __popc(m);
		}
		if (FUSE) {
			__syncwarp();
			group_normalise(R, rbase, written, lane);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// K2 for sparse primaries: ONE THREAD per primary with at most R.small_t (<= SMALL_T) candidate tuples.  An all-sky
// match leaves ~99 % of the primaries with nothing but their no-counterpart row; a warp per primary (k_rows) then runs
// one lane in 32 behind a chain of dependent loads.  Here 32 primaries share a warp, the stores of consecutive
// primaries coalesce, and the group normalisation is a scalar loop over <= SMALL_T rows with the reference's formulas
// (__init__.py:423-457).  Same rows, same order, same arithmetic as k_rows.
// ---------------------------------------------------------------------------------------------------------
constexpr int SMALL_T = 8;

template <int NC, bool FUSE>
__global__ void __launch_bounds__(128)
k_rows_small(RowParams R)
{
	if (!gate_open(R.gate)) return;
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= R.np) return;
	const ConstTables *__restrict__ T = R.T;
	int nl[NC];
	long long lo[NC];
	int ntup = 1;
	nl[0] = 0; lo[0] = 0;
#pragma unroll
	for (int c = 1; c < NC; c++) {
		lo[c] = R.L.off[c][p];
		nl[c] = (int) (R.L.off[c][p + 1] - lo[c]);
		ntup = ntup <= SMALL_T ? ntup * (nl[c] + 1) : ntup;
	}
	if (ntup > R.small_t) return;
	const long long rbase = R.row_off[p];
	const double *mat = NC > 2 ? R.mat + R.mat_off[p] : nullptr;
	const long long gp = R.first + p;
	const long long wp = R.flat.err > 0.0 ? flat_hash_pack(R.ra[0][gp], R.dec[0][gp], R.flat) : 0ll;
	double v[SMALL_T];
	int nrow = 0;
	for (int t = 0; t < ntup; t++) {
		int dg[NC];
		long long sidx[NC];
		double sep[NC * (NC - 1) / 2];
		double sig[NC];
		unsigned present = 1u;
		bool ok = true;
		sidx[0] = gp;
		int rem = t;
#pragma unroll
		for (int c = NC - 1; c >= 1; c--) {
			dg[c] = rem % (nl[c] + 1);
			rem /= (nl[c] + 1);
		}
#pragma unroll
		for (int c = 1; c < NC; c++) {
			if (dg[c] > 0) {
				long long e = lo[c] + dg[c] - 1;
				sidx[c] = R.L.s[c][e];
				sep[pair_index(0, c, NC)] = R.L.sep[c][e];
				present |= 1u << c;
			} else {
				sidx[c] = -1;
				sep[pair_index(0, c, NC)] = nan("");
			}
		}
		if (NC > 2) {
			long long bo = 0;
#pragma unroll
			for (int a = 1; a < NC; a++) {
#pragma unroll
				for (int b = a + 1; b < NC; b++) {
					double s = nan("");
					if (dg[a] > 0 && dg[b] > 0) {
						s = mat[bo + (long long) (dg[a] - 1) * nl[b] + (dg[b] - 1)];
						ok = ok && (s < R.pair_radius[pair_index(a, b, NC)]);
					}
					sep[pair_index(a, b, NC)] = s;
					bo += (long long) nl[a] * nl[b];
				}
			}
		}
		if (R.flat.err > 0.0) ok = ok && tuple_in_one_bucket<NC>(R, wp, dg, lo);
		if (!ok) continue;
		const long long row = rbase + nrow;
		double smax = 0.0;
		if (R.sep_f32) {
#pragma unroll
			for (int k = 0; k < NC * (NC - 1) / 2; k++) sep[k] = (double) (float) sep[k];
		}
#pragma unroll
		for (int c = 0; c < NC; c++) R.C.idx[c][row] = sidx[c];
#pragma unroll
		for (int k = 0; k < NC * (NC - 1) / 2; k++) {
			R.C.sep[k][row] = sep[k];
			if (sep[k] > smax) smax = sep[k];
		}
		R.C.sepmax[row] = smax;
		R.C.ncat[row] = __popc(present);
		double lbf = 0.0;
		if (present != 1u) {
			if (R.ell) {
				double esig[MAXC], esep[MAXP];
				ell_prepare(R, present, sidx, esig, esep);
				lbf = log_bf_ref<0>(T, NC, present, esig, esep);
			} else {
				sig[0] = R.err[0][gp];
#pragma unroll
				for (int c = 1; c < NC; c++)
					if (present >> c & 1u) sig[c] = R.err[c][sidx[c]];
				lbf = log_bf_ref<NC>(T, NC, present, sig, sep, R.sep_f32 != 0);
			}
		}
		const unsigned smask = present >> 1;
		const double prior = T->prior[smask], l10p = T->log10prior[smask];
		R.C.lbf_u[row] = lbf;
		R.C.lbf[row] = lbf;
		R.C.dist_post[row] = posterior_ref(prior, l10p, lbf);
		if (FUSE) {
			double total = lbf + row_bias(R, row, sidx);
			R.C.p_single[row] = posterior_ref(prior, l10p, total);
			v[nrow] = total + l10p;
		}
		nrow++;
	}
	if (!FUSE) return;
	// group normalisation, scalar (__init__.py:423-457)
	double m_all = -INFINITY, m_rest = -INFINITY;
	for (int k = 0; k < nrow; k++) {
		m_all = fmax(m_all, v[k]);
		if (k > 0) m_rest = fmax(m_rest, v[k]);
	}
	double s_all = 0.0, s_rest = 0.0;
	for (int k = 0; k < nrow; k++) {
		s_all += exp10(v[k] - m_all);
		if (k > 0) s_rest += exp10(v[k] - m_rest);
	}
	const double bfsum = log10(s_all) + m_all;
	const double bfsum1 = nrow > 1 ? log10(s_rest) + m_rest : 0.0;
	const double p_any = 1 - exp10(v[0] - bfsum);
	double best = 0.0;
	for (int k = 1; k < nrow; k++) {
		v[k] = exp10(v[k] - bfsum1);
		best = fmax(best, v[k]);
	}
	v[0] = 0.0;
	for (int k = 0; k < nrow; k++) {
		R.C.p_i[rbase + k] = v[k];
		R.C.p_any[rbase + k] = p_any;
		R.C.flag[rbase + k] = (v[k] == best) ? 1 : (v[k] > R.ratio_secondary * best ? 2 : 0);
	}
}

// ---------------------------------------------------------------------------------------------------------
// K2 for two catalogues (the streaming configuration): one warp per primary, everything of the group staged in
// shared memory: the unsorted matches are rank-sorted by secondary index (fastskymatch.py:181,217), rows are
// scored and written with coalesced stores, and -- FUSE -- the group's log-sum-exp runs on the shared copy of
// the log-weights, so no column is ever read back from global memory.
// ---------------------------------------------------------------------------------------------------------
constexpr int R2_WARPS = 8;
constexpr int R2_CAP = 128;   // matches per primary handled in shared memory; larger groups take k_rows2_big

constexpr int R2_NB = 128;    // buckets of the in-group sort
#ifndef NWB_R2_MINBLOCKS
#define NWB_R2_MINBLOCKS 4
#endif
#ifndef NWB_R2_GRID_PER_SM
#define NWB_R2_GRID_PER_SM 4
#endif


struct __align__(16) R2Smem {
	int s_in[R2_CAP];
	int s[R2_CAP];
	double sep[R2_CAP];
	double v[R2_CAP + 2];           // (+2 keeps hist 16-byte aligned)
	int hist[R2_NB + 4];            // bucket counts, then exclusive bucket starts (R2_NB + 1 used)
	int key[R2_CAP];                // the indices again, in bucket order
	unsigned char pos[R2_CAP];      // arrival position of an element inside its bucket
};

template <bool FUSE, bool SHARE>
__global__ void __launch_bounds__(R2_WARPS * 32, NWB_R2_MINBLOCKS)
k_rows2(RowParams R)
{
	__shared__ R2Smem smem[R2_WARPS];
	const int lane = threadIdx.x & 31;
	R2Smem &M = smem[threadIdx.x >> 5];
	const ConstTables *__restrict__ T = R.T;
	if (!guard_ok(R)) return;
	R2Memo memo;
	double m_sig0 = -1.0, w0 = 0.0, lw0 = 0.0;
	const int nwarps = gridDim.x * R2_WARPS;
	for (int p = blockIdx.x * R2_WARPS + (threadIdx.x >> 5); p < R.np; p += nwarps) {
	const int n = R.S1.cnt[p];
	const long long rbase = R.row_off[p];
	const long long gp = R.first + p;
	const double sig0 = R.err[0][gp];
	if (sig0 != m_sig0) {
		m_sig0 = sig0;
		w0 = 1.0 / (sig0 * sig0);
		lw0 = log(w0);
	}
	__syncwarp();
	if (n <= R2_CAP) {
		// Sort the group by secondary index: bucket the indices over [min, max] of the group (monotone map, so an
		// element's rank = elements in lower buckets + smaller elements of its own bucket), ~1 element per bucket.
		int smin = 0x7fffffff, smax = -1;
		for (int e = lane; e < n; e += 32) {
			Slot16 x = store_get(R.S1, p, e);
			M.s_in[e] = x.s;
			M.v[e] = x.sep;   // parked here until ranked
			smin = min(smin, x.s);
			smax = max(smax, x.s);
		}
		*reinterpret_cast<int4 *>(&M.hist[4 * lane]) = make_int4(0, 0, 0, 0);
		smin = __reduce_min_sync(NWB_FULL, smin);
		smax = __reduce_max_sync(NWB_FULL, smax);
		const float bscale = (float) R2_NB / ((float) (smax - smin) + 1.0f);
		__syncwarp();
		for (int e = lane; e < n; e += 32) {
			int b = min(R2_NB - 1, (int) ((float) (M.s_in[e] - smin) * bscale));
			M.pos[e] = (unsigned char) atomicAdd(&M.hist[b], 1);
		}
		__syncwarp();
		{
			int4 c = *reinterpret_cast<const int4 *>(&M.hist[4 * lane]);
			int tot = c.x + c.y + c.z + c.w;
			int incl = tot;
			for (int o = 1; o < 32; o <<= 1) {
				int y = __shfl_up_sync(NWB_FULL, incl, o);
				if (lane >= o) incl += y;
			}
			int base = incl - tot;
			__syncwarp();
			*reinterpret_cast<int4 *>(&M.hist[4 * lane]) = make_int4(base, base + c.x, base + c.x + c.y, base + c.x + c.y + c.z);
			if (lane == 31) M.hist[R2_NB] = incl;
		}
		__syncwarp();
		for (int e = lane; e < n; e += 32) {
			int b = min(R2_NB - 1, (int) ((float) (M.s_in[e] - smin) * bscale));
			M.key[M.hist[b] + M.pos[e]] = M.s_in[e];
		}
		__syncwarp();
		for (int e = lane; e < n; e += 32) {
			int mine = M.s_in[e];
			int b = min(R2_NB - 1, (int) ((float) (mine - smin) * bscale));
			int lo = M.hist[b], hi = M.hist[b + 1];
			int rank = lo;
			for (int j = lo; j < hi; j++) rank += M.key[j] < mine;
			M.s[rank] = mine;
			M.sep[rank] = M.v[e];
		}
		__syncwarp();
		// rows in order: row 0 = no counterpart, row k = k-th smallest secondary index
		const int rows = n + 1;
		double m_rest = -INFINITY;   // max of the log-weights of rows 1.. (this lane's share)
		for (int k = lane; k < rows; k += 32) {
			double v = 0.0;
			rows2_write<FUSE, SHARE>(R, T, rbase + k, gp, k == 0 ? -1 : (long long) M.s[k - 1], k == 0 ? 0.0 : M.sep[k - 1],
				w0, lw0, memo, v);
			if (FUSE) {
				M.v[k] = v;
				if (k > 0) m_rest = fmax(m_rest, v);
			}
		}
		if (!FUSE) continue;
		// Group normalisation (__init__.py:423-457) on the shared copy of the log-weights, with ONE exp10 per row:
		//   t_k = 10^(v_k - m_rest) (k >= 1),  s_rest = sum t_k,  p_i = t_k / s_rest  [= 10^(v_k - bfsum1), since
		//   bfsum1 = log10(s_rest) + m_rest],  s_all = s_rest * 10^(m_rest - m_all) + 10^(v_0 - m_all).
		// Same numbers as the reference's formulas up to the rounding of the exponent arguments (~1e-14 relative,
		// four orders below the parity tolerance).  Every lane only touches its own k = lane + 32 j, except v[0].
		__syncwarp();
		const double v0 = M.v[0];
		m_rest = warp_max(m_rest);
		double s_rest = 0.0;
		for (int k = lane + (lane == 0 ? 32 : 0); k < rows; k += 32) {
			double t = nwb_exp10(M.v[k] - m_rest);
			M.v[k] = t;
			s_rest += t;
		}
		s_rest = warp_sum(s_rest);
		double p_any, rinv;
		group_p_any(rows, v0, m_rest, s_rest, p_any, rinv);
		// lone no-counterpart row: bfsum = v0 exactly, p_any = 1 - 10^0 = 0
		__syncwarp();
		const double best = rinv;   // the largest t_k is exactly 1, so max p_i = 1 * rinv (0 for a lone row)
		// SHARE: dist_post = 1/(1 + (1 - prior) 10^(-v)) with 10^(-v_k) = 10^(-m_rest) / t_k, i.e.
		// t_k / (t_k + (1 - prior) 10^(-m_rest)): no second exponential per row.  Outside |m_rest| <= 250 (10^(-m_rest)
		// near the ends of the double range) and for t_k = 0 the direct form is used.
		const bool direct = SHARE && !(fabs(m_rest) <= 250.0);
		const double oscale = (SHARE && rows > 1 && !direct) ? (1 - T->prior[1]) * nwb_exp10(-m_rest) : 0.0;
		const double omp = 1 - T->prior[1], l10p1 = T->log10prior[1];
		for (int k = lane; k < rows; k += 32) {
			long long row = rbase + k;
			double t = M.v[k];
			double pi = k == 0 ? 0.0 : t * rinv;
			R.C.p_i[row] = pi;
			R.C.p_any[row] = p_any;
			R.C.flag[row] = (pi == best) ? 1 : (pi > R.ratio_secondary * best ? 2 : 0);
			if (SHARE) {
				double post = 1.0;   // row 0: prior = 1, (1 - prior) * 10^0 = 0
				if (k > 0) post = shared_post(direct, t, oscale, omp, R.C.lbf[row], l10p1);
				R.C.dist_post[row] = post;
				R.C.p_single[row] = post;
			}
		}
	} else {
		// big group: rank against global memory, write rows, normalise through the p_i column
		for (int e = lane; e < n; e += 32) {
			Slot16 me = store_get(R.S1, p, e);
			int rank = 0;
			for (int f = 0; f < n; f++) rank += store_get(R.S1, p, f).s < me.s;
			double v = 0.0;
			rows2_write<FUSE, false>(R, T, rbase + 1 + rank, gp, me.s, me.sep, w0, lw0, memo, v);
			if (FUSE) R.C.p_i[rbase + 1 + rank] = v;
		}
		if (lane == 0) {
			double v = 0.0;
			rows2_write<FUSE, false>(R, T, rbase, gp, -1, 0.0, w0, lw0, memo, v);
			if (FUSE) R.C.p_i[rbase] = v;
		}
		if (FUSE) {
			__syncwarp();
			group_normalise(R, rbase, (long long) n + 1, lane);
		}
	}
	}
}

// K3 (unfused): one warp per primary
__global__ void __launch_bounds__(256)
k_final(RowParams R)
{
	if (!gate_open(R.gate)) return;
	const int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	const ConstTables *__restrict__ T = R.T;
	for (int p = warp; p < R.np; p += nwarps) {
		long long r0 = R.row_off[p], n = R.row_off[p + 1] - r0;
		if (n <= R.small_t) continue;   // k_final_small's
		for (long long k = lane; k < n; k += 32) {
			long long row = r0 + k;
			long long sidx[MAXC];
			unsigned smask = 0;
			for (int c = 0; c < R.ncat; c++) {
				sidx[c] = R.C.idx[c][row];
				if (c > 0 && sidx[c] >= 0) smask |= 1u << (c - 1);
			}
			double total = R.C.lbf[row] + row_bias(R, row, sidx);
			double prior = T->prior[smask], l10p = T->log10prior[smask];
			R.C.p_single[row] = posterior_ref(prior, l10p, total);
			R.C.p_i[row] = total + l10p;
		}
		__syncwarp();
		group_normalise(R, r0, n, lane);
	}
}

// K3 for sparse primaries: ONE THREAD per primary with at most R.small_t rows (see k_rows_small): an all-sky match leaves
// ~99 % of the groups with nothing but their no-counterpart row, for which a warp per primary is 31 idle lanes behind a
// chain of dependent loads.  Same formulas (__init__.py:423-457), scalar.
__global__ void __launch_bounds__(128)
k_final_small(RowParams R)
{
	if (!gate_open(R.gate)) return;
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= R.np) return;
	const ConstTables *__restrict__ T = R.T;
	const long long r0 = R.row_off[p];
	const int n = (int) min((long long) (SMALL_T + 1), R.row_off[p + 1] - r0);
	if (n > R.small_t || n <= 0) return;
	double v[SMALL_T];
	for (int k = 0; k < n; k++) {
		const long long row = r0 + k;
		long long sidx[MAXC];
		unsigned smask = 0;
		for (int c = 0; c < R.ncat; c++) {
			sidx[c] = R.C.idx[c][row];
			if (c > 0 && sidx[c] >= 0) smask |= 1u << (c - 1);
		}
		const double total = R.C.lbf[row] + row_bias(R, row, sidx);
		const double prior = T->prior[smask], l10p = T->log10prior[smask];
		R.C.p_single[row] = posterior_ref(prior, l10p, total);
		v[k] = total + l10p;
	}
	double m_all = -INFINITY, m_rest = -INFINITY;
	for (int k = 0; k < n; k++) {
		m_all = fmax(m_all, v[k]);
		if (k > 0) m_rest = fmax(m_rest, v[k]);
	}
	double s_all = 0.0, s_rest = 0.0;
	for (int k = 0; k < n; k++) {
		s_all += exp10(v[k] - m_all);
		if (k > 0) s_rest += exp10(v[k] - m_rest);
	}
	const double bfsum = log10(s_all) + m_all;
	const double bfsum1 = n > 1 ? log10(s_rest) + m_rest : 0.0;
	const double p_any = 1 - exp10(v[0] - bfsum);
	double best = 0.0;
	for (int k = 1; k < n; k++) {
		v[k] = exp10(v[k] - bfsum1);
		best = fmax(best, v[k]);
	}
	v[0] = 0.0;
	for (int k = 0; k < n; k++) {
		R.C.p_i[r0 + k] = v[k];
		R.C.p_any[r0 + k] = p_any;
		R.C.flag[r0 + k] = (v[k] == best) ? 1 : (v[k] > R.ratio_secondary * best ? 2 : 0);
	}
}

// K2b: the CLI's unrelated-association correction (nway.py:366-421), one warp per primary.
// For every set M of >= 2 secondary catalogues: best(M) = max(0, max over rows j with ncat_j > 2 of
// log_bf(sub-association M & present_j of row j) + log10(nu[A0]/prod nu_plus[A])), taken when the
// intersection has >= 2 members; every row whose ABSENT set is exactly M gets best(M) added.
__global__ void k_correct_cli(RowParams R)
{
	if (!gate_open(R.gate)) return;
	const int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	const ConstTables *__restrict__ T = R.T;
	const int nc = R.ncat;
	const unsigned all = (1u << (nc - 1)) - 1;   // over secondaries, bit c-1
	for (int p = warp; p < R.np; p += nwarps) {
		long long r0 = R.row_off[p], n = R.row_off[p + 1] - r0;
		if (n < 3) continue;   // a correction needs a row with two more secondaries than another (nway.py:380-400): none in a group of 1 or 2
		for (unsigned M = 3; M <= all; M++) {
			if (__popc(M) < 2) continue;
			// does any row of the group miss exactly M?
			bool need = false;
			for (long long k = lane; k < n; k += 32) {
				unsigned pres = 0;
				for (int c = 1; c < nc; c++)
					if (R.C.idx[c][r0 + k] >= 0) pres |= 1u << (c - 1);
				need = need || ((~pres & all) == M);
			}
			if (!__any_sync(NWB_FULL, need)) continue;
			double best = 0.0;
			for (long long k = lane; k < n; k += 32) {
				long long row = r0 + k;
				if (!(R.C.ncat[row] > 2)) continue;
				unsigned pres = 0;
				for (int c = 1; c < nc; c++)
					if (R.C.idx[c][row] >= 0) pres |= 1u << (c - 1);
				unsigned A = M & pres;
				if (__popc(A) < 2) continue;
				double sig[MAXC], sep[MAXP];
				if (R.ell) {
					long long sidx[MAXC];
					for (int c = 1; c < nc; c++) sidx[c] = R.C.idx[c][row];
					ell_prepare(R, A << 1, sidx, sig, sep);
				} else {
					for (int c = 1; c < nc; c++)
						if (A >> (c - 1) & 1u) sig[c] = R.err[c][R.C.idx[c][row]];
					for (int a = 1; a < nc; a++)
						for (int b = a + 1; b < nc; b++)
							if ((A >> (a - 1) & 1u) && (A >> (b - 1) & 1u))
								sep[pair_index(a, b, nc)] = R.C.sep[pair_index(a, b, nc)][row];
				}
				double lp = log_bf_ref<0>(T, nc, A << 1, sig, sep) + T->sub_log10prior[A];
				if (lp > best) best = lp;
			}
			best = warp_max(best);
			if (best > 0) {
				for (long long k = lane; k < n; k += 32) {
					long long row = r0 + k;
					unsigned pres = 0;
					for (int c = 1; c < nc; c++)
						if (R.C.idx[c][row] >= 0) pres |= 1u << (c - 1);
					if ((~pres & all) == M) {
						double lbf = R.C.lbf[row] + best;
						R.C.lbf[row] = lbf;
						R.C.dist_post[row] = posterior_ref(T->prior[pres], T->log10prior[pres], lbf);
					}
				}
			}
			__syncwarp();
		}
	}
}

// min / max of a column (is a catalogue's positional error one constant?)
__global__ void k_minmax(long long n, const double *__restrict__ x, double *__restrict__ out /* [2], pre-set */)
{
	double lo = 1e300, hi = -1e300;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		double v = x[i];
		if (!(v == v)) { lo = -1e300; hi = 1e300; }   // NaN: never "constant"
		lo = fmin(lo, v); hi = fmax(hi, v);
	}
	for (int o = 16; o > 0; o >>= 1) {
		lo = fmin(lo, __shfl_xor_sync(NWB_FULL, lo, o));
		hi = fmax(hi, __shfl_xor_sync(NWB_FULL, hi, o));
	}
	if ((threadIdx.x & 31) == 0) {
		// doubles of one sign order like their bit patterns; errors are positive, anything else disables the shortcut
		if (lo > 0) atomicMin(reinterpret_cast<unsigned long long *>(out), (unsigned long long) __double_as_longlong(lo));
		else atomicMin(reinterpret_cast<unsigned long long *>(out), 0ull);
		if (hi > 0) atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long) __double_as_longlong(hi));
		else atomicMax(reinterpret_cast<unsigned long long *>(out + 1), 0x7ff0000000000000ull);
	}
}

// smallest / largest ra, largest |dec| and the number of NaNs of a catalogue: what the reference's choice between its
// flat-sky and its HEALPix hash depends on (fastskymatch.py:94-98).  Order-preserving integer images as in k_prim_prep.
__global__ void k_radec_bounds(long long n, const double *__restrict__ ra, const double *__restrict__ dec,
	unsigned long long *__restrict__ out /* [0] min ra (pre-set ~0), [1] max ra, [2] max |dec|, [3] NaNs (pre-set 0) */)
{
	double lo = INFINITY, hi = -INFINITY, ad = 0.0;
	unsigned long long nans = 0;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const double r = ra[i], d = dec[i];
		if (r != r || d != d) nans++;
		lo = fmin(lo, r); hi = fmax(hi, r); ad = fmax(ad, fabs(d));
	}
	for (int o = 16; o > 0; o >>= 1) {
		lo = fmin(lo, __shfl_xor_sync(NWB_FULL, lo, o));
		hi = fmax(hi, __shfl_xor_sync(NWB_FULL, hi, o));
		ad = fmax(ad, __shfl_xor_sync(NWB_FULL, ad, o));
		nans += __shfl_xor_sync(NWB_FULL, nans, o);
	}
	if ((threadIdx.x & 31) == 0) {
		auto key = [](double x) {
			unsigned long long u = (unsigned long long) __double_as_longlong(x);
			return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
		};
		atomicMin(out + 0, key(lo));
		atomicMax(out + 1, key(hi));
		atomicMax(out + 2, key(ad));
		if (nans) atomicAdd(out + 3, nans);
	}
}

// ---------------------------------------------------------------------------------------------------------
// N1: automatic magnitude histograms (nwaylib/__init__.py:324-366, nway.py:455-503) -- the selection half on the device
// ---------------------------------------------------------------------------------------------------------
// per row: bit 0 = the catalogue has a counterpart in this row, bit 1 = row selected as a secure match, bit 2 = row
// "possible" (its source is removed from the field-source histogram); isel / idef feed the two prefix sums
__global__ void k_hist_flags(long long nrows, const long long *__restrict__ res, const double *__restrict__ sepmax,
	const double *__restrict__ dist_post, int by_radius, double thr_sel, double thr_possible,
	unsigned char *__restrict__ flag, int *__restrict__ isel, int *__restrict__ idef)
{
	long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) return;
	const bool def = res[r] != -1;
	const double x = by_radius ? sepmax[r] : dist_post[r];
	const bool sel = def && (by_radius ? x < thr_sel : x > thr_sel);
	const bool pos = def && (by_radius ? x < thr_possible : x > thr_possible);
	flag[r] = (unsigned char) ((def ? 1 : 0) | (sel ? 2 : 0) | (pos ? 4 : 0));
	isel[r] = sel;
	idef[r] = def;
}

// numpy.unique(res[selection], return_index=True): the first occurrence of a source inside the selected sub-array is
// the smallest selected-position -> atomicMin.  Weights: the reference compresses them by res_defined (API,
// __init__.py:337) or by the selection (command-line program, nway.py:471) and then indexes them with the
// selected-positions -- W holds that compressed array.
__global__ void k_hist_mark(long long nrows, const long long *__restrict__ res, const unsigned char *__restrict__ flag,
	const int *__restrict__ selpos, const int *__restrict__ defpos, const double *__restrict__ sw, int weights_cli,
	int *__restrict__ first, unsigned char *__restrict__ possible, double *__restrict__ W)
{
	long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) return;
	const unsigned f = flag[r];
	if (!(f & 1u)) return;
	const long long s = res[r];
	if (sw && !weights_cli) W[defpos[r]] = sw[r];
	if (f & 2u) {
		atomicMin(&first[s], selpos[r]);
		if (sw && weights_cli) W[selpos[r]] = sw[r];
	}
	if (f & 4u) possible[s] = 1;
}

// per source of the catalogue: selected?  field source (finite magnitude, not "possible")?  min / max of the field
// sources' magnitudes (order-preserving integer images, as in k_prim_prep) and the three counts the reference logs
__global__ void k_hist_sources(long long n, const double *__restrict__ mag, const int *__restrict__ first,
	const unsigned char *__restrict__ possible, int *__restrict__ selflag, unsigned long long *__restrict__ stats
	/* [0] possible, [1] others, [2] valid, [3] min key, [4] max key; zeroed except [3] = ~0 */)
{
	long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long npos = 0, noth = 0, nval = 0, kmin = ~0ull, kmax = 0ull;
	if (s < n) {
		double m = mag[s];
		bool valid = isfinite(m) && m != -99.0;
		bool pos = possible[s] != 0;
		selflag[s] = first[s] != 0x7f7f7f7f;   // the memset pattern = never selected
		npos = pos; nval = valid;
		if (valid && !pos) {
			noth = 1;
			unsigned long long u = (unsigned long long) __double_as_longlong(m);
			u ^= (u >> 63) ? ~0ull : 0x8000000000000000ull;
			kmin = kmax = u;
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		npos += __shfl_xor_sync(NWB_FULL, npos, o);
		noth += __shfl_xor_sync(NWB_FULL, noth, o);
		nval += __shfl_xor_sync(NWB_FULL, nval, o);
		unsigned long long a = __shfl_xor_sync(NWB_FULL, kmin, o), b = __shfl_xor_sync(NWB_FULL, kmax, o);
		kmin = a < kmin ? a : kmin;
		kmax = b > kmax ? b : kmax;
	}
	if ((threadIdx.x & 31) == 0) {
		if (npos) atomicAdd(stats + 0, npos);
		if (noth) atomicAdd(stats + 1, noth);
		if (nval) atomicAdd(stats + 2, nval);
		if (noth) { atomicMin(stats + 3, kmin); atomicMax(stats + 4, kmax); }
	}
}

// the selected sources in ascending index order (= numpy.unique's order): magnitude and weight
__global__ void k_hist_gather(long long n, const double *__restrict__ mag, const int *__restrict__ first,
	const int *__restrict__ selflag, const int *__restrict__ selrank, const double *__restrict__ W,
	double *__restrict__ out_mag, double *__restrict__ out_w)
{
	long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || !selflag[s]) return;
	double m = mag[s];
	out_mag[selrank[s]] = m == -99.0 ? nan("") : m;
	out_w[selrank[s]] = W ? W[first[s]] : 1.0;
}

// numpy.histogram(mag[others], bins=edges): counts per bin, last bin closed on the right
__global__ void k_hist_count(long long n, const double *__restrict__ mag, const unsigned char *__restrict__ possible,
	int nbins, const double *__restrict__ edges, unsigned long long *__restrict__ counts)
{
	__shared__ unsigned int sc[MAXB];
	__shared__ double se[MAXB + 1];
	for (int k = threadIdx.x; k < nbins; k += blockDim.x) sc[k] = 0;
	for (int k = threadIdx.x; k <= nbins; k += blockDim.x) se[k] = edges[k];
	__syncthreads();
	for (long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long) gridDim.x * blockDim.x) {
		double m = mag[s];
		if (!(isfinite(m) && m != -99.0) || possible[s]) continue;
		if (!(m >= se[0] && m <= se[nbins])) continue;
		int k = 0;
		for (int j = 1; j < nbins; j++)
			if (se[j] <= m) k = j;
		atomicAdd(&sc[k], 1u);
	}
	__syncthreads();
	for (int k = threadIdx.x; k < nbins; k += blockDim.x)
		if (sc[k]) atomicAdd(counts + k, (unsigned long long) sc[k]);
}

// ---------------------------------------------------------------------------------------------------------
// a13 truncation + element-wise surface
// ---------------------------------------------------------------------------------------------------------
__global__ void k_keep_flags(long long n, const double *__restrict__ p_i, double min_prob, int *__restrict__ keep)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) keep[i] = !(p_i[i] < min_prob);
}

__global__ void k_compact8(long long n, const int *__restrict__ keep, const long long *__restrict__ pos,
	const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && keep[i]) dst[pos[i]] = src[i];
}

__global__ void k_dist(long long n, const double *__restrict__ ra1, const double *__restrict__ dec1,
	const double *__restrict__ ra2, const double *__restrict__ dec2, double *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double s1, c1, s2, c2, sd, cd;
	sincos_ref(deg2rad_ref(dec1[i]), &s1, &c1);
	sincos_ref(deg2rad_ref(dec2[i]), &s2, &c2);
	sincos_ref(deg2rad_ref(ra2[i]) - deg2rad_ref(ra1[i]), &sd, &cd);
	double num1 = c2 * sd;
	double num2 = c1 * s2 - s1 * c2 * cd;
	double den = s1 * s2 + c1 * c2 * cd;
	out[i] = atan2(hypot(num1, num2), den) * 180 / NWB_PI;
}

__global__ void k_log_bf(long long n, int ncat, const double *__restrict__ sep, const double *__restrict__ err,
	const ConstTables *__restrict__ T, double *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double sig[MAXC], sp[MAXP];
	for (int c = 0; c < ncat; c++) sig[c] = err[(long long) c * n + i];
	for (int a = 0; a < ncat; a++)
		for (int b = a + 1; b < ncat; b++)
			sp[pair_index(a, b, ncat)] = sep[((long long) a * ncat + b) * n + i];
	out[i] = log_bf_ref<0>(T, ncat, (1u << ncat) - 1, sig, sp);
}

// bayesdistance.log_bf_elliptical (:207-240) on caller-supplied offsets: sep_ra / sep_dec are ncat*ncat blocks of n
// (only a < b read), err is ncat blocks of (sigma_x | sigma_y | rho), each n long
__global__ void k_log_bf_ell(long long n, int ncat, const double *__restrict__ sra, const double *__restrict__ sde,
	const double *__restrict__ err, const ConstTables *__restrict__ T, double *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double sx[MAXC], sy[MAXC], rho[MAXC], sig[MAXC], sp[MAXP];
	for (int c = 0; c < ncat; c++) {
		sx[c] = err[((long long) c * 3 + 0) * n + i];
		sy[c] = err[((long long) c * 3 + 1) * n + i];
		rho[c] = err[((long long) c * 3 + 2) * n + i];
		sig[c] = sqrt((sx[c] * sx[c] + sy[c] * sy[c]) / 2);
	}
	for (int a = 0; a < ncat; a++)
		for (int b = a + 1; b < ncat; b++) {
			long long k = ((long long) a * ncat + b) * n + i;
			sp[pair_index(a, b, ncat)] = ell_rescaled_sep(sra[k], sde[k], sx[a], sy[a], rho[a], sx[b], sy[b], rho[b], sig[a], sig[b]);
		}
	out[i] = log_bf_ref<0>(T, ncat, (1u << ncat) - 1, sig, sp);
}

// tangent-plane offsets (arcsec) of the rows of the result table for the catalogue pair (a, b), a < b, measured like
// fastskymatch.match_multiple does for circular=False (:299-331): frame centred on the source of the later catalogue
// b, the earlier one is the target.  NaN where either source is absent.
__global__ void k_row_offsets(long long nrows, const long long *__restrict__ ia, const long long *__restrict__ ib,
	const double *__restrict__ ra_a, const double *__restrict__ dec_a, const double *__restrict__ ra_b,
	const double *__restrict__ dec_b, double *__restrict__ dra, double *__restrict__ ddec)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nrows) return;
	long long sa = ia[i], sb = ib[i];
	double x = nan(""), y = nan("");
	if (sa >= 0 && sb >= 0) offsets_ref(ra_b[sb], dec_b[sb], ra_a[sa], dec_a[sa], x, y);
	dra[i] = x;
	ddec[i] = y;
}

// Score a caller-supplied candidate list (nwb_score_rows): one thread per row of `idx` (R x ncat row indices, -1 = the
// catalogue takes no part) -- what the reference computes for every tuple crossproduct() hands it, BEFORE the radius
// filter: the separations of every pair (fastskymatch.py:26-47 via __init__.py:143-168), Separation_max (:166), ncat
// (:177), the log Bayes factor of the present catalogues and its prior (__init__.py:220-259, bayesdistance.py:64-86),
// dist_post (bayesdistance.py:26-32).  Separates an arithmetic mismatch from an enumeration mismatch.
__global__ void k_score_rows(RowParams R, long long nrows, const long long *__restrict__ idx, double *__restrict__ sep_out /* npairs x nrows or null */,
	double *__restrict__ sepmax, long long *__restrict__ ncat_out, double *__restrict__ lbf_out, double *__restrict__ post_out)
{
	const long long row = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= nrows) return;
	const int nc = R.ncat;
	long long sidx[MAXC];
	double lon[MAXC], sl[MAXC], cl[MAXC], sig[MAXC], sep[MAXP];
	unsigned present = 0;
	for (int c = 0; c < nc; c++) {
		const long long i = idx[row * nc + c];
		sidx[c] = i;
		if (i >= 0 && i < R.n[c]) {
			present |= 1u << c;
			lon[c] = deg2rad_ref(R.ra[c][i]);
			sincos_ref(deg2rad_ref(R.dec[c][i]), &sl[c], &cl[c]);
			if (!R.ell) sig[c] = R.err[c][i];
		} else {
			sidx[c] = -1;
		}
	}
	double smax = 0.0;
	for (int a = 0; a < nc; a++)
		for (int b = a + 1; b < nc; b++) {
			double v = nan("");
			if ((present >> a & 1u) && (present >> b & 1u)) {
				v = sep_arcsec_ref(lon[a], sl[a], cl[a], lon[b], sl[b], cl[b]);
				if (R.sep_f32) v = (double) (float) v;
				if (v > smax) smax = v;
			}
			sep[pair_index(a, b, nc)] = v;
			if (sep_out) sep_out[(long long) pair_index(a, b, nc) * nrows + row] = v;
		}
	sepmax[row] = smax;
	ncat_out[row] = __popc(present);
	double lbf;
	if (R.ell) {
		double esig[MAXC], esep[MAXP];
		ell_prepare(R, present, sidx, esig, esep);
		lbf = log_bf_ref<0>(R.T, nc, present, esig, esep);
	} else {
		lbf = log_bf_ref<0>(R.T, nc, present, sig, sep, R.sep_f32 != 0);
	}
	lbf_out[row] = lbf;
	const unsigned smask = present >> 1;
	post_out[row] = posterior_ref(R.T->prior[smask], R.T->log10prior[smask], lbf);
}

__global__ void k_posterior(long long n, const double *__restrict__ prior, const double *__restrict__ lbf,
	double *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = posterior_ref(prior[i], log10(prior[i]), lbf[i]);
}

}  // namespace nwb

// ---- table all-gather over peer memory (nwb_gather_*) -------------------------------------------------------------
// Every rank places its own shard of the output table -- ncols columns of `rows` 8-byte values, where the row kernels
// wrote them -- straight into its final position in the gathered table of EVERY rank (dst[r]: the ranks' gather
// buffers, the peers' mapped through cudaIpc; the own one is an ordinary copy): stores over NVLink from the SMs, no
// staging, no unpacking afterwards.  Work item = (chunk of a column, peer), peers fastest so that all links are busy
// at any time.  The destination row offset is only 8-byte aligned: up to 15 scalar head elements per chunk bring it to a
// 128-byte line, from there on every warp store is four whole lines (8 GPUs, 8 x 590 MB: 6.49 ms; 7.42 ms with stores
// that were only 16-byte aligned -- partial lines cost NVLink packets).
struct PushArgs {
	const char *src;             // the shard: column k at src + k * src_stride
	long long src_stride;        // bytes
	const long long *counts;     // device: rows of every rank's shard (e.g. an all-gather of the contexts' row counts) ...
	long long counts_val[16];    // ... or, if counts is null, the same by value
	long long cap_rows;          // rows the gathered table holds
	long long chunk;             // elements per work item (a multiple of 16); 0 = sized by the kernel
	int ncols, world, rank;
	char *dst[16];               // the ranks' gathered tables (this set)
	long long dst_stride;        // bytes between the gathered table's columns
	const long long *abort;      // device word, non-zero = the count exchange before this kernel failed (k_peer_sync); may be null
};

constexpr int PUSH_THREADS = 512;

template <int ST>
__device__ __forceinline__ void push_store(double2 *p, double2 v)
{
	if (ST == 1) __stcs(p, v);
	else if (ST == 2) __stwt(p, v);
	else *p = v;
}

// ST: flavour of the 16-byte store (0 plain, 1 streaming, 2 write-through; 8 GPUs: 6.49 / 6.37 / 6.38 ms, tools/bench_push.py)
template <int ST>
__global__ void __launch_bounds__(PUSH_THREADS) k_table_push(PushArgs A)
{
	long long rows = 0, off = 0, total = 0;
	for (int r = 0; r < A.world; r++) {
		const long long c = A.counts ? __ldg(A.counts + r) : A.counts_val[r];
		if (r < A.rank) off += c;
		if (r == A.rank) rows = c;
		total += c;
	}
	if (total > A.cap_rows || (A.abort && __ldg(A.abort) != 0)) return;   // the host sees the same counts / flag and reports it
	const long long dst_off = off * 8;
	// work items of up to 512 KB (long runs per link: 6.28 ms instead of 6.49 ms for 8 x 590 MB), smaller for a small table so
	// that every block still gets a few
	long long chunk = A.chunk;
	if (chunk == 0) chunk = max(2048LL, min(65536LL, (rows * A.ncols * A.world / (4LL * gridDim.x) + 15) / 16 * 16));
	const long long per_col = (rows + chunk - 1) / chunk;
	const long long items = per_col * A.ncols * A.world;
	for (long long id = blockIdx.x; id < items; id += gridDim.x) {
		const int peer = (int) ((id + A.rank) % A.world);
		const long long t = id / A.world;
		const int col = (int) (t % A.ncols);
		const long long first = (t / A.ncols) * chunk;
		const long long n = min(chunk, rows - first);
		const double *__restrict__ s = (const double *) (A.src + (long long) col * A.src_stride) + first;
		double *__restrict__ d = (double *) (A.dst[peer] + (long long) col * A.dst_stride + dst_off) + first;
		// up to 15 leading elements bring the destination to a 128-byte line: from there every warp store is four whole lines
		const int head = (int) min(n, (long long) ((16 - (int) (((unsigned long long) d >> 3) & 15)) & 15));
		if ((int) threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
		const long long body = (n - head) >> 1;
		const double *sb = s + head;
		double2 *db = (double2 *) (d + head);
		if ((((unsigned long long) sb) & 15) == 0) {   // source and destination equally aligned: 16-byte loads
			const double2 *sb2 = (const double2 *) sb;
#pragma unroll 4
			for (long long i = threadIdx.x; i < body; i += PUSH_THREADS) push_store<ST>(db + i, __ldcs(sb2 + i));
		} else {
#pragma unroll 4
			for (long long i = threadIdx.x; i < body; i += PUSH_THREADS) {
				double2 v;
				v.x = __ldcs(sb + 2 * i);
				v.y = __ldcs(sb + 2 * i + 1);
				push_store<ST>(db + i, v);
			}
		}
		if (threadIdx.x == 32 && ((n - head) & 1)) d[n - 1] = s[n - 1];
	}
}
