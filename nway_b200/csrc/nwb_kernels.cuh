// Kernels of the nway match path (sm_100a, fp64 CUDA cores -- there is no dense contraction on this path).
//
// Pipeline (one shard of primaries per context):
//   k_prim_prep     per primary: lon, sin/cos(lat), search box           -> P_* arrays           (K0)
//   k_prim_cells    count / fill: primary -> every grid cell its box overlaps                    (K0)
//   k_pairs         stream a secondary catalogue once: cell -> box test -> exact separation,
//                   warp-aggregated append of (primary, secondary, sep)                          (K1)
//   k_scatter       pair records -> per-primary segments                                         (K1)
//   k_sort_lists    per primary: rank-sort the segment by secondary index                        (K1)
//   k_count_rows    N >= 3: secondary-secondary separations + number of valid tuples             (K2)
//   k_rows          per primary: enumerate tuples in lexicographic order, score, write columns;
//                   optionally fused with the group normalisation                                (K2[+K3])
//   k_correct_cli   nway.py:366-421                                                              (K2b)
//   k_final         bias lookup, p_single, per-primary log-sum-exp, p_any, p_i, match_flag       (K3)
#pragma once
#include "nwb_device.cuh"

namespace nwb {

// ---------------------------------------------------------------------------------------------------------
// grid over (ra, dec): declination bands of height h, each cut into nra[b] cells along ra
// ---------------------------------------------------------------------------------------------------------
struct Grid {
	double dec_lo, inv_h;
	double ra_org, ra_span;
	int nbands;
	int full_circle;
	const int *nra;         // [nbands]
	const int *base;        // [nbands] first cell of the band
	const double *inv_w;    // [nbands] cells per degree of ra
	long long ncells;
};

struct Entry {   // 32 bytes: one primary as seen from one cell
	double ra_n, dec, dra;
	int p, pad;
};

struct PairRec {   // 16 bytes
	int p, s;
	double sep;
};

__device__ __forceinline__ double wrap360(double x)
{
	double y = x - 360.0 * floor(x / 360.0);
	return (y >= 360.0 || y < 0.0) ? 0.0 : y;
}

__device__ __forceinline__ int band_of(const Grid &G, double dec)
{
	double t = floor((dec - G.dec_lo) * G.inv_h);
	if (!(t >= 0.0)) return -1;
	if (t >= (double) G.nbands) return G.nbands;
	return (int) t;
}

__device__ __forceinline__ int racell_of(const Grid &G, int b, double x /* wrap360(ra - ra_org) */)
{
	int n = G.nra[b];
	int i = (int) (x * G.inv_w[b]);
	return i >= n ? n - 1 : (i < 0 ? 0 : i);
}

// ---------------------------------------------------------------------------------------------------------
// K0: primaries
// ---------------------------------------------------------------------------------------------------------
struct PrimArrays {
	double *lon, *slat, *clat;   // for the exact formula
	double *ra_n, *dec, *dra;    // search box (degrees); dra >= 180 means "all ra"
};

// box margins: rb (deg) is the search radius inflated by 1e-9 relative + 1e-12, so that rounding in the
// box test can never reject a pair the exact formula would accept.
__global__ void k_prim_prep(int np, long long first, const double *__restrict__ ra, const double *__restrict__ dec,
	double rb, PrimArrays P, double *__restrict__ red /* [gridDim.x][6] */)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	double v[6] = {1e300, -1e300, 1e300, -1e300, 1e300, -1e300};   // dec min/max, A lo/hi, B lo/hi
	if (i < np) {
		double r = ra[first + i], d = dec[first + i];
		double lat = deg2rad_ref(d);
		double sl, cl;
		sincos(lat, &sl, &cl);
		P.lon[i] = deg2rad_ref(r);
		P.slat[i] = sl;
		P.clat[i] = cl;
		double rn = wrap360(r);
		double dra;
		if (fabs(d) + rb >= 89.999) {
			dra = 360.0;
		} else {
			double s = sin(rb / 180 * NWB_PI) / cos((fabs(d)) / 180 * NWB_PI);
			dra = s >= 1.0 ? 360.0 : asin(s) * 180 / NWB_PI * (1 + 1e-9) + 1e-12;
		}
		P.ra_n[i] = rn;
		P.dec[i] = d;
		P.dra[i] = dra;
		double rn_b = wrap360(rn + 180.0);
		v[0] = d; v[1] = d;
		v[2] = rn - dra; v[3] = rn + dra;
		v[4] = rn_b - dra; v[5] = rn_b + dra;
	}
	// block reduce
	__shared__ double sm[6][32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 6; k++) {
		double x = v[k];
		for (int o = 16; o > 0; o >>= 1) {
			double y = __shfl_xor_sync(NWB_FULL, x, o);
			x = (k & 1) ? fmax(x, y) : fmin(x, y);
		}
		if (lane == 0) sm[k][w] = x;
	}
	__syncthreads();
	if (w == 0) {
		int nw = blockDim.x >> 5;
#pragma unroll
		for (int k = 0; k < 6; k++) {
			double x = lane < nw ? sm[k][lane] : ((k & 1) ? -1e300 : 1e300);
			for (int o = 16; o > 0; o >>= 1) {
				double y = __shfl_xor_sync(NWB_FULL, x, o);
				x = (k & 1) ? fmax(x, y) : fmin(x, y);
			}
			if (lane == 0) red[blockIdx.x * 6 + k] = x;
		}
	}
}

__global__ void k_reduce6(int nblocks, const double *__restrict__ red, double *__restrict__ out)
{
	int lane = threadIdx.x;
	for (int k = 0; k < 6; k++) {
		double x = (k & 1) ? -1e300 : 1e300;
		for (int i = lane; i < nblocks; i += 32) {
			double y = red[i * 6 + k];
			x = (k & 1) ? fmax(x, y) : fmin(x, y);
		}
		for (int o = 16; o > 0; o >>= 1) {
			double y = __shfl_xor_sync(NWB_FULL, x, o);
			x = (k & 1) ? fmax(x, y) : fmin(x, y);
		}
		if (lane == 0) out[k] = x;
	}
}

// FILL = false: cellcnt[cell] += 1 for every cell the primary's (slightly inflated) box overlaps.
// FILL = true : write the entry at cstart[cell] + slot.
template <bool FILL>
__global__ void k_prim_cells(int np, Grid G, PrimArrays P, double rb_ins, double dra_eps,
	int *__restrict__ cellcnt, const int *__restrict__ cstart, Entry *__restrict__ entries)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= np) return;
	double d = P.dec[i], rn = P.ra_n[i], dra = P.dra[i];
	int b0 = band_of(G, d - rb_ins), b1 = band_of(G, d + rb_ins);
	b0 = max(b0, 0);
	b1 = min(b1, G.nbands - 1);
	Entry en;
	en.ra_n = rn; en.dec = d; en.dra = dra; en.p = i; en.pad = 0;
	double di = dra + dra_eps;
	for (int b = b0; b <= b1; b++) {
		int n = G.nra[b];
		int i0, cnt;
		double cellw = G.ra_span / n;
		if (G.full_circle) {
			if (2 * di + 2 * cellw >= 360.0) { i0 = 0; cnt = n; }
			else {
				i0 = racell_of(G, b, wrap360(rn - di - G.ra_org));
				int i1 = racell_of(G, b, wrap360(rn + di - G.ra_org));
				cnt = (i1 - i0 + n) % n + 1;
			}
		} else {
			// the grid's ra window was built from min(rn - dra) .. max(rn + dra) with a margin: no wrap inside
			double x0 = wrap360(rn - G.ra_org) - di, x1 = wrap360(rn - G.ra_org) + di;
			i0 = racell_of(G, b, fmax(x0, 0.0));
			int i1 = racell_of(G, b, fmin(x1, G.ra_span));
			cnt = i1 - i0 + 1;
		}
		int cb = G.base[b];
		for (int k = 0; k < cnt; k++) {
			int cell = cb + (i0 + k) % n;
			if (FILL) {
				int slot = atomicAdd(&cellcnt[cell], 1);
				entries[cstart[cell] + slot] = en;
			} else {
				atomicAdd(&cellcnt[cell], 1);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// K1: stream one secondary catalogue
// ---------------------------------------------------------------------------------------------------------
// One thread per secondary source, coalesced streaming loads of (ra, dec) -- 16 algorithmic bytes per source,
// read exactly once.  The cell lookup and the primary records come from L2-resident tables.  Matches are
// appended with one atomicAdd per warp.
__global__ void __launch_bounds__(256)
k_pairs(long long n, const double *__restrict__ ra, const double *__restrict__ dec, Grid G,
	const int *__restrict__ cstart, const Entry *__restrict__ entries, PrimArrays P, double rb, double radius,
	PairRec *__restrict__ out, unsigned long long cap, unsigned long long *__restrict__ out_count,
	int *__restrict__ cnt)
{
	const int lane = threadIdx.x & 31;
	long long stride = (long long) gridDim.x * blockDim.x;
	long long nround = (n + 31) / 32 * 32;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
		int e0 = 0, e1 = 0;
		double r = 0, d = 0, rn = 0;
		if (i < n) {
			r = __ldcs(ra + i);
			d = __ldcs(dec + i);
			int b = band_of(G, d);
			if (b >= 0 && b < G.nbands) {
				double x = wrap360(r - G.ra_org);
				if (G.full_circle || x <= G.ra_span) {
					int cell = G.base[b] + racell_of(G, b, x);
					e0 = cstart[cell];
					e1 = cstart[cell + 1];
					rn = wrap360(r);
				}
			}
		}
		bool have_trig = false;
		double lon2 = 0, slat2 = 0, clat2 = 0;
		for (int e = e0; __any_sync(NWB_FULL, e < e1); e++) {
			bool hit = false;
			int p = 0;
			double sep = 0;
			if (e < e1) {
				Entry en = entries[e];
				if (fabs(d - en.dec) <= rb) {
					double dr = rn - en.ra_n;
					if (dr > 180.0) dr -= 360.0;
					else if (dr < -180.0) dr += 360.0;
					if (fabs(dr) <= en.dra) {
						if (!have_trig) {
							sincos(deg2rad_ref(d), &slat2, &clat2);
							lon2 = deg2rad_ref(r);
							have_trig = true;
						}
						p = en.p;
						sep = sep_arcsec_ref(P.lon[p], P.slat[p], P.clat[p], lon2, slat2, clat2);
						hit = sep < radius;
					}
				}
			}
			unsigned m = __ballot_sync(NWB_FULL, hit);
			if (m) {
				unsigned long long basepos = 0;
				int leader = __ffs(m) - 1;
				if (lane == leader) basepos = atomicAdd(out_count, (unsigned long long) __popc(m));
				basepos = __shfl_sync(NWB_FULL, basepos, leader);
				if (hit) {
					unsigned long long pos = basepos + __popc(m & ((1u << lane) - 1));
					if (pos < cap) {
						PairRec rec;
						rec.p = p; rec.s = (int) i; rec.sep = sep;
						out[pos] = rec;
						atomicAdd(&cnt[p], 1);
					}
				}
			}
		}
	}
}

__global__ void k_scatter(long long npairs, const PairRec *__restrict__ recs, const long long *__restrict__ seg_off,
	int *__restrict__ fill, int *__restrict__ seg_s, double *__restrict__ seg_sep)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= npairs) return;
	PairRec r = recs[i];
	long long pos = seg_off[r.p] + atomicAdd(&fill[r.p], 1);
	seg_s[pos] = r.s;
	seg_sep[pos] = r.sep;
}

// One warp per primary: out-of-place rank sort of its segment by secondary index (the reference's sorted()
// of every bucket list, fastskymatch.py:181).  WITH_TRIG additionally stores lon, sin/cos(lat) of the
// secondary for the secondary-secondary separations of N >= 3.
template <bool WITH_TRIG>
__global__ void k_sort_lists(int np, const long long *__restrict__ seg_off, const int *__restrict__ seg_s,
	const double *__restrict__ seg_sep, int *__restrict__ L_s, double *__restrict__ L_sep,
	const double *__restrict__ ra, const double *__restrict__ dec, double *__restrict__ L_lon,
	double *__restrict__ L_slat, double *__restrict__ L_clat)
{
	int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int p = warp; p < np; p += nwarps) {
		long long lo = seg_off[p];
		int n = (int) (seg_off[p + 1] - lo);
		for (int e = lane; e < n; e += 32) {
			int s = seg_s[lo + e];
			int rank = 0;
			for (int f = 0; f < n; f++) rank += seg_s[lo + f] < s;
			L_s[lo + rank] = s;
			L_sep[lo + rank] = seg_sep[lo + e];
			if (WITH_TRIG) {
				double sl, cl;
				sincos(deg2rad_ref(dec[s]), &sl, &cl);
				L_lon[lo + rank] = deg2rad_ref(ra[s]);
				L_slat[lo + rank] = sl;
				L_clat[lo + rank] = cl;
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------
// K2 / K3 parameter block
// ---------------------------------------------------------------------------------------------------------
struct Lists {
	const long long *off[MAXC];   // [c] for c >= 1: segment offsets, np+1
	const int *s[MAXC];
	const double *sep[MAXC];
	const double *lon[MAXC], *slat[MAXC], *clat[MAXC];
};

struct Columns {
	long long *idx[MAXC];
	double *sep[MAXP];
	double *sepmax;
	long long *ncat;
	double *lbf_u, *lbf, *dist_post;
	double *bias[MAXM];
	double *p_single;
	long long *flag;
	double *p_any, *p_i;
};

struct RowParams {
	int ncat, nmag, np;
	long long first;                 // global index of local primary 0
	double radius, ratio_secondary;
	const double *err[MAXC];         // sigma columns (circular)
	const ConstTables *T;
	Lists L;
	Columns C;
	const long long *row_off;        // [np+1]
	const long long *mat_off;        // [np+1] (N >= 3)
	double *mat;                     // secondary-secondary separations
};

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
__global__ void k_mat_sizes(int np, int ncat, Lists L, long long *__restrict__ sizes)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= np) return;
	long long tot = 0;
	for (int a = 1; a < ncat; a++)
		for (int b = a + 1; b < ncat; b++)
			tot += (L.off[a][p + 1] - L.off[a][p]) * (L.off[b][p + 1] - L.off[b][p]);
	sizes[p] = tot;
}

// N >= 3: one warp per primary computes every secondary-secondary separation once and counts the tuples
// that survive the pairwise radius filter (__init__.py:166,180).
template <int NC>
__global__ void k_count_rows(RowParams R, long long *__restrict__ rows)
{
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
						ok = ok && (s < R.radius);
					}
					bo += (long long) nl[a] * nl[b];
				}
			}
			count += ok;
		}
		count = warp_sum_ll(count);
		if (lane == 0) rows[p] = count;
	}
}

// N == 2: rows per primary = matches + 1
__global__ void k_rows_per_primary_2(int np, const long long *__restrict__ off, long long *__restrict__ rows)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < np) rows[p] = off[p + 1] - off[p] + 1;
}

// ---------------------------------------------------------------------------------------------------------
// per-row finalisation pieces shared by k_rows<FUSE> and k_final
// ---------------------------------------------------------------------------------------------------------
// bias lookup for one row: returns sum of weights in the reference's order ((0 + w1) + w2 ...), writes bias cols
__device__ __forceinline__ double row_bias(const RowParams &R, long long row, const long long *sidx /* [ncat] */)
{
	double wsum = 0.0;
	for (int j = 0; j < R.nmag; j++) {
		const MagTable &M = R.T->mag[j];
		long long s = sidx[M.cat];
		double m = -99.0;
		if (s >= 0) {
			m = M.mag[s];
			if (!isfinite(m)) m = -99.0;
		}
		double b;
		double w = mag_weight(M, m, b);
		R.C.bias[j][row] = b;
		wsum = wsum + w;
	}
	return wsum;
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
		const long long rbase = R.row_off[p];
		const double *mat = NC > 2 ? R.mat + R.mat_off[p] : nullptr;
		const long long gp = R.first + p;
		const double sig0 = R.err[0][gp];
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
								ok = ok && (s < R.radius);
							}
							sep[pair_index(a, b, NC)] = s;
							bo += (long long) nl[a] * nl[b];
						}
					}
				}
			}
			unsigned m = __ballot_sync(NWB_FULL, ok);
			if (ok) {
				long long row = rbase + written + __popc(m & ((1u << lane) - 1));
				double smax = 0.0;
#pragma unroll
				for (int c = 0; c < NC; c++) R.C.idx[c][row] = sidx[c];
#pragma unroll
				for (int k = 0; k < NC * (NC - 1) / 2; k++) {
					R.C.sep[k][row] = sep[k];
					if (sep[k] > smax) smax = sep[k];   // NaN compares false: nanmax
				}
				R.C.sepmax[row] = smax;
				R.C.ncat[row] = __popc(present);
#pragma unroll
				for (int c = 1; c < NC; c++)
					if (present >> c & 1u) sig[c] = R.err[c][sidx[c]];
				double lbf = log_bf_ref<NC>(T, NC, present, sig, sep);
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

// K3 (unfused): one warp per primary
__global__ void __launch_bounds__(256)
k_final(RowParams R)
{
	const int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	const ConstTables *__restrict__ T = R.T;
	for (int p = warp; p < R.np; p += nwarps) {
		long long r0 = R.row_off[p], n = R.row_off[p + 1] - r0;
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

// K2b: the CLI's unrelated-association correction (nway.py:366-421), one warp per primary.
// For every set M of >= 2 secondary catalogues: best(M) = max(0, max over rows j with ncat_j > 2 of
// log_bf(sub-association M & present_j of row j) + log10(nu[A0]/prod nu_plus[A])), taken when the
// intersection has >= 2 members; every row whose ABSENT set is exactly M gets best(M) added.
__global__ void k_correct_cli(RowParams R)
{
	const int lane = threadIdx.x & 31;
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	const ConstTables *__restrict__ T = R.T;
	const int nc = R.ncat;
	const unsigned all = (1u << (nc - 1)) - 1;   // over secondaries, bit c-1
	for (int p = warp; p < R.np; p += nwarps) {
		long long r0 = R.row_off[p], n = R.row_off[p + 1] - r0;
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
				for (int c = 1; c < nc; c++)
					if (A >> (c - 1) & 1u) sig[c] = R.err[c][R.C.idx[c][row]];
				for (int a = 1; a < nc; a++)
					for (int b = a + 1; b < nc; b++)
						if ((A >> (a - 1) & 1u) && (A >> (b - 1) & 1u))
							sep[pair_index(a, b, nc)] = R.C.sep[pair_index(a, b, nc)][row];
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
	sincos(deg2rad_ref(dec1[i]), &s1, &c1);
	sincos(deg2rad_ref(dec2[i]), &s2, &c2);
	sincos(deg2rad_ref(ra2[i]) - deg2rad_ref(ra1[i]), &sd, &cd);
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

__global__ void k_posterior(long long n, const double *__restrict__ prior, const double *__restrict__ lbf,
	double *__restrict__ out)
{
	long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = posterior_ref(prior[i], log10(prior[i]), lbf[i]);
}

}  // namespace nwb
