// What a row of the output table is made of, as plain per-row functions: the pair store k_pairs fills, the parameter
// block of the row kernels, the magnitude-prior look-up of a row and the scoring of one row of a two-catalogue match
// (rows2_write).  Kept apart from the kernels so that tests/emu/rows_emu.cpp can run exactly these functions on the
// host and compare whole tables with an independent CPU computation.
#pragma once
#include "nwb_device.cuh"
#include "nwb_grid.cuh"

namespace nwb {

struct Slot16 {   // 16 bytes: one match of a primary
	int s, pad;
	double sep;
};

struct SpillRec {   // 24 bytes: a match that did not fit the primary's slots
	int p, slot, s, pad;
	double sep;
};

// where the matches of one secondary catalogue live: C slots per primary, in arrival order (cnt[p] of them are
// valid), plus -- only if some primary overflowed -- per-primary spill segments.
struct PairStore {
	const Slot16 *base;
	int C;
	const int *cnt;
	const long long *spill_off;   // nullptr when nothing spilled
	const Slot16 *spill;
};

__device__ __forceinline__ Slot16 store_get(const PairStore &S, int p, int e)
{
	if (e < S.C) return S.base[(size_t) p * S.C + e];
	return S.spill[S.spill_off[p] + (e - S.C)];
}

// ---------------------------------------------------------------------------------------------------------
// K2 / K3 parameter block
// ---------------------------------------------------------------------------------------------------------
struct Lists {
	const long long *off[MAXC];   // [c] for c >= 1: segment offsets, np+1
	const int *s[MAXC];
	const double *sep[MAXC];
	const double *lon[MAXC], *slat[MAXC], *clat[MAXC];
	const long long *ij[MAXC];    // NWB_COMPAT_FLAT_HASH: the reference's flat-sky cell of every list entry (flat_hash_pack)
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
	double pair_radius[MAXP];        // per catalogue pair: min(radius, --prefilter-pair radius), arcsec (fastskymatch.py:184-208)
	const double *err[MAXC];         // sigma columns (circular); elliptical: sigma_x | sigma_y | rho, each n[c] long
	int ell;                         // elliptical mode (nway.py:346-354): every catalogue carries a triple
	int sep_f32;                     // nway.py compatibility: separations / offsets pass through float32
	FlatHash flat;                   // .err > 0: NWB_COMPAT_FLAT_HASH in force; the reference's bucket size in degrees (radius / 60. / 60)
	int small_t;                     // primaries with at most this many candidate tuples are handled by k_rows_small (0 = off)
	long long n[MAXC];               // catalogue sizes (stride of the error triple)
	const double *ra[MAXC], *dec[MAXC];
	const ConstTables *T;
	Lists L;
	Columns C;
	const long long *row_off;        // [np+1]
	const long long *mat_off;        // [np+1] (N >= 3)
	double *mat;                     // secondary-secondary separations
	PairStore S1;                    // N == 2: the matches of catalogue 1, unsorted, straight from k_pairs
	int err1_const;                  // catalogue 1 carries one positional error for all sources:
	double err1_value;               //   no per-row gather (one random 32-byte sector per row saved)
	// speculative launch (no host sync between K1 and K2): the kernel runs only if the status words written by
	// k_collect_status say that everything it depends on is valid and fits; the host checks the same words after
	// its single sync and re-runs the non-speculative path otherwise
	const long long *guard;
	long long max_rows, entries_cap;
	// N >= 3 / elliptical, speculative launch: the stage runs only if the word the gate kernel (k_spec_gate) wrote for it
	// says that everything it depends on fits the buffers of the previous match
	const int *gate;
};

__device__ __forceinline__ bool gate_open(const int *gate) { return !gate || *reinterpret_cast<const volatile int *>(gate) != 0; }

__device__ __forceinline__ bool guard_ok(const RowParams &R)
{
	if (!R.guard) return true;
	return R.guard[9] == 0 && R.guard[1] == 0 && R.guard[8] <= R.max_rows && R.guard[0] <= R.entries_cap;
}

// NWB_COMPAT_FLAT_HASH: does the reference's flat-sky hash hold this tuple?  It does iff one bucket received all of its
// present members, i.e. their cells span at most one step in i and in j (fastskymatch.py:125-132, 164-181).
// wp = cell of the primary, dg[c] = 0 (absent) or 1 + position in catalogue c's list, lo[c] = start of that list.
template <int NC>
__device__ __forceinline__ bool tuple_in_one_bucket(const RowParams &R, long long wp, const int *dg, const long long *lo)
{
	int imin = flat_hash_i(wp), imax = imin, jmin = flat_hash_j(wp), jmax = jmin;
#pragma unroll
	for (int c = 1; c < NC; c++)
		if (dg[c] > 0) {
			const long long w = R.L.ij[c][lo[c] + dg[c] - 1];
			const int i = flat_hash_i(w), j = flat_hash_j(w);
			imin = min(imin, i); imax = max(imax, i);
			jmin = min(jmin, j); jmax = max(jmax, j);
		}
	return imax - imin <= 1 && jmax - jmin <= 1;
}

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

// per-lane memo of the error-dependent terms of the 2-catalogue Bayes factor: catalogues very often carry one
// positional error for all sources (or a handful of values), and then w, log(w), log(w0 + w1) need not be
// recomputed row after row.  Same expressions, same bits -- just not evaluated twice for equal inputs.
struct R2Memo {
	double s1 = -1.0, w1 = 0.0, lw1 = 0.0;      // key s1
	double kw0 = -1.0, kw1 = -1.0, wsum = 0.0, lwsum = 0.0, rwsum = 0.0;   // key (w0, w1); rwsum = RN(1 / wsum)
};

// SHARE (fused, no magnitude priors): dist_post and p_single are not written here but by the normalisation, which
// gets 10^(-v) from the exponential it evaluates anyway
template <bool FUSE, bool SHARE>
__device__ __forceinline__ void rows2_write(const RowParams &R, const ConstTables *__restrict__ T, long long row,
	long long gp, long long sidx1, double sep, double w0, double lw0, R2Memo &memo, double &v_out)
{
	const bool present = sidx1 >= 0;
	if (R.sep_f32) sep = (double) (float) sep;
	R.C.idx[0][row] = gp;
	R.C.idx[1][row] = sidx1;
	R.C.sep[0][row] = present ? sep : nan("");
	R.C.sepmax[row] = present ? sep : 0.0;
	R.C.ncat[row] = present ? 2 : 1;
	double lbf = 0.0;
	if (present) {
		// bayesdistance.py:64-86 for n = 2, same operation order as log_bf_ref<2>
		double s1 = R.err1_const ? R.err1_value : R.err[1][sidx1];
		if (s1 != memo.s1) {
			memo.s1 = s1;
			memo.w1 = 1.0 / (s1 * s1);
			memo.lw1 = log(memo.w1);
		}
		double w1 = memo.w1;
		if (w0 != memo.kw0 || w1 != memo.kw1) {
			memo.kw0 = w0; memo.kw1 = w1;
			memo.wsum = w0 + w1;
			memo.lwsum = log(memo.wsum);
			memo.rwsum = 1.0 / memo.wsum;
		}
		double wsum = memo.wsum;
		double slog = lw0 + memo.lw1 - memo.lwsum;
		double q = w0 * w1 * (R.sep_f32 ? (double) __fmul_rn((float) sep, (float) sep) : sep * sep);
		// -q / 2 / wsum: the halving is exact, the quotient comes from the memoised reciprocal (quotient_by_reciprocal)
		const double exponent = quotient_by_reciprocal(-q / 2, wsum, memo.rwsum);
		lbf = (T->norm[2] + slog + exponent) * T->log10e;
	}
	unsigned smask = present ? 1u : 0u;
	double prior = T->prior[smask], l10p = T->log10prior[smask];
	R.C.lbf_u[row] = lbf;
	R.C.lbf[row] = lbf;
	if (SHARE) {
		v_out = lbf + l10p;
		return;
	}
	double post = 1. / (1 + (1 - prior) * nwb_exp10(-lbf - l10p));   // bayesdistance.py:32
	R.C.dist_post[row] = post;
	if (FUSE) {
		double total = lbf;
		double ps = post;
		if (R.nmag > 0) {
			long long sidx[2] = {gp, sidx1};
			total = lbf + row_bias(R, row, sidx);
			ps = posterior_ref(prior, l10p, total);
		}
		R.C.p_single[row] = ps;
		v_out = total + l10p;
	}
}

}  // namespace nwb
