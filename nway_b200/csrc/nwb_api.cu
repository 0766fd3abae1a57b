// C ABI of the B200-native nway match path: context, buffers, stage orchestration.  See include/nwayb200.h.
#include "../../include/nwayb200.h"
#include "nwb_kernels.cuh"
#include "nwb_peer.cuh"
#include "nwb_grid_host.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace nwb;

#ifndef NWB_R2_SHARE
#define NWB_R2_SHARE 1
#endif
#ifndef NWB_CELL_FACTOR
#define NWB_CELL_FACTOR 1.0   // smallest cell edge in units of the search radius (experiment knob; >= 1)
#endif

namespace {

std::string g_create_error;

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
};

struct CatSlot {
	bool set = false;
	int64_t n = 0;
	const double *ra = nullptr, *dec = nullptr, *err = nullptr, *mags = nullptr;
	int err_kind = NWB_ERR_CIRCULAR, m = 0;
	double area = 0;
	DevBuf own;   // one allocation holding ra|dec|err|mags when copied from the host
	bool err_const = false;   // every source has the same (positive, circular) error
	double err_value = 0;
	bool bounds_known = false;   // NWB_COMPAT_FLAT_HASH: smallest / largest ra, largest |dec|, any NaN (k_radec_bounds)
	double ra_min = 0, ra_max = 0, absdec_max = 0;
	bool has_nan = false;
};

struct HostMagHist {
	bool set = false;
	int nbins = 0;
	double edges[MAXB + 1], weight[MAXB], bias[MAXB];
};

}  // namespace

struct nwb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr, own_stream = nullptr;
	std::string err;
	int ncat = 0;
	CatSlot cat[MAXC];
	HostMagHist hist[MAXC][MAXM];
	bool params_set = false, tables_set = false;
	double radius = 0, ratio_secondary = 0.5;
	double pc[MAXC];
	int unrelated_mode = NWB_UNRELATED_API;
	int compat = 0;
	// shard mode (nwb_shard_*): streaming split by secondary rows, matches scattered into the owners' exchange buffers
	struct Shard {
		bool on = false, connected = false;
		int rank = 0, world = 1;
		int64_t block = 0;              // primaries per rank (the last rank may own fewer)
		DevBuf xch, d_peers;            // this rank's exchange buffer; device array of every rank's buffer base
		void *peer[16] = {nullptr};     // host copy; [rank] = xch.p, the others opened through cudaIpc
		size_t off_cnt[MAXC] = {0}, off_slot[MAXC] = {0}, off_spill[MAXC] = {0}, off_spillcnt[MAXC] = {0};
		size_t zero_bytes = 0, bytes = 0;   // counters + match counts come first: zeroed before every match
		size_t set_bytes = 0;               // TWO such sets: match e uses set e % 2 while set (e + 1) % 2 is zeroed for the next one
		long long epoch = 0;
		int C[MAXC] = {0};
		unsigned long long spill_cap = 0;
		size_t off_bar = 0;                 // behind the two sets: 16 barrier words (k_peer_sync), tags only grow
		unsigned long long bar_tag = 0;
	} shard;
	// table all-gather over peer memory (nwb_gather_*)
	struct Gather {
		bool on = false, connected = false;
		int rank = 0, world = 1;
		DevBuf buf;                     // TWO sets of set_bytes: push e goes to set e % 2 of every rank
		void *peer[16] = {nullptr};     // [rank] = buf.p, the others opened through cudaIpc
		size_t set_bytes = 0, stride = 0;   // bytes per set / between the columns of a gathered table
		int64_t cap_rows = 0;
		int ncols = 0;
		long long epoch = 0;
		cudaStream_t lane[16] = {nullptr};   // copy-engine variant: one stream per destination
		cudaEvent_t fork = nullptr, join[16] = {nullptr};
		// engine 2 (no collective library at all): the buffer starts with a header of flag words -- row counts, their tags,
		// completion tags (k_peer_sync) -- the sets follow at GATHER_HEADER
		DevBuf d_counts;                // [17]: the ranks' row counts as collected by k_peer_sync, [16] = abort
		long long *h_words = nullptr;   // pinned, mapped: the same for the host ([16] = error)
	} gather;
	// N >= 3 / elliptical: what the previous match of this shape left behind -- buffer capacities and launch decisions -- so
	// that the next one can enqueue its whole pipeline behind device-side gates (k_spec_gate) without host round trips
	struct GenCaps {
		bool valid = false;
		int nc = 0, ell = 0;
		int64_t np = -1;
		long long cap_list[MAXC] = {0}, cap_mat = 0;
		int big_sort[MAXC] = {0}, any_big = 0;
	} gen;
	// what nwb_bench_skeleton needs of the last match: K1 arguments per secondary catalogue and the grid's buffers
	K1Args last_k1[MAXC];
	bool last_k1_dense = false;
	const int *last_etotal = nullptr;
	double flat_err = 0;             // > 0: this match applies the reference's flat-sky bucket predicate (NWB_COMPAT_FLAT_HASH)
	bool any_big = true;             // some primary has more candidate tuples than k_rows_small handles
	double prefilter[MAXP];          // per catalogue pair, arcsec; +inf = none
	bool prefilter_on = false;
	ConstTables tables;   // host copy
	int64_t first = 0, count = -1;

	// device scratch (grow-only)
	DevBuf d_tables, d_prim, d_red, d_bands, d_cellcnt, d_entries, d_cub, d_pairs, d_paircount;
	DevBuf d_cnt[MAXC], d_segoff[MAXC], d_seg_s[MAXC], d_seg_sep[MAXC], d_Ls[MAXC], d_Lsep[MAXC], d_Ltrig[MAXC];
	DevBuf d_rows, d_rowoff, d_matsz, d_matoff, d_mat, d_cols, d_cols2, d_keep, d_keeppos, d_misc;
	DevBuf d_spill, d_status, d_spilloff[MAXC], d_spillseg[MAXC], d_cells, d_worklist, d_surv;
	DevBuf d_hrow, d_hsrc, d_hout;   // automatic histograms: per-row / per-source scratch, compact sample
	int hist_cat = -1, hist_k = -1;  // the (catalogue, magnitude) whose 'possible' marks d_hsrc holds
	size_t entries_cap = 0;
	unsigned long long spill_cap = 0;
	long long *h_status = nullptr;   // pinned
	int k1_occ[16] = {0}, num_sms = 0, filter_occ = 0, skel_blocks = 0;   // resident blocks per SM of every k_pairs instantiation
	// grid geometry of the previous match, re-used when the primaries' bounding box and the radius are unchanged
	bool geom_valid = false;
	double geom_rb = 0;
	int64_t geom_np = -1, geom_first = -1;
	BoundsKey geom_key;
	Grid geom_G;
	double geom_occ = 1.0;
	int64_t cols_cap_rows = 0;
	int cols_cap_ncols = 0;
	bool timing_dirty = false, tables_dirty = true, shard_events_valid = false;
	int timing_ncat = 0;

	// result
	bool pending = false;            // nwb_match_async enqueued a match that nwb_match_wait has not collected yet
	int pending_fuse = 0;
	bool matched = false, finalized = false;
	int64_t nrows = 0, np = 0;
	int ncols = 0;
	int res_ncat = 0, res_nmag = 0;
	Columns cols;
	RowParams rp;

	cudaEvent_t ev[8];
	cudaEvent_t kev[2 * MAXC + 2];   // around each k_pairs launch, and around k_rows
	float ms[NWB_T_COUNT];
	int64_t launches = 0;
	int64_t stats[4] = {0, 0, 0, 0};
};

namespace {

int fail(nwb_ctx *c, int code, const std::string &msg)
{
	if (c) c->err = msg; else g_create_error = msg;
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(ctx, NWB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

#define LAUNCH(ctx, kernel, grid, block, ...) do { \
	kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
	(ctx)->launches++; \
	cudaError_t e_ = cudaGetLastError(); \
	if (e_ != cudaSuccess) return fail(ctx, NWB_ERR_CUDA, std::string(#kernel) + ": " + cudaGetErrorString(e_)); } while (0)

int ensure(nwb_ctx *ctx, DevBuf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) return NWB_OK;
	if (b.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
	size_t want = std::max<size_t>(bytes + bytes / 8, 256);
	cudaError_t e = cudaMalloc(&b.p, want);
	if (e != cudaSuccess) {
		cudaGetLastError();
		e = cudaMalloc(&b.p, std::max<size_t>(bytes, 256));
		want = std::max<size_t>(bytes, 256);
		if (e != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return fail(ctx, NWB_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed"); }
	}
	b.cap = want;
	return NWB_OK;
}

#define ENSURE(buf, bytes) do { int r_ = ensure(ctx, buf, bytes); if (r_) return r_; } while (0)

void release(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

struct CastLL {
	__host__ __device__ long long operator()(const int &x) const { return (long long) x; }
};

// exclusive prefix sum of n ints into n long longs (n includes a trailing 0 so out[n-1] is the total)
int scan_int_to_ll(nwb_ctx *ctx, const int *in, long long *out, int64_t n)
{
	cub::TransformInputIterator<long long, CastLL, const int *> it(in, CastLL());
	size_t bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int) n, ctx->stream));
	ENSURE(ctx->d_cub, bytes);
	CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, it, out, (int) n, ctx->stream));
	return NWB_OK;
}

// rows per primary of a two-catalogue match = matches + the no-counterpart row; element np is the trailing 0
struct RowsOfPrimary {
	const int *cnt;
	int np;
	__host__ __device__ long long operator()(const int &p) const { return p < np ? (long long) cnt[p] + 1 : 0ll; }
};

// row_off[0 .. np] (row_off[np] = R) straight from the match counters: no separate "rows per primary" pass
int scan_rows2(nwb_ctx *ctx, const int *cnt, long long *out, int64_t np)
{
	cub::CountingInputIterator<int> idx(0);
	cub::TransformInputIterator<long long, RowsOfPrimary, cub::CountingInputIterator<int>> it(idx, RowsOfPrimary{cnt, (int) np});
	size_t bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int) np + 1, ctx->stream));
	ENSURE(ctx->d_cub, bytes);
	CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, it, out, (int) np + 1, ctx->stream));
	return NWB_OK;
}

int scan_ll(nwb_ctx *ctx, const long long *in, long long *out, int64_t n)
{
	size_t bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int) n, ctx->stream));
	ENSURE(ctx->d_cub, bytes);
	CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, in, out, (int) n, ctx->stream));
	return NWB_OK;
}

int scan_int(nwb_ctx *ctx, const int *in, int *out, int64_t n)
{
	size_t bytes = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int) n, ctx->stream));
	ENSURE(ctx->d_cub, bytes);
	CU(cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, in, out, (int) n, ctx->stream));
	return NWB_OK;
}

// the scalar tables, with the reference's expressions (used when the caller did not supply its own)
void default_tables(nwb_ctx *ctx)
{
	ConstTables &T = ctx->tables;
	const int nc = ctx->ncat;
	const double log_arcsec2rad = std::log(3600 * 180 / M_PI);          // bayesdistance.py:15
	for (int n = 0; n <= MAXC; n++)
		T.norm[n] = (n - 1) * std::log(2.0) + 2 * (n - 1) * log_arcsec2rad;   // bayesdistance.py:76
	T.log10e = std::log10(M_E);
	const double area_total = 4 * M_PI * ((180 / M_PI) * (180 / M_PI));  // __init__.py:205
	double nu[MAXC], nup[MAXC];
	for (int c = 0; c < nc; c++) {
		double n = (double) ctx->cat[c].n, area = ctx->cat[c].area * 1.0;
		nu[c] = n / area * area_total;
		nup[c] = (n + 1) / area * area_total;
	}
	nup[0] = nu[0];
	for (unsigned mask = 0; mask < (1u << (nc - 1)); mask++) {
		double pcprod = ctx->pc[0], nuprod = nup[0];     // numpy.prod: left to right, index 0 included
		for (int c = 1; c < nc; c++)
			if (mask >> (c - 1) & 1u) { pcprod *= ctx->pc[c]; nuprod *= nup[c]; }
		T.prior[mask] = nu[0] * pcprod / nuprod;         // __init__.py:254
		T.log10prior[mask] = std::log10(T.prior[mask]);
		// sub-association prior of the CLI correction: nu[A0] / prod(nu_plus[A])   (nway.py:395)
		double sub = 1.0;
		int a0 = -1;
		bool firstp = true;
		double prod = 1.0;
		for (int c = 1; c < nc; c++)
			if (mask >> (c - 1) & 1u) {
				if (a0 < 0) a0 = c;
				if (firstp) { prod = nup[c]; firstp = false; } else prod *= nup[c];
			}
		if (a0 >= 0) sub = nu[a0] / prod;
		T.sub_log10prior[mask] = std::log10(sub);
	}
}

int upload_tables(nwb_ctx *ctx)
{
	if (!ctx->tables_dirty && ctx->d_tables.p) return NWB_OK;   // nothing changed since the last upload
	ConstTables &T = ctx->tables;
	int nmag = 0;
	for (int c = 0; c < ctx->ncat; c++) {
		for (int k = 0; k < ctx->cat[c].m; k++) {
			if (nmag >= MAXM) return fail(ctx, NWB_ERR_ARG, "too many magnitude columns");
			MagTable &M = T.mag[nmag];
			const HostMagHist &H = ctx->hist[c][k];
			M.cat = c;
			M.mag = ctx->cat[c].mags + (int64_t) k * ctx->cat[c].n;
			if (H.set) {
				M.nbins = H.nbins;
				memcpy(M.edges, H.edges, sizeof(double) * (H.nbins + 1));
				memcpy(M.weight, H.weight, sizeof(double) * H.nbins);
				memcpy(M.bias, H.bias, sizeof(double) * H.nbins);
			} else {
				// no prior yet: weight 0 / bias 1 everywhere (pass 1 of an 'auto' run)
				M.nbins = 1;
				M.edges[0] = 1.0; M.edges[1] = 0.0;   // empty range: every lookup is "outside"
				M.weight[0] = 0.0; M.bias[0] = 1.0;
			}
			nmag++;
		}
	}
	ctx->res_nmag = nmag;
	ENSURE(ctx->d_tables, sizeof(ConstTables));
	CU(cudaMemcpyAsync(ctx->d_tables.p, &T, sizeof(ConstTables), cudaMemcpyHostToDevice, ctx->stream));
	ctx->tables_dirty = false;
	return NWB_OK;
}

inline int grid_for(long long n, int block) { return (int) std::min<long long>((n + block - 1) / block, 1 << 30); }

int column_lookup(nwb_ctx *ctx, int column, void **out)
{
	const Columns &C = ctx->cols;
	int nc = ctx->res_ncat, npairs = nc * (nc - 1) / 2;
	void *p = nullptr;
	if (column >= NWB_COL_IDX && column < NWB_COL_IDX + nc) p = C.idx[column - NWB_COL_IDX];
	else if (column >= NWB_COL_SEP && column < NWB_COL_SEP + npairs) p = C.sep[column - NWB_COL_SEP];
	else if (column >= NWB_COL_BIAS && column < NWB_COL_BIAS + ctx->res_nmag) p = C.bias[column - NWB_COL_BIAS];
	else switch (column) {
		case NWB_COL_SEPMAX: p = C.sepmax; break;
		case NWB_COL_NCAT: p = C.ncat; break;
		case NWB_COL_LOGBF_UNCORR: p = C.lbf_u; break;
		case NWB_COL_LOGBF: p = C.lbf; break;
		case NWB_COL_DIST_POST: p = C.dist_post; break;
		case NWB_COL_P_SINGLE: p = C.p_single; break;
		case NWB_COL_MATCH_FLAG: p = C.flag; break;
		case NWB_COL_P_ANY: p = C.p_any; break;
		case NWB_COL_P_I: p = C.p_i; break;
		default: break;
	}
	if (!p) return fail(ctx, NWB_ERR_ARG, "unknown column selector " + std::to_string(column));
	*out = p;
	return NWB_OK;
}

// lay the SoA columns out in one allocation
int layout_columns(nwb_ctx *ctx, DevBuf &buf, int64_t R, int nc, int nmag, Columns &C, int &ncols)
{
	int npairs = nc * (nc - 1) / 2;
	ncols = nc + npairs + 9 + nmag;
	size_t stride = ((size_t) std::max<int64_t>(R, 1) * 8 + 255) / 256 * 256;
	ENSURE(buf, stride * ncols);
	char *base = (char *) buf.p;
	int k = 0;
	auto next = [&]() { return (void *) (base + stride * (k++)); };
	for (int c = 0; c < nc; c++) C.idx[c] = (long long *) next();
	for (int i = 0; i < npairs; i++) C.sep[i] = (double *) next();
	C.sepmax = (double *) next();
	C.ncat = (long long *) next();
	C.lbf_u = (double *) next();
	C.lbf = (double *) next();
	C.dist_post = (double *) next();
	for (int j = 0; j < nmag; j++) C.bias[j] = (double *) next();
	C.p_single = (double *) next();
	C.flag = (long long *) next();
	C.p_any = (double *) next();
	C.p_i = (double *) next();
	return NWB_OK;
}

template <int NC>
int launch_rows(nwb_ctx *ctx, const RowParams &rp, bool fuse, int grid, bool any_big = true)
{
	if (rp.small_t > 0) {   // sparse primaries: one thread each
		int sgrid = (rp.np + 127) / 128;
		if (fuse) LAUNCH(ctx, (k_rows_small<NC, true>), sgrid, 128, rp);
		else LAUNCH(ctx, (k_rows_small<NC, false>), sgrid, 128, rp);
	}
	if (!any_big) return NWB_OK;   // no primary is left for the warp-per-primary kernel
	if (fuse) LAUNCH(ctx, (k_rows<NC, true>), grid, 256, rp);
	else LAUNCH(ctx, (k_rows<NC, false>), grid, 256, rp);
	return NWB_OK;
}

template <int NC>
int launch_count(nwb_ctx *ctx, const RowParams &rp, long long *rows, int grid, bool any_big = true)
{
	if (rp.small_t > 0) LAUNCH(ctx, (k_count_rows_small<NC>), (rp.np + 127) / 128, 128, rp, rows);
	if (any_big) LAUNCH(ctx, (k_count_rows<NC>), grid, 256, rp, rows);
	return NWB_OK;
}

// k_pairs<DENSE, FLAT, SKEL>: band table in shared memory and no occupancy bitmap / the reference's flat-sky bucket
// predicate on every match (NWB_COMPAT_FLAT_HASH) / the memory-system skeleton (nwb_bench_skeleton)
int launch_pairs(nwb_ctx *ctx, bool dense, bool flat, bool skel, bool scat, int n, const double *ra, const double *dec, const Grid &G,
	const int *etotal, const CellRec *cells, const Entry *entries, long long entries_cap, const K1Args &ka)
{
	// persistent: exactly one wave of resident blocks of THIS instantiation (they differ in registers), each striding over
	// the catalogue -- a grid sized for another variant's occupancy would leave SMs with an uneven number of blocks
	const int which = (dense ? 1 : 0) | (flat ? 2 : 0) | (skel ? 4 : 0) | (scat ? 8 : 0);
	if (ctx->num_sms <= 0) {
		int nsm = 0;
		CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
		ctx->num_sms = std::max(nsm, 1);
	}
	// the instantiations that exist: (dense, flat) x { plain, scatter } + (dense) x skeleton
#define NWB_KP_DISPATCH(DO) \
	if (skel) { if (dense) DO(true, false, true, false); else DO(false, false, true, false); } \
	else if (scat) { if (flat) { if (dense) DO(true, true, false, true); else DO(false, true, false, true); } \
		else { if (dense) DO(true, false, false, true); else DO(false, false, false, true); } } \
	else if (flat) { if (dense) DO(true, true, false, false); else DO(false, true, false, false); } \
	else { if (dense) DO(true, false, false, false); else DO(false, false, false, false); }
	if (ctx->k1_occ[which] <= 0) {
		int nb = 0;
#define NWB_OCC(D, F, S, X) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (k_pairs<D, F, S, X>), K1_WARPS * 32, 0))
		NWB_KP_DISPATCH(NWB_OCC)
#undef NWB_OCC
		ctx->k1_occ[which] = std::max(nb, 1);
	}
	int occ = ctx->k1_occ[which];
	size_t dyn = 0;
	if (skel && ctx->skel_blocks > 0 && ctx->skel_blocks < occ) {
		// the skeleton at the residency of the kernel it is compared with: it needs fewer registers, and MORE resident warps
		// make this access pattern slower, not faster (profiles/r02_kpairs_variants.txt) -- dynamic shared memory it never
		// touches caps the blocks per SM
		occ = ctx->skel_blocks;
		cudaFuncAttributes fa;
		if (dense) CU(cudaFuncGetAttributes(&fa, k_pairs<true, false, true, false>)); else CU(cudaFuncGetAttributes(&fa, k_pairs<false, false, true, false>));
		const size_t per_block = (size_t) 227 * 1024 / (size_t) occ;
		if (per_block > fa.sharedSizeBytes + 2048) dyn = per_block - fa.sharedSizeBytes - 1024;
		if (dense) CU(cudaFuncSetAttribute(k_pairs<true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
		else CU(cudaFuncSetAttribute(k_pairs<false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
	}
	const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(((int64_t) n + K1_WARPS * 32 - 1) / (K1_WARPS * 32), (int64_t) ctx->num_sms * occ));
	if ((int64_t) n + (int64_t) ka.s_base + (int64_t) grid * K1_WARPS * 32 + 64 > 0x7fffffffll)
		return fail(ctx, NWB_ERR_ARG, "catalogue too large: secondary indices are 32-bit");
#define NWB_KP(D, F, S, X) do { \
	k_pairs<D, F, S, X><<<grid, K1_WARPS * 32, dyn, ctx->stream>>>(n, ra, dec, G, etotal, cells, entries, entries_cap, ka); \
	ctx->launches++; \
	cudaError_t e_ = cudaGetLastError(); \
	if (e_ != cudaSuccess) return fail(ctx, NWB_ERR_CUDA, std::string("k_pairs: ") + cudaGetErrorString(e_)); } while (0)
	NWB_KP_DISPATCH(NWB_KP)
#undef NWB_KP
#undef NWB_KP_DISPATCH
	return NWB_OK;
}

}  // namespace

constexpr size_t GATHER_HEADER = 4096;           // flag words of the gather buffer: payload 0, count tags 256, completion tags 512
constexpr size_t GATHER_OFF_COUNT = 0, GATHER_OFF_COUNT_TAG = 256, GATHER_OFF_DONE_TAG = 512;
constexpr unsigned long long PEER_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;   // a rank that never arrives: give up after 10 s

// shard mode: close the peers' exchange buffers, free the own one
static void shard_release(nwb_ctx *ctx)
{
	nwb_ctx::Shard &S = ctx->shard;
	for (int r = 0; r < S.world && r < 16; r++)
		if (r != S.rank && S.peer[r]) cudaIpcCloseMemHandle(S.peer[r]);
	for (auto &p : S.peer) p = nullptr;
	release(S.xch);
	release(S.d_peers);
	S.on = S.connected = false;
}

// table all-gather: close the peers' buffers, free the own one
static void gather_release(nwb_ctx *ctx)
{
	nwb_ctx::Gather &T = ctx->gather;
	for (int r = 0; r < T.world && r < 16; r++)
		if (r != T.rank && T.peer[r]) cudaIpcCloseMemHandle(T.peer[r]);
	for (auto &p : T.peer) p = nullptr;
	for (auto &l : T.lane) if (l) { cudaStreamDestroy(l); l = nullptr; }
	for (auto &e : T.join) if (e) { cudaEventDestroy(e); e = nullptr; }
	if (T.fork) { cudaEventDestroy(T.fork); T.fork = nullptr; }
	if (T.h_words) { cudaFreeHost(T.h_words); T.h_words = nullptr; }
	release(T.d_counts);
	release(T.buf);
	T.on = T.connected = false;
}

// =========================================================================================================
extern "C" {

int nwb_version(void) { return 100; }

int nwb_create(int device, nwb_ctx **out)
{
	nwb_ctx *ctx = nullptr;
	if (!out) return fail(nullptr, NWB_ERR_ARG, "out is NULL");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(nullptr, NWB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return fail(nullptr, NWB_ERR_ARG, "device index out of range");
	e = cudaSetDevice(device);
	if (e != cudaSuccess) return fail(nullptr, NWB_ERR_CUDA, cudaGetErrorString(e));
	ctx = new nwb_ctx();
	ctx->device = device;
	e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
	ctx->stream = ctx->own_stream;
	if (e != cudaSuccess) { delete ctx; return fail(nullptr, NWB_ERR_CUDA, cudaGetErrorString(e)); }
	for (auto &ev : ctx->ev) cudaEventCreate(&ev);
	for (auto &ev : ctx->kev) cudaEventCreate(&ev);
	for (auto &m : ctx->ms) m = 0;
	for (int c = 0; c < MAXC; c++) ctx->pc[c] = 1.0;
	for (int k = 0; k < MAXP; k++) ctx->prefilter[k] = INFINITY;
	memset(&ctx->tables, 0, sizeof(ctx->tables));
	*out = ctx;
	return NWB_OK;
}

void nwb_destroy(nwb_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	DevBuf *single[] = {&ctx->d_tables, &ctx->d_prim, &ctx->d_red, &ctx->d_bands, &ctx->d_cellcnt,
		&ctx->d_entries, &ctx->d_cub, &ctx->d_pairs, &ctx->d_paircount, &ctx->d_rows, &ctx->d_rowoff, &ctx->d_matsz,
		&ctx->d_matoff, &ctx->d_mat, &ctx->d_cols, &ctx->d_cols2, &ctx->d_keep, &ctx->d_keeppos, &ctx->d_misc, &ctx->d_spill, &ctx->d_status, &ctx->d_cells, &ctx->d_worklist, &ctx->d_surv,
		&ctx->d_hrow, &ctx->d_hsrc, &ctx->d_hout};
	for (DevBuf *b : single) release(*b);
	for (int c = 0; c < MAXC; c++) {
		release(ctx->d_cnt[c]); release(ctx->d_segoff[c]); release(ctx->d_seg_s[c]); release(ctx->d_seg_sep[c]);
		release(ctx->d_Ls[c]); release(ctx->d_Lsep[c]); release(ctx->d_Ltrig[c]); release(ctx->cat[c].own);
		release(ctx->d_spilloff[c]); release(ctx->d_spillseg[c]);
	}
	shard_release(ctx);
	gather_release(ctx);
	if (ctx->h_status) cudaFreeHost(ctx->h_status);
	for (auto &ev : ctx->ev) cudaEventDestroy(ev);
	for (auto &ev : ctx->kev) cudaEventDestroy(ev);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

int nwb_set_stream(nwb_ctx *ctx, void *cuda_stream)
{
	if (!ctx) return NWB_ERR_ARG;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
	return NWB_OK;
}

const char *nwb_last_error(nwb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int nwb_set_catalogue(nwb_ctx *ctx, int c, int ncat, int64_t n, const double *ra, const double *dec,
	const double *err, int err_kind, const double *mags, int m, double area, int on_device)
{
	if (!ctx) return NWB_ERR_ARG;
	if (ncat < 2 || ncat > MAXC) return fail(ctx, NWB_ERR_ARG, "ncat must be 2.." + std::to_string(MAXC));
	if (c < 0 || c >= ncat) return fail(ctx, NWB_ERR_ARG, "catalogue index out of range");
	if (n < 0 || n >= (1ll << 31) - 64) return fail(ctx, NWB_ERR_ARG, "catalogue size must be < 2^31");
	if (!ra || !dec || !err) if (n > 0) return fail(ctx, NWB_ERR_ARG, "ra/dec/err is NULL");
	if (err_kind != NWB_ERR_CIRCULAR && err_kind != NWB_ERR_ELLIPSE) return fail(ctx, NWB_ERR_ARG, "bad err_kind");
	if (m < 0 || m > MAXM || (m > 0 && !mags)) return fail(ctx, NWB_ERR_ARG, "bad magnitude columns");
	if (!(area > 0)) return fail(ctx, NWB_ERR_ARG, "area must be > 0");
	CU(cudaSetDevice(ctx->device));
	if (ctx->ncat != ncat) {
		for (int k = 0; k < MAXC; k++) { ctx->cat[k].set = false; for (auto &h : ctx->hist[k]) h.set = false; }
		ctx->ncat = ncat;
	}
	CatSlot &S = ctx->cat[c];
	S.n = n; S.err_kind = err_kind; S.m = m; S.area = area;
	for (auto &h : ctx->hist[c]) h.set = false;
	int ecols = err_kind;
	if (on_device) {
		S.ra = ra; S.dec = dec; S.err = err; S.mags = mags;
	} else {
		size_t cols = 2 + ecols + m;
		ENSURE(S.own, std::max<size_t>(1, cols * n) * sizeof(double));
		double *d = (double *) S.own.p;
		CU(cudaMemcpyAsync(d, ra, n * 8, cudaMemcpyHostToDevice, ctx->stream));
		CU(cudaMemcpyAsync(d + n, dec, n * 8, cudaMemcpyHostToDevice, ctx->stream));
		CU(cudaMemcpyAsync(d + 2 * n, err, (size_t) ecols * n * 8, cudaMemcpyHostToDevice, ctx->stream));
		if (m) CU(cudaMemcpyAsync(d + (2 + ecols) * n, mags, (size_t) m * n * 8, cudaMemcpyHostToDevice, ctx->stream));
		S.ra = d; S.dec = d + n; S.err = d + 2 * n; S.mags = m ? d + (2 + ecols) * n : nullptr;
	}
	S.set = true;
	S.bounds_known = false;
	ctx->tables_dirty = true;
	ctx->matched = ctx->finalized = false;
	ctx->pending = false;   // a match still in flight belongs to the old catalogue: abandoned
	// one positional error for the whole catalogue?  (lets the row kernels skip a random gather per row)
	S.err_const = false;
	if (n > 0 && err_kind == NWB_ERR_CIRCULAR) {
		ENSURE(ctx->d_status, 64 * sizeof(long long));
		if (!ctx->h_status) CU(cudaHostAlloc((void **) &ctx->h_status, 64 * sizeof(long long), cudaHostAllocMapped));
		unsigned long long init[2] = {0x7ff0000000000000ull, 0ull};
		unsigned long long *d_mm = (unsigned long long *) ctx->d_status.p + 32;
		CU(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
		LAUNCH(ctx, k_minmax, (int) std::min<int64_t>((n + 255) / 256, 148 * 8), 256, (long long) n, S.err, (double *) d_mm);
		LAUNCH(ctx, k_words_to_host, 1, 32, (const long long *) d_mm, ctx->h_status + 40, 2);
		CU(cudaStreamSynchronize(ctx->stream));
		double lo, hi;
		memcpy(&lo, ctx->h_status + 40, 8);
		memcpy(&hi, ctx->h_status + 41, 8);
		S.err_const = lo == hi && lo > 0 && std::isfinite(lo);
		S.err_value = lo;
	}
	return NWB_OK;
}

int nwb_set_params(nwb_ctx *ctx, double match_radius_arcsec, const double *completeness,
	double prob_ratio_secondary, int unrelated_mode)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!(match_radius_arcsec > 0) || !(match_radius_arcsec < 3600.0 * 30))
		return fail(ctx, NWB_ERR_ARG, "match radius must be in (0, 30 deg)");
	if (ctx->ncat < 2) return fail(ctx, NWB_ERR_ARG, "set the catalogues first");
	if (!completeness) return fail(ctx, NWB_ERR_ARG, "completeness is NULL");
	if (completeness[0] != 1.0) return fail(ctx, NWB_ERR_ARG, "completeness[0] must be 1");
	if (unrelated_mode != NWB_UNRELATED_API && unrelated_mode != NWB_UNRELATED_CLI)
		return fail(ctx, NWB_ERR_ARG, "bad unrelated_mode");
	ctx->radius = match_radius_arcsec;
	for (int c = 0; c < ctx->ncat; c++) ctx->pc[c] = completeness[c];
	ctx->ratio_secondary = prob_ratio_secondary;
	ctx->unrelated_mode = unrelated_mode;
	ctx->params_set = true;
	ctx->tables_set = false;
	ctx->tables_dirty = true;
	return NWB_OK;
}

int nwb_set_prefilter(nwb_ctx *ctx, int npairs, const int *cat_a, const int *cat_b, const double *radius_arcsec)
{
	if (!ctx) return NWB_ERR_ARG;
	for (int k = 0; k < MAXP; k++) ctx->prefilter[k] = INFINITY;
	ctx->prefilter_on = false;
	if (npairs < 0 || (npairs > 0 && (!cat_a || !cat_b || !radius_arcsec))) return fail(ctx, NWB_ERR_ARG, "bad prefilter list");
	if (npairs > 0 && ctx->ncat < 2) return fail(ctx, NWB_ERR_ARG, "set the catalogues before the prefilter");
	for (int k = 0; k < npairs; k++) {
		int a = std::min(cat_a[k], cat_b[k]), b = std::max(cat_a[k], cat_b[k]);
		if (a < 0 || b >= ctx->ncat || a == b) return fail(ctx, NWB_ERR_ARG, "prefilter: bad catalogue pair");
		if (!(radius_arcsec[k] >= 0)) return fail(ctx, NWB_ERR_ARG, "prefilter: radius must be >= 0");
		double &slot = ctx->prefilter[pair_index(a, b, ctx->ncat)];
		slot = std::min(slot, radius_arcsec[k]);
		ctx->prefilter_on = true;
	}
	return NWB_OK;
}

int nwb_set_compat(nwb_ctx *ctx, int flags)
{
	if (!ctx) return NWB_ERR_ARG;
	if (flags & ~(NWB_COMPAT_SEP_F32 | NWB_COMPAT_FLAT_HASH)) return fail(ctx, NWB_ERR_ARG, "unknown compatibility flag");
	ctx->compat = flags;
	return NWB_OK;
}

int nwb_flat_hash_applied(nwb_ctx *ctx, int *applied)
{
	if (!ctx || !applied) return NWB_ERR_ARG;
	*applied = ctx->flat_err > 0.0 ? 1 : 0;
	return NWB_OK;
}

// optional: scalar tables computed by the caller with the reference's own numpy expressions
int nwb_set_tables(nwb_ctx *ctx, const double *norm /* ncat+1 */, double log10e, const double *prior,
	const double *log10prior, const double *sub_log10prior /* each 2^(ncat-1) */)
{
	if (!ctx || !ctx->params_set) return fail(ctx, NWB_ERR_ARG, "nwb_set_params first");
	int nm = 1 << (ctx->ncat - 1);
	for (int n = 0; n <= ctx->ncat; n++) ctx->tables.norm[n] = norm[n];
	ctx->tables.log10e = log10e;
	for (int k = 0; k < nm; k++) {
		ctx->tables.prior[k] = prior[k];
		ctx->tables.log10prior[k] = log10prior[k];
		ctx->tables.sub_log10prior[k] = sub_log10prior[k];
	}
	ctx->tables_set = true;
	ctx->tables_dirty = true;
	return NWB_OK;
}

int nwb_set_maghist(nwb_ctx *ctx, int c, int k, int nbins, const double *edges, const double *weight,
	const double *bias)
{
	if (!ctx) return NWB_ERR_ARG;
	if (c < 0 || c >= ctx->ncat || !ctx->cat[c].set) return fail(ctx, NWB_ERR_ARG, "catalogue not set");
	if (k < 0 || k >= ctx->cat[c].m) return fail(ctx, NWB_ERR_ARG, "magnitude column out of range");
	if (nbins < 1 || nbins > MAXB) return fail(ctx, NWB_ERR_ARG, "nbins must be 1.." + std::to_string(MAXB));
	HostMagHist &H = ctx->hist[c][k];
	H.nbins = nbins;
	for (int i = 0; i <= nbins; i++) H.edges[i] = edges[i];
	for (int i = 0; i < nbins; i++) {
		H.weight[i] = std::isnan(weight[i]) ? 0.0 : weight[i];
		H.bias[i] = std::isnan(weight[i]) ? 1.0 : bias[i];
	}
	H.set = true;
	ctx->tables_dirty = true;
	return NWB_OK;
}

int nwb_set_primary_range(nwb_ctx *ctx, int64_t first, int64_t count)
{
	if (!ctx) return NWB_ERR_ARG;
	if (first < 0 || count < -1) return fail(ctx, NWB_ERR_ARG, "bad primary range");
	ctx->first = first;
	ctx->count = count;
	return NWB_OK;
}

int nwb_sync(nwb_ctx *ctx)
{
	if (!ctx) return NWB_ERR_ARG;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

static int run_final(nwb_ctx *ctx)
{
	int grid = (int) std::min<int64_t>((ctx->np * 32 + 255) / 256, 148 * 64);
	grid = std::max(grid, 1);
	// groups of a few rows (sparse primaries): one thread each; the warp-per-primary kernel takes the rest
	if (ctx->rp.small_t > 0) LAUNCH(ctx, k_final_small, (int) ((ctx->np + 127) / 128), 128, ctx->rp);
	LAUNCH(ctx, k_final, grid, 256, ctx->rp);
	return NWB_OK;
}

// the scalar part of the row-kernel parameter block
static int fill_row_params(nwb_ctx *ctx, const PairStore *stores, int64_t first, int64_t np, bool ell)
{
	const int nc = ctx->ncat;
	RowParams &rp = ctx->rp;
	memset(&rp, 0, sizeof(rp));
	rp.ncat = nc; rp.nmag = ctx->res_nmag; rp.np = (int) np; rp.first = first;
	rp.radius = ctx->radius; rp.ratio_secondary = ctx->ratio_secondary;
	for (int c = 0; c < nc; c++) {
		rp.err[c] = ctx->cat[c].err; rp.n[c] = ctx->cat[c].n; rp.ra[c] = ctx->cat[c].ra; rp.dec[c] = ctx->cat[c].dec;
	}
	rp.ell = ell ? 1 : 0;
	rp.sep_f32 = (ctx->compat & NWB_COMPAT_SEP_F32) ? 1 : 0;
	rp.small_t = SMALL_T;
	rp.flat.err = ctx->flat_err; rp.flat.rerr = ctx->flat_err > 0.0 ? 1.0 / ctx->flat_err : 0.0;
	for (int k = 0; k < MAXP; k++) rp.pair_radius[k] = ctx->prefilter_on ? std::min(ctx->radius, ctx->prefilter[k]) : ctx->radius;
	rp.T = (const ConstTables *) ctx->d_tables.p;
	rp.S1 = stores[1];
	rp.err1_const = ctx->cat[1].err_const ? 1 : 0;
	rp.err1_value = ctx->cat[1].err_value;
	rp.guard = nullptr;
	return NWB_OK;
}

// slots per primary for the matches of catalogue c: expected number + 6 sigma (a uniform field almost never spills),
// capped so that the pair store of np primaries stays below ~6 GB
static int slots_per_primary(const nwb_ctx *ctx, int c, int64_t np)
{
	const int nc = ctx->ncat;
	const double r_deg = ctx->radius / 3600.0;
	const double mu = (double) ctx->cat[c].n * (M_PI * r_deg * r_deg) / ctx->cat[c].area;
	const double want = std::ceil(mu + 6 * std::sqrt(mu) + 2);
	const double cap_mem = std::floor(6e9 / 16.0 / (double) std::max<int64_t>(np, 1) / (nc - 1));
	int C = (int) std::max(2.0, std::min(std::min(want, cap_mem), 1e6));
	C = (C + 1) / 2 * 2;
	if (ctx->cat[c].n == 0) C = 2;
	return C;
}

// NWB_COMPAT_FLAT_HASH: would the reference's crossproduct() take its flat-sky branch for these catalogues and this
// radius (fastskymatch.py:94-98: err < 1 deg, every ra in (10 err, 360 - 10 err), every |dec| < 45)?  The bounds of a
// catalogue are reduced on the device once per nwb_set_catalogue.  *flat_err = the bucket size in degrees, or 0.
static int flat_hash_decision(nwb_ctx *ctx, double *flat_err)
{
	*flat_err = 0.0;
	if (!(ctx->compat & NWB_COMPAT_FLAT_HASH)) return NWB_OK;
	const double err = ctx->radius / 60. / 60;   // the reference's expression (__init__.py:128, nway.py:214)
	if (!ctx->h_status) CU(cudaHostAlloc((void **) &ctx->h_status, 64 * sizeof(long long), cudaHostAllocMapped));
	ENSURE(ctx->d_status, 64 * sizeof(long long));
	bool flat = err < 1;
	for (int c = 0; c < ctx->ncat && flat; c++) {
		CatSlot &S = ctx->cat[c];
		if (!S.bounds_known) {
			if (S.n > 0) {
				unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
				unsigned long long *d_b = (unsigned long long *) ctx->d_status.p + 48;
				CU(cudaMemcpyAsync(d_b, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
				LAUNCH(ctx, k_radec_bounds, (int) std::min<int64_t>((S.n + 255) / 256, 148 * 8), 256, (long long) S.n, S.ra, S.dec, d_b);
				LAUNCH(ctx, k_words_to_host, 1, 32, (const long long *) d_b, ctx->h_status + 56, 4);
				CU(cudaStreamSynchronize(ctx->stream));
				double v[3];
				for (int k = 0; k < 3; k++) {
					unsigned long long u = (unsigned long long) ctx->h_status[56 + k];
					u ^= (u >> 63) ? 0x8000000000000000ull : ~0ull;
					memcpy(&v[k], &u, 8);
				}
				S.ra_min = v[0]; S.ra_max = v[1]; S.absdec_max = v[2];
				S.has_nan = ctx->h_status[59] != 0;
			} else {
				S.ra_min = INFINITY; S.ra_max = -INFINITY; S.absdec_max = 0; S.has_nan = false;   // all() of nothing is true
			}
			S.bounds_known = true;
		}
		flat = flat && !S.has_nan && S.ra_min > 10 * err && S.ra_max < 360 - 10 * err && S.absdec_max < 45;
	}
	if (!flat) return NWB_OK;
	if (!(360.0 / err < 1073741000.0))
		return fail(ctx, NWB_ERR_ARG, "NWB_COMPAT_FLAT_HASH: the radius is too small for 32-bit flat-sky cells");
	*flat_err = err;
	return NWB_OK;
}

// phase: 0 = the whole match.  Shard mode (nwb_shard_match) runs it in two halves with a barrier between the ranks in
// between: 1 = grid over ALL primaries + this rank's slice of every secondary catalogue streamed, matches scattered to
// the owners of the primaries; 2 = lists / rows / normalisation of the primaries this rank owns.
static int match_impl(nwb_ctx *ctx, int fuse_final, int64_t *nrows, bool allow_cached, bool defer, int phase = 0)
{
	const bool shard = ctx->shard.on;
	if ((phase != 0) != shard) return fail(ctx, NWB_ERR_STATE, shard ? "shard mode: use nwb_shard_match" : "nwb_shard_setup first");
	if (shard && !ctx->shard.connected) return fail(ctx, NWB_ERR_STATE, "nwb_shard_connect first");
	const int nc = ctx->ncat;
	if (nc < 2) return fail(ctx, NWB_ERR_ARG, "no catalogues");
	for (int c = 0; c < nc; c++) {
		if (!ctx->cat[c].set) return fail(ctx, NWB_ERR_ARG, "catalogue " + std::to_string(c) + " not set");
	}
	bool ell = false;
	for (int c = 0; c < nc; c++) ell = ell || ctx->cat[c].err_kind == NWB_ERR_ELLIPSE;
	if (ell)
		for (int c = 0; c < nc; c++)
			if (ctx->cat[c].err_kind != NWB_ERR_ELLIPSE)
				return fail(ctx, NWB_ERR_ARG, "elliptical mode: every catalogue must carry (sigma_x, sigma_y, rho); pass (s, s, 0) for circular ones (nway.py:79-88)");
	if (!ctx->params_set) return fail(ctx, NWB_ERR_ARG, "nwb_set_params not called");
	CU(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	ctx->matched = ctx->finalized = false;
	ctx->pending = false;
	ctx->launches = 0;
	// two views of the primary catalogue: the rows this context produces (first, np) and the primaries its grid indexes
	// (gfirst, gnp) -- the same unless the streaming is sharded, where every rank's grid holds ALL primaries
	int64_t first = ctx->first, np = ctx->count < 0 ? ctx->cat[0].n - ctx->first : ctx->count;
	if (shard) {
		first = std::min<int64_t>(ctx->shard.rank * ctx->shard.block, ctx->cat[0].n);
		np = std::min<int64_t>(ctx->shard.block, ctx->cat[0].n - first);
	}
	if (first + np > ctx->cat[0].n || np < 0) return fail(ctx, NWB_ERR_ARG, "primary range exceeds the catalogue");
	const int64_t gfirst = shard ? 0 : first, gnp = shard ? ctx->cat[0].n : np;
	ctx->np = np;
	if (gnp == 0 || (np == 0 && phase != 1)) { if (nrows) *nrows = 0; ctx->nrows = 0; return fail(ctx, NWB_ERR_EMPTY, "No matches."); }
	if (!ctx->tables_set && ctx->tables_dirty) default_tables(ctx);
	{ int r = upload_tables(ctx); if (r) return r; }
	const bool cli = ctx->unrelated_mode == NWB_UNRELATED_CLI && nc >= 3;
	const bool fuse = fuse_final && !cli;
	double flat_err = 0.0;
	{ int r = flat_hash_decision(ctx, &flat_err); if (r) return r; }
	ctx->flat_err = flat_err;
	FlatHash flat;
	flat.err = flat_err; flat.rerr = flat_err > 0.0 ? 1.0 / flat_err : 0.0;
	if (!ctx->h_status) CU(cudaHostAlloc((void **) &ctx->h_status, 64 * sizeof(long long), cudaHostAllocMapped));
	long long *hs = ctx->h_status;

	if (phase != 2) CU(cudaEventRecord(ctx->ev[0], st));
	// ---- K0: primaries -> bounding box (the one unavoidable early sync: the grid geometry is chosen on the host)
	const double r_deg = ctx->radius / 3600.0;
	const double rb = r_deg * (1 + 1e-9) + 1e-12;
	const double rb_ins = rb + 1e-9, dra_eps = 1e-9;
	const double entry_tau_max = (rb_ins * M_PI / 180 > 0.02) ? -1.0 : 0.02;   // the pole rule of pretest_constants (Grid::tau_max)
	ENSURE(ctx->d_prim, (size_t) gnp * 8 * sizeof(double));
	PrimArrays P;
	{
		double *b = (double *) ctx->d_prim.p;
		P.rec = (PrimRec *) b; P.clat = b + 4 * gnp; P.ra_n = b + 5 * gnp; P.dec = b + 6 * gnp; P.dra = b + 7 * gnp;
	}
	const int gblocks = grid_for(gnp, 256);   // one thread per primary of the grid
	int pblocks = grid_for(std::max<int64_t>(np, 1), 256);   // ... per primary whose rows this context writes
	ENSURE(ctx->d_red, 8 * sizeof(double));
	unsigned long long *d_red = (unsigned long long *) ctx->d_red.p;
	const bool use_cached = allow_cached && ctx->geom_valid && ctx->geom_rb == rb && ctx->geom_np == gnp && ctx->geom_first == gfirst;
	if (!use_cached && phase != 2) {
		CU(cudaMemsetAsync(d_red, 0, 6 * sizeof(unsigned long long), st));
		LAUNCH(ctx, (k_prim_prep<false>), gblocks, 256, (int) gnp, (long long) gfirst, ctx->cat[0].ra, ctx->cat[0].dec, rb, P, d_red,
			Grid(), rb_ins, dra_eps, (int *) nullptr, entry_tau_max, flat, (CellRec *) nullptr, (OverflowItem *) nullptr, (int *) nullptr, (long long) 0);
		// the grid geometry is chosen on the host from the bounding box: one sync.  It is kept for the next match
		// on this context, which only has to verify (on the device) that the box is still the same.
		unsigned long long *raw = (unsigned long long *) (hs + 32);
		CU(cudaMemcpyAsync(raw, d_red, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		double red[6];
		for (int k = 0; k < 6; k++) {   // undo the order-preserving encoding; minima were stored negated
			ctx->geom_key.k[k] = raw[k];
			unsigned long long u = raw[k];
			u ^= (u >> 63) ? 0x8000000000000000ull : ~0ull;
			memcpy(&red[k], &u, 8);
			if (!(k & 1)) red[k] = -red[k];
		}
		HostGrid HG;
		// cells: ~16 per primary is plenty (each primary occupies 1-9 of them), and the 32-byte records of at most 2 M
		// cells (64 MB) stay in the 126 MB L2 next to the streamed catalogue
		long long max_cells = std::min<long long>(2ll << 20, std::max<long long>(1ll << 16, 16 * (long long) gnp));
		build_grid(red, rb_ins, rb_ins * NWB_CELL_FACTOR, max_cells, HG);
		pretest_constants(HG, rb_ins);
		size_t nb = (size_t) HG.g.nbands;
		ENSURE(ctx->d_bands, nb * (sizeof(BandRec) + sizeof(float)) + 64);
		CU(cudaMemcpyAsync(ctx->d_bands.p, HG.bands.data(), nb * sizeof(BandRec), cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync((char *) ctx->d_bands.p + nb * sizeof(BandRec), HG.kx.data(), nb * sizeof(float), cudaMemcpyHostToDevice, st));
		CU(cudaStreamSynchronize(st));   // HG.bands / HG.kx are temporaries
		HG.g.bands = (const BandRec *) ctx->d_bands.p;
		HG.g.kx = (const float *) ((const char *) ctx->d_bands.p + nb * sizeof(BandRec));
		{
			// the regular bitmap of the pure stream (k_filter): as many ra cells per band as the widest band of the grid has
			int widest = 1;
			for (const BandRec &B : HG.bands) widest = std::max(widest, B.nra);
			// ... scaled down, both ways alike, until it fits the shared memory of one block (KF_SMEM_WORDS)
			const double kf_bits = 32.0 * KF_SMEM_WORDS;
			const double f = std::sqrt(std::min(1.0, kf_bits / ((double) HG.g.nbands * (double) ((widest + 31) / 32 * 32))));
			HG.g.nr2 = std::max(32, (int) ((double) ((widest + 31) / 32 * 32) * f) / 32 * 32);
			HG.g.nb2 = (int) std::max<long long>(1, std::min<long long>(HG.g.nbands, (long long) kf_bits / HG.g.nr2));
			// the cell functions ROUND (kf_round): scales for nr2 - 1 and nb2 - 1 intervals, a hair low, so that coordinates inside
			// the grid need no clamping (kf_occupied)
			HG.g.inv_w2 = (double) (HG.g.nr2 - 1) / HG.g.ra_span * (1.0 - 1e-12);
			HG.g.band2_scale = (double) (HG.g.nb2 - 1) / (double) HG.g.nbands * (1.0 - 1e-12);
			HG.g.bits2 = nullptr;
		}
		ctx->geom_G = HG.g;
		ctx->geom_rb = rb; ctx->geom_np = gnp; ctx->geom_first = gfirst;
		ctx->geom_valid = true;
	}
	{
		// cell records, then (sparse primaries only) the occupancy bitmap behind them
		Grid &g = ctx->geom_G;
		const size_t rec_bytes = ((size_t) g.ncells * sizeof(CellRec) + 255) / 256 * 256;
		const size_t bit_bytes = (((size_t) g.ncells / 32 + 2) * sizeof(unsigned) + 255) / 256 * 256;
		const size_t bit2_bytes = ((size_t) g.nb2 * g.nr2 / 32 + 8) * sizeof(unsigned);
		ENSURE(ctx->d_cells, rec_bytes + bit_bytes + bit2_bytes);
		const double s_deg = 1.0 / g.inv_h;
		const double reach = 1.0 + 2.0 * rb_ins / s_deg;   // cells a primary's box spans along one axis, on average
		const bool sparse = (double) gnp * reach * reach < 0.5 * (double) g.ncells;
		g.bits = sparse ? (const unsigned *) ((const char *) ctx->d_cells.p + rec_bytes) : nullptr;
		g.bits2 = sparse ? (const unsigned *) ((const char *) ctx->d_cells.p + rec_bytes + bit_bytes) : nullptr;
		ctx->geom_occ = (double) gnp * reach * reach / (double) g.ncells;   // expected share of occupied cells (an upper estimate)
	}
	const Grid G = ctx->geom_G;
	// one zero-initialised block: cell counters | per-catalogue match counters | scalar counters
	const size_t ncell1 = (size_t) G.ncells + 1;
	const size_t cnt_stride = ((size_t) np + 1 + 3) / 4 * 4;
	const size_t zero_ints = (ncell1 + 3) / 4 * 4 + cnt_stride * (nc - 1) + 8 * MAXC;
	ENSURE(ctx->d_cellcnt, zero_ints * sizeof(int));
	int *d_cellcnt = (int *) ctx->d_cellcnt.p;
	int *d_cnt[MAXC] = {nullptr};
	for (int c = 1; c < nc; c++) d_cnt[c] = d_cellcnt + (ncell1 + 3) / 4 * 4 + cnt_stride * (c - 1);
	unsigned long long *d_spillcount = (unsigned long long *) (d_cellcnt + (ncell1 + 3) / 4 * 4 + cnt_stride * (nc - 1));
	int *d_etotal = (int *) (d_spillcount + 12);   // [0] overflow entries of the cell lists, [1] registrations, [2] K0 work list
	unsigned long long *d_survn = (unsigned long long *) (d_etotal + 8);   // [c]: survivors of k_filter for catalogue c
	const size_t xset = shard ? (size_t) (ctx->shard.epoch & 1) * ctx->shard.set_bytes : 0;   // this match's set of the exchange buffer
	char *xch = (char *) ctx->shard.xch.p + xset;
	if (shard) {   // match counters and spill counters of the own primaries live in the exchange buffer, where the peers write
		for (int c = 1; c < nc; c++) d_cnt[c] = (int *) (xch + ctx->shard.off_cnt[c]);
		d_spillcount = (unsigned long long *) (xch + ctx->shard.off_spillcnt[0]);
	}

	// slots per primary and catalogue (slots_per_primary); shard mode: the exchange buffer's, fixed by nwb_shard_setup
	int Cs[MAXC] = {0};
	size_t base_off[MAXC + 1] = {0};
	for (int c = 1; c < nc; c++) {
		Cs[c] = shard ? ctx->shard.C[c] : slots_per_primary(ctx, c, np);
		base_off[c + 1] = base_off[c] + (shard ? 0 : (size_t) np * Cs[c]);
	}
	ENSURE(ctx->d_pairs, std::max<size_t>(1, base_off[nc]) * sizeof(Slot16));
	Slot16 *d_base = (Slot16 *) ctx->d_pairs.p;
	if (ctx->spill_cap < 65536) ctx->spill_cap = 65536;
	if (shard) ctx->spill_cap = ctx->shard.spill_cap;
	if (ctx->entries_cap < (size_t) gnp * 12 + 4096) ctx->entries_cap = (size_t) gnp * 12 + 4096;
	ENSURE(ctx->d_rows, (size_t) (np + 1) * sizeof(long long));
	ENSURE(ctx->d_rowoff, (size_t) (np + 1) * sizeof(long long));
	long long *d_rows = (long long *) ctx->d_rows.p, *d_rowoff = (long long *) ctx->d_rowoff.p;
	ENSURE(ctx->d_status, 64 * sizeof(long long));
	long long *d_status = (long long *) ctx->d_status.p;

	const bool generic = nc > 2 || ell;   // N == 2 with circular errors takes the specialised row kernel
	PairStore stores[MAXC];
	memset(stores, 0, sizeof(stores));
	long long R = 0;
	bool done = false, speculated = false;
	for (int attempt = 0; !done && attempt < 4; attempt++) {
		ENSURE(ctx->d_entries, ctx->entries_cap * sizeof(Entry));
		if (!shard) ENSURE(ctx->d_spill, (size_t) ctx->spill_cap * (nc - 1) * sizeof(SpillRec));
		Entry *d_entries = (Entry *) ctx->d_entries.p;
		SpillRec *d_spill = (SpillRec *) ctx->d_spill.p;
		CellRec *d_cells = (CellRec *) ctx->d_cells.p;
		if (phase == 2) {
			// shard mode, second half: the grid and the streaming were phase 1
		} else if (attempt == 0 && use_cached) {
			// known geometry: the primaries are counted into their cells by the preparation kernel itself
			LAUNCH(ctx, k_zero, (int) std::min<size_t>((zero_ints / 4 + 255) / 256, 148 * 8), 256, (int4 *) d_cellcnt, (long long) (zero_ints / 4), d_red, 6);
			// ... and placed: the first three of a cell inline, the rest noted in the work list (one pass over the primaries)
			ENSURE(ctx->d_worklist, ctx->entries_cap * sizeof(OverflowItem));
			LAUNCH(ctx, (k_prim_prep<true>), gblocks, 256, (int) gnp, (long long) gfirst, ctx->cat[0].ra, ctx->cat[0].dec, rb, P, d_red,
				G, rb_ins, dra_eps, d_cellcnt, entry_tau_max, flat, d_cells, (OverflowItem *) ctx->d_worklist.p, d_etotal + 2, (long long) ctx->entries_cap);
			LAUNCH(ctx, k_cell_headers, grid_for(G.ncells, 256), 256, (long long) G.ncells, (const int *) d_cellcnt, d_cells, (unsigned *) G.bits, d_etotal);
			LAUNCH(ctx, k_fill_overflow, 148 * 2, 256, G, P, (const OverflowItem *) ctx->d_worklist.p, (const int *) (d_etotal + 2), (long long) ctx->entries_cap,
				(const CellRec *) d_cells, d_entries, (const int *) d_etotal, (long long) ctx->entries_cap);
		} else {
			CU(cudaMemsetAsync(d_cellcnt, 0, zero_ints * sizeof(int), st));
			LAUNCH(ctx, (k_prim_cells<false>), grid_for(gnp * 4, 256), 256, (int) gnp, G, P, rb_ins, dra_eps, d_cellcnt, (CellRec *) nullptr,
				(Entry *) nullptr, (const int *) nullptr, (long long) 0);
			LAUNCH(ctx, k_cell_headers, grid_for(G.ncells, 256), 256, (long long) G.ncells, (const int *) d_cellcnt, d_cells, (unsigned *) G.bits, d_etotal);
			LAUNCH(ctx, (k_prim_cells<true>), grid_for(gnp * 4, 256), 256, (int) gnp, G, P, rb_ins, dra_eps, d_cellcnt, d_cells,
				d_entries, (const int *) d_etotal, (long long) ctx->entries_cap);
		}
		if (attempt == 0 && phase != 2) CU(cudaEventRecord(ctx->ev[1], st));

		// the regular bitmap of the pure stream, when some catalogue will take the two-kernel route (see below)
		if (phase != 2 && G.bits2 && ctx->geom_occ < 0.2) {
			bool any_long = false;
			for (int c = 1; c < nc; c++) {
				const int64_t n = ctx->cat[c].n;
				const int64_t cnt = shard ? n * (ctx->shard.rank + 1) / ctx->shard.world - n * ctx->shard.rank / ctx->shard.world : n;
				any_long = any_long || cnt >= (1 << 20);
			}
			if (any_long) {
				CU(cudaMemsetAsync((void *) G.bits2, 0, ((size_t) G.nb2 * G.nr2 / 32 + 8) * sizeof(unsigned), st));
				LAUNCH(ctx, k_prim_bits2, gblocks, 256, (int) gnp, G, P, rb_ins, dra_eps, (unsigned *) G.bits2);
			}
		}
		// ---- K1: stream the secondaries ----------------------------------------------------------------
		for (int c = 1; c < nc; c++) {
			int64_t n = ctx->cat[c].n;
			stores[c].base = shard ? (const Slot16 *) (xch + ctx->shard.off_slot[c]) : d_base + base_off[c];
			stores[c].C = Cs[c];
			stores[c].cnt = d_cnt[c];
			stores[c].spill_off = nullptr;
			stores[c].spill = nullptr;
			if (n == 0 || phase == 2) continue;
			CU(cudaEventRecord(ctx->kev[2 * c], st));
			K1Args ka;
			memset(&ka, 0, sizeof(ka));
			ka.P = P; ka.radius = ctx->prefilter_on ? std::min(ctx->radius, ctx->prefilter[pair_index(0, c, nc)]) : ctx->radius; ka.base = d_base + base_off[c]; ka.C = Cs[c]; ka.cnt = d_cnt[c];
			ka.spill = d_spill + (size_t) ctx->spill_cap * (c - 1); ka.spill_cap = (unsigned long long) ctx->spill_cap;
			ka.spill_count = d_spillcount + c;
			ka.flat = flat;
			// shard mode: this rank streams its slice [s_first, s_first + s_count) of the catalogue; the matches go to the owners
			int64_t s_first = 0, s_count = n;
			if (shard) {
				s_first = n * ctx->shard.rank / ctx->shard.world;
				s_count = n * (ctx->shard.rank + 1) / ctx->shard.world - s_first;
				ka.s_base = (int) s_first;
				ka.x_block = (int) ctx->shard.block;
				ka.x_peers = (char *const *) ctx->shard.d_peers.p;
				ka.x_cnt_off = (long long) (xset + ctx->shard.off_cnt[c]); ka.x_slot_off = (long long) (xset + ctx->shard.off_slot[c]);
				ka.x_spill_off = (long long) (xset + ctx->shard.off_spill[c]); ka.x_spillcnt_off = (long long) (xset + ctx->shard.off_spillcnt[c]);
			}
			ctx->last_k1[c] = ka; ctx->last_k1_dense = G.nbands <= K1_SBANDS && !G.bits;
			ctx->last_etotal = d_etotal;
			// sparse primaries and a long catalogue: the stream as two kernels -- k_filter (coordinates -> bitmap bit, survivors
			// appended to a list) at full occupancy, then k_pairs over the few per cent that survive.  Should the list overflow
			// (denser than estimated) the second k_pairs launch streams the catalogue directly; otherwise it does nothing.
			const bool two_kernels = G.bits && G.bits2 && s_count >= (1 << 20) && ctx->geom_occ < 0.2;
			if (two_kernels && s_count > 0) {
				if (ctx->filter_occ <= 0) {
					int nsm = 0;
					CU(cudaFuncSetAttribute(k_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, KF_SMEM_WORDS * (int) sizeof(unsigned)));
					CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
					ctx->filter_occ = 1;   // one block of KF_THREADS per SM: the bitmap takes its shared memory
					ctx->num_sms = std::max(nsm, 1);
				}
				// one segment of survivor records per block: expected survivors x 3 + a margin (the blocks stride evenly over the
				// catalogue, so a uniform sky fills them alike; a patch far denser than the average overflows a segment)
				const int nseg = ctx->num_sms * ctx->filter_occ;
				// expected share of occupied cells of the coarse bitmap (an upper estimate, as geom_occ is for the grid)
				const double occ2 = std::min(1.0, ctx->geom_occ * (double) G.ncells / ((double) G.nb2 * (double) G.nr2) * 2.0);
				const int segcap = (int) std::min<double>(2e9 / nseg, ((double) s_count / nseg) * (3.0 * occ2 + 0.02) + 512.0);
				const size_t cap = (size_t) nseg * segcap;
				ENSURE(ctx->d_surv, cap * (sizeof(int) + sizeof(double2)) + ((size_t) nseg + 1) * sizeof(int) + 512);
				double2 *d_surv_rd = (double2 *) ctx->d_surv.p;   // coordinates first (16-byte aligned), indices behind, then the counts
				int *d_surv_idx = (int *) (d_surv_rd + cap);
				int *d_surv_cnt = d_surv_idx + cap;
				CU(cudaMemsetAsync(d_surv_cnt + nseg, 0, sizeof(int), st));
				{
					const size_t smem = ((size_t) G.nb2 * (G.nr2 / 32) + 3) / 4 * 4 * sizeof(unsigned);
					k_filter<<<nseg, KF_THREADS, smem, st>>>((int) s_count, ctx->cat[c].ra + s_first, ctx->cat[c].dec + s_first, G,
						d_surv_idx, d_surv_rd, d_surv_cnt, segcap);
					ctx->launches++;
					cudaError_t e_ = cudaGetLastError();
					if (e_ != cudaSuccess) return fail(ctx, NWB_ERR_CUDA, std::string("k_filter: ") + cudaGetErrorString(e_));
				}
				ka.surv = d_surv_idx; ka.surv_rd = d_surv_rd; ka.surv_cnt = d_surv_cnt; ka.surv_nseg = nseg; ka.surv_segcap = segcap;
				for (int mode = 1; mode <= 2; mode++) {
					ka.surv_mode = mode;
					int r = launch_pairs(ctx, false, flat_err > 0.0, false, shard, (int) s_count, ctx->cat[c].ra + s_first, ctx->cat[c].dec + s_first, G,
						(const int *) d_etotal, (const CellRec *) d_cells, (const Entry *) d_entries, (long long) ctx->entries_cap, ka);
					if (r) return r;
				}
				ka.surv_mode = 0;
				ctx->last_k1[c] = ka;
			} else if (s_count > 0) {
				int r = launch_pairs(ctx, ctx->last_k1_dense, flat_err > 0.0, false, shard, (int) s_count, ctx->cat[c].ra + s_first, ctx->cat[c].dec + s_first, G,
					(const int *) d_etotal, (const CellRec *) d_cells, (const Entry *) d_entries, (long long) ctx->entries_cap, ka);
				if (r) return r;
			}
			CU(cudaEventRecord(ctx->kev[2 * c + 1], st));
		}
		if (attempt == 0 && phase != 2) CU(cudaEventRecord(ctx->ev[2], st));
		if (phase == 1) {   // the caller puts a barrier between the ranks here: every rank's matches must have arrived
			ctx->shard_events_valid = true;
			return NWB_OK;
		}
		if (!generic && np <= 4 * RO_THREADS * 4) {
			// a few thousand primaries (the launch-latency-bound regime): row offsets + status words in one single-block launch
			LAUNCH(ctx, k_rowoff_status, 1, RO_THREADS, (int) np, (const int *) d_cnt[1], d_rowoff, nc, (const int *) d_etotal,
				(const unsigned long long *) d_spillcount, (const unsigned long long *) d_red, ctx->geom_key, d_status, hs);
		} else {
			if (!generic) { int r = scan_rows2(ctx, (const int *) d_cnt[1], d_rowoff, np); if (r) return r; }
			LAUNCH(ctx, k_collect_status, 1, 32, nc, (const int *) d_etotal, (const unsigned long long *) d_spillcount,
				!generic ? (const long long *) d_rowoff + np : (const long long *) nullptr, (const unsigned long long *) d_red,
				ctx->geom_key, d_status, hs);
		}
		// speculative K2: if the table of the previous match was big enough, launch the row kernel right away; it
		// checks the status words on the device.  One host sync per match instead of three.
		speculated = false;
		if (!generic && !shard && ctx->cols_cap_rows > 0 && ctx->cols_cap_ncols == 2 + 1 + 9 + ctx->res_nmag) {
			int r = fill_row_params(ctx, stores, first, np, ell);
			if (r) return r;
			ctx->rp.row_off = d_rowoff;
			{ int r2 = layout_columns(ctx, ctx->d_cols, ctx->cols_cap_rows, nc, ctx->res_nmag, ctx->cols, ctx->ncols); if (r2) return r2; }
			ctx->rp.C = ctx->cols;
			ctx->rp.guard = d_status;
			ctx->rp.max_rows = ctx->cols_cap_rows;
			ctx->rp.entries_cap = (long long) ctx->entries_cap;
			CU(cudaEventRecord(ctx->ev[3], st));
			CU(cudaEventRecord(ctx->kev[0], st));
			int grid2 = (int) std::min<int64_t>((np + R2_WARPS - 1) / R2_WARPS, 148 * NWB_R2_GRID_PER_SM);
			if (fuse && ctx->res_nmag == 0 && NWB_R2_SHARE) LAUNCH(ctx, (k_rows2<true, true>), grid2, R2_WARPS * 32, ctx->rp);
			else if (fuse) LAUNCH(ctx, (k_rows2<true, false>), grid2, R2_WARPS * 32, ctx->rp);
			else LAUNCH(ctx, (k_rows2<false, false>), grid2, R2_WARPS * 32, ctx->rp);
			CU(cudaEventRecord(ctx->kev[1], st));
			speculated = true;
		}
		if (defer && speculated && use_cached && attempt == 0) {
			// nwb_match_async: everything of this match is in the stream (the row kernel checks the status words on the
			// device); nwb_match_wait looks at the same words on the host and redoes the match if they say no
			CU(cudaEventRecord(ctx->ev[4], st));
			CU(cudaEventRecord(ctx->ev[5], st));
			ctx->pending = true;
			ctx->pending_fuse = fuse_final;
			return NWB_OK;
		}
		CU(cudaStreamSynchronize(st));
		if (hs[9] != 0) {   // the primaries moved: the cached geometry is stale
			ctx->geom_valid = false;
			return 1;       // caller retries without the cache
		}
		done = true;
		if (shard && hs[9] != 0) { ctx->geom_valid = false; return 1; }
		if ((size_t) hs[0] > ctx->entries_cap) { ctx->entries_cap = (size_t) hs[0] + 1024; done = false; }
		for (int c = 1; c < nc; c++)
			if ((unsigned long long) hs[c] > ctx->spill_cap) { ctx->spill_cap = (unsigned long long) hs[c] + 1024; done = false; }
		R = hs[8];
		ctx->stats[3] = hs[10];
		if (speculated && !(done && hs[1] == 0 && R <= ctx->cols_cap_rows)) speculated = false;   // the kernel declined
		if (shard && !done) {
			// the streaming is shared between the ranks: a bigger buffer takes effect when ALL of them redo the match
			for (int c = 1; c < nc; c++)
				if ((unsigned long long) hs[c] > ctx->shard.spill_cap)
					return fail(ctx, NWB_ERR_NOMEM, "shard mode: more overflowing matches than the exchange buffer holds (a very clustered catalogue): nwb_shard_setup with a larger spill capacity");
			return 1;
		}
	}
	if (!done) return fail(ctx, NWB_ERR_NOMEM, "grid / spill buffers kept overflowing");
	ctx->stats[2] = G.ncells;

	// ---- overflowed primaries (rare): spill records -> per-primary spill segments ------------------------
	for (int c = 1; c < nc; c++) {
		long long nsp = hs[c];
		if (nsp == 0) continue;
		ENSURE(ctx->d_spilloff[c], (size_t) (np + 1) * sizeof(long long));
		ENSURE(ctx->d_spillseg[c], (size_t) nsp * sizeof(Slot16));
		ENSURE(ctx->d_matsz, (size_t) (np + 1) * sizeof(long long));
		int *sizes = (int *) ctx->d_matsz.p;
		CU(cudaMemsetAsync(sizes, 0, (size_t) (np + 1) * sizeof(int), st));
		LAUNCH(ctx, k_spill_sizes, pblocks, 256, (int) np, (const int *) d_cnt[c], Cs[c], sizes);
		{ int r = scan_int_to_ll(ctx, sizes, (long long *) ctx->d_spilloff[c].p, np + 1); if (r) return r; }
		LAUNCH(ctx, k_spill_scatter, grid_for(nsp, 256), 256, nsp, shard ? (const SpillRec *) (xch + ctx->shard.off_spill[c]) : (const SpillRec *) ctx->d_spill.p + (size_t) ctx->spill_cap * (c - 1),
			Cs[c], (const long long *) ctx->d_spilloff[c].p, (Slot16 *) ctx->d_spillseg[c].p);
		stores[c].spill_off = (const long long *) ctx->d_spilloff[c].p;
		stores[c].spill = (const Slot16 *) ctx->d_spillseg[c].p;
	}

	RowParams &rp = ctx->rp;
	if (!speculated) { int r = fill_row_params(ctx, stores, first, np, ell); if (r) return r; }
	int wgrid = std::max(1, (int) std::min<int64_t>((np * 32 + 255) / 256, 148 * 64));
	ctx->stats[1] = 0;

	// ---- N >= 3 / elliptical, speculative: lists, separation scratch, row count, rows and normalisation enqueued back to
	// back behind device-side gates, ONE synchronisation at the end; taken when the previous match of this context had
	// the same shape (its buffers and launch decisions are reused) and nothing spilled.  If a gate stays shut -- more
	// matches than the lists hold, more rows than the table, a primary that needs the warp-per-primary kernels where
	// none were launched -- the stage-by-stage path below redoes the work.
	bool generic_done = false;
	{
		bool can = generic && ctx->gen.valid && ctx->gen.nc == nc && ctx->gen.ell == (int) ell && ctx->gen.np == np &&
			ctx->cols_cap_rows > 0 && ctx->cols_cap_ncols == nc + nc * (nc - 1) / 2 + 9 + ctx->res_nmag;
		for (int c = 1; c < nc; c++) can = can && hs[c] == 0;
		if (can) {
			const nwb_ctx::GenCaps &Gc = ctx->gen;
			int *d_gates = (int *) (d_status + 52);
			Lists L;
			memset(&L, 0, sizeof(L));
			GateArgs ga;
			memset(&ga, 0, sizeof(ga));
			ga.ncat = nc; ga.gates = d_gates; ga.host = hs; ga.maxcnt = (const int *) (d_status + 40);
			CU(cudaMemsetAsync(d_status + 40, 0, 4 * sizeof(long long), st));
			for (int c = 1; c < nc; c++) {
				ENSURE(ctx->d_segoff[c], (size_t) (np + 1) * sizeof(long long));
				long long *off = (long long *) ctx->d_segoff[c].p;
				{ int r = scan_int_to_ll(ctx, (const int *) d_cnt[c], off, np + 1); if (r) return r; }
				LAUNCH(ctx, k_max_int, (int) std::min<int64_t>((np + 255) / 256, 148 * 8), 256, (long long) np, (const int *) d_cnt[c], (int *) (d_status + 40) + c);
				ga.seg_total[c] = off + np; ga.cap_list[c] = Gc.cap_list[c]; ga.big_sort[c] = Gc.big_sort[c];
			}
			ga.level = 1;
			LAUNCH(ctx, k_spec_gate, 1, 32, ga, SMALL_N, SMALL_T);
			for (int c = 1; c < nc; c++) {
				const size_t cap = (size_t) Gc.cap_list[c];
				double *tr = (double *) ctx->d_Ltrig[c].p;
				const long long *off = (const long long *) ctx->d_segoff[c].p;
				LAUNCH(ctx, k_sort_lists_small, pblocks, 256, (int) np, stores[c], off, (int *) ctx->d_Ls[c].p, (double *) ctx->d_Lsep[c].p,
					ctx->cat[c].ra, ctx->cat[c].dec, tr, tr + cap, tr + 2 * cap, (long long *) (tr + 3 * cap), flat, (const int *) (d_gates + 1));
				if (Gc.big_sort[c])
					LAUNCH(ctx, k_sort_lists, wgrid, 256, (int) np, stores[c], off, (int *) ctx->d_Ls[c].p, (double *) ctx->d_Lsep[c].p,
						ctx->cat[c].ra, ctx->cat[c].dec, tr, tr + cap, tr + 2 * cap, SMALL_N, (long long *) (tr + 3 * cap), flat, (const int *) (d_gates + 1));
				L.off[c] = off;
				L.s[c] = (const int *) ctx->d_Ls[c].p;
				L.sep[c] = (const double *) ctx->d_Lsep[c].p;
				L.lon[c] = tr; L.slat[c] = tr + cap; L.clat[c] = tr + 2 * cap;
				L.ij[c] = (const long long *) (tr + 3 * cap);
			}
			rp.L = L;
			CU(cudaEventRecord(ctx->ev[3], st));
			ENSURE(ctx->d_matsz, (size_t) (np + 1) * sizeof(long long));
			ENSURE(ctx->d_matoff, (size_t) (np + 1) * sizeof(long long));
			long long *d_matsz = (long long *) ctx->d_matsz.p, *d_matoff = (long long *) ctx->d_matoff.p;
			CU(cudaMemsetAsync(d_matsz, 0, (size_t) (np + 1) * sizeof(long long), st));
			unsigned long long *d_maxtup = (unsigned long long *) (d_status + 44);
			CU(cudaMemsetAsync(d_maxtup, 0, sizeof(unsigned long long), st));
			LAUNCH(ctx, k_mat_sizes, pblocks, 256, (int) np, nc, L, d_matsz, d_maxtup, (const int *) (d_gates + 1));
			{ int r = scan_ll(ctx, d_matsz, d_matoff, np + 1); if (r) return r; }
			ga.level = 2; ga.mat_total = d_matoff + np; ga.cap_mat = Gc.cap_mat; ga.max_tuples = d_maxtup; ga.any_big = Gc.any_big;
			LAUNCH(ctx, k_spec_gate, 1, 32, ga, SMALL_N, SMALL_T);
			rp.mat_off = d_matoff;
			rp.mat = (double *) ctx->d_mat.p;
			CU(cudaMemsetAsync(d_rows, 0, (size_t) (np + 1) * sizeof(long long), st));
			rp.gate = d_gates + 2;
			const bool big = Gc.any_big != 0;
			int r = 0;
			switch (nc) {
				case 2: r = launch_count<2>(ctx, rp, d_rows, wgrid, big); break;
				case 3: r = launch_count<3>(ctx, rp, d_rows, wgrid, big); break;
				case 4: r = launch_count<4>(ctx, rp, d_rows, wgrid, big); break;
				case 5: r = launch_count<5>(ctx, rp, d_rows, wgrid, big); break;
				case 6: r = launch_count<6>(ctx, rp, d_rows, wgrid, big); break;
				case 7: r = launch_count<7>(ctx, rp, d_rows, wgrid, big); break;
				default: r = launch_count<8>(ctx, rp, d_rows, wgrid, big); break;
			}
			if (r) return r;
			{ int r2 = scan_ll(ctx, d_rows, d_rowoff, np + 1); if (r2) return r2; }
			ga.level = 3; ga.rows_total = d_rowoff + np; ga.cap_rows = ctx->cols_cap_rows;
			LAUNCH(ctx, k_spec_gate, 1, 32, ga, SMALL_N, SMALL_T);
			rp.row_off = d_rowoff;
			{ int r2 = layout_columns(ctx, ctx->d_cols, ctx->cols_cap_rows, nc, ctx->res_nmag, ctx->cols, ctx->ncols); if (r2) return r2; }
			rp.C = ctx->cols;
			rp.gate = d_gates + 3;
			CU(cudaEventRecord(ctx->kev[0], st));
			switch (nc) {
				case 2: r = launch_rows<2>(ctx, rp, fuse, wgrid, big); break;
				case 3: r = launch_rows<3>(ctx, rp, fuse, wgrid, big); break;
				case 4: r = launch_rows<4>(ctx, rp, fuse, wgrid, big); break;
				case 5: r = launch_rows<5>(ctx, rp, fuse, wgrid, big); break;
				case 6: r = launch_rows<6>(ctx, rp, fuse, wgrid, big); break;
				case 7: r = launch_rows<7>(ctx, rp, fuse, wgrid, big); break;
				default: r = launch_rows<8>(ctx, rp, fuse, wgrid, big); break;
			}
			if (r) return r;
			CU(cudaEventRecord(ctx->kev[1], st));
			if (cli) LAUNCH(ctx, k_correct_cli, wgrid, 256, rp);
			CU(cudaEventRecord(ctx->ev[4], st));
			if (fuse_final && !fuse) { int r2 = run_final(ctx); if (r2) return r2; }
			CU(cudaEventRecord(ctx->ev[5], st));
			CU(cudaStreamSynchronize(st));
			rp.gate = nullptr;
			if (hs[63] == 1) {
				generic_done = true;
				R = hs[25];
				ctx->any_big = big;
				for (int c = 1; c < nc; c++) ctx->stats[1] += (int64_t) hs[16 + c];
			}
		}
	}
	if (generic_done) {
		ctx->res_ncat = nc;
		ctx->timing_dirty = true;
		ctx->timing_ncat = nc;
		ctx->nrows = R;
		ctx->matched = true;
		ctx->finalized = fuse_final != 0;
		if (nrows) *nrows = R;
		if (R == 0) return fail(ctx, NWB_ERR_EMPTY, "No matches.");
		return NWB_OK;
	}

	if (generic) {
		// ---- lists: N >= 3 (and the elliptical mode) need the matches sorted and compact -----------------------------------------
		Lists L;
		memset(&L, 0, sizeof(L));
		for (int c = 1; c < nc; c++) {
			ENSURE(ctx->d_segoff[c], (size_t) (np + 1) * sizeof(long long));
			long long *off = (long long *) ctx->d_segoff[c].p;
			{ int r = scan_int_to_ll(ctx, (const int *) d_cnt[c], off, np + 1); if (r) return r; }   // cnt[np] == 0
			CU(cudaMemcpyAsync(hs + 16 + c, off + np, sizeof(long long), cudaMemcpyDeviceToHost, st));
			// largest match count of any primary: decides whether the warp-per-primary sort is needed at all
			int *d_maxcnt = (int *) (d_status + 40) + c;
			CU(cudaMemsetAsync(d_maxcnt, 0, sizeof(int), st));
			LAUNCH(ctx, k_max_int, (int) std::min<int64_t>((np + 255) / 256, 148 * 8), 256, (long long) np, (const int *) d_cnt[c], d_maxcnt);
		}
		CU(cudaMemcpyAsync(hs + 40, d_status + 40, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		const int *h_maxcnt = (const int *) (hs + 40);
		for (int c = 1; c < nc; c++) {
			size_t npair = (size_t) hs[16 + c];
			ctx->stats[1] += (int64_t) npair;
			ENSURE(ctx->d_Ls[c], std::max<size_t>(1, npair) * sizeof(int));
			ENSURE(ctx->d_Lsep[c], std::max<size_t>(1, npair) * sizeof(double));
			ENSURE(ctx->d_Ltrig[c], std::max<size_t>(1, npair) * 4 * sizeof(double));   // lon | sin lat | cos lat | flat-sky cell
			double *tr = (double *) ctx->d_Ltrig[c].p;
			const long long *off = (const long long *) ctx->d_segoff[c].p;
			if (npair) {
				LAUNCH(ctx, k_sort_lists_small, pblocks, 256, (int) np, stores[c], off, (int *) ctx->d_Ls[c].p, (double *) ctx->d_Lsep[c].p,
					ctx->cat[c].ra, ctx->cat[c].dec, tr, tr + npair, tr + 2 * npair, (long long *) (tr + 3 * npair), flat);
				if (h_maxcnt[c] > SMALL_N)
					LAUNCH(ctx, k_sort_lists, wgrid, 256, (int) np, stores[c], off, (int *) ctx->d_Ls[c].p, (double *) ctx->d_Lsep[c].p,
						ctx->cat[c].ra, ctx->cat[c].dec, tr, tr + npair, tr + 2 * npair, SMALL_N, (long long *) (tr + 3 * npair), flat);
			}
			L.off[c] = off;
			L.s[c] = (const int *) ctx->d_Ls[c].p;
			L.sep[c] = (const double *) ctx->d_Lsep[c].p;
			L.lon[c] = tr; L.slat[c] = tr + npair; L.clat[c] = tr + 2 * npair;
			L.ij[c] = (const long long *) (tr + 3 * npair);
		}
		rp.L = L;
		CU(cudaEventRecord(ctx->ev[3], st));
		ENSURE(ctx->d_matsz, (size_t) (np + 1) * sizeof(long long));
		ENSURE(ctx->d_matoff, (size_t) (np + 1) * sizeof(long long));
		long long *d_matsz = (long long *) ctx->d_matsz.p, *d_matoff = (long long *) ctx->d_matoff.p;
		CU(cudaMemsetAsync(d_matsz, 0, (size_t) (np + 1) * sizeof(long long), st));
		unsigned long long *d_maxtup = (unsigned long long *) (d_status + 44);
		CU(cudaMemsetAsync(d_maxtup, 0, sizeof(unsigned long long), st));
		LAUNCH(ctx, k_mat_sizes, pblocks, 256, (int) np, nc, L, d_matsz, d_maxtup);
		{ int r = scan_ll(ctx, d_matsz, d_matoff, np + 1); if (r) return r; }
		CU(cudaMemcpyAsync(hs + 24, d_matoff + np, sizeof(long long), cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(hs + 26, d_maxtup, sizeof(long long), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		long long mat_total = hs[24];
		const bool any_big = (unsigned long long) hs[26] > (unsigned long long) SMALL_T;   // else every primary is a "small" one
		ctx->any_big = any_big;
		ENSURE(ctx->d_mat, std::max<size_t>(1, (size_t) mat_total) * sizeof(double));
		rp.mat_off = d_matoff;
		rp.mat = (double *) ctx->d_mat.p;
		CU(cudaMemsetAsync(d_rows, 0, (size_t) (np + 1) * sizeof(long long), st));
		int r = 0;
		switch (nc) {
			case 2: r = launch_count<2>(ctx, rp, d_rows, wgrid, any_big); break;
			case 3: r = launch_count<3>(ctx, rp, d_rows, wgrid, any_big); break;
			case 4: r = launch_count<4>(ctx, rp, d_rows, wgrid, any_big); break;
			case 5: r = launch_count<5>(ctx, rp, d_rows, wgrid, any_big); break;
			case 6: r = launch_count<6>(ctx, rp, d_rows, wgrid, any_big); break;
			case 7: r = launch_count<7>(ctx, rp, d_rows, wgrid, any_big); break;
			default: r = launch_count<8>(ctx, rp, d_rows, wgrid, any_big); break;
		}
		if (r) return r;
		{ int r2 = scan_ll(ctx, d_rows, d_rowoff, np + 1); if (r2) return r2; }
		CU(cudaMemcpyAsync(hs + 25, d_rowoff + np, sizeof(long long), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		R = hs[25];
		// what the next match of this shape may assume (speculative pipeline above)
		ctx->gen.valid = true; ctx->gen.nc = nc; ctx->gen.ell = (int) ell; ctx->gen.np = np;
		ctx->gen.cap_mat = (long long) (ctx->d_mat.cap / sizeof(double));
		ctx->gen.any_big = any_big ? 1 : 0;
		for (int c = 1; c < nc; c++) {
			ctx->gen.cap_list[c] = (long long) std::min(std::min(ctx->d_Ls[c].cap / sizeof(int), ctx->d_Lsep[c].cap / sizeof(double)), ctx->d_Ltrig[c].cap / (4 * sizeof(double)));
			ctx->gen.big_sort[c] = h_maxcnt[c] > SMALL_N ? 1 : 0;
		}
	} else {
		ctx->stats[1] = R - np;
		if (!speculated) CU(cudaEventRecord(ctx->ev[3], st));
	}

	// ---- K2: rows ------------------------------------------------------------------------------------------
	ctx->res_ncat = nc;
	if (!speculated) {
		rp.row_off = d_rowoff;
		// keep the capacity of the previous table (grow-only) so that the next match can launch K2 speculatively
		int64_t cap_rows = std::max<int64_t>(R + R / 16 + 1024, ctx->cols_cap_rows);
		if (ctx->cols_cap_ncols != nc + nc * (nc - 1) / 2 + 9 + ctx->res_nmag) cap_rows = R + R / 16 + 1024;
		{ int r = layout_columns(ctx, ctx->d_cols, cap_rows, nc, ctx->res_nmag, ctx->cols, ctx->ncols); if (r) return r; }
		ctx->cols_cap_rows = cap_rows;
		ctx->cols_cap_ncols = ctx->ncols;
		rp.C = ctx->cols;
		CU(cudaEventRecord(ctx->kev[0], st));
		int r = 0;
		switch (nc) {
			case 2: {
				if (generic) { r = launch_rows<2>(ctx, rp, fuse, wgrid, ctx->any_big); break; }
				int grid2 = (int) std::min<int64_t>((np + R2_WARPS - 1) / R2_WARPS, 148 * NWB_R2_GRID_PER_SM);
				if (fuse && ctx->res_nmag == 0 && NWB_R2_SHARE) LAUNCH(ctx, (k_rows2<true, true>), grid2, R2_WARPS * 32, rp);
				else if (fuse) LAUNCH(ctx, (k_rows2<true, false>), grid2, R2_WARPS * 32, rp);
				else LAUNCH(ctx, (k_rows2<false, false>), grid2, R2_WARPS * 32, rp);
				break;
			}
			case 3: r = launch_rows<3>(ctx, rp, fuse, wgrid, ctx->any_big); break;
			case 4: r = launch_rows<4>(ctx, rp, fuse, wgrid, ctx->any_big); break;
			case 5: r = launch_rows<5>(ctx, rp, fuse, wgrid, ctx->any_big); break;
			case 6: r = launch_rows<6>(ctx, rp, fuse, wgrid, ctx->any_big); break;
			case 7: r = launch_rows<7>(ctx, rp, fuse, wgrid, ctx->any_big); break;
			default: r = launch_rows<8>(ctx, rp, fuse, wgrid, ctx->any_big); break;
		}
		if (r) return r;
		CU(cudaEventRecord(ctx->kev[1], st));
	}
	rp.guard = nullptr;   // k_final / later calls are not speculative
	if (cli) LAUNCH(ctx, k_correct_cli, wgrid, 256, rp);
	CU(cudaEventRecord(ctx->ev[4], st));
	if (fuse_final && !fuse) { int r = run_final(ctx); if (r) return r; }
	CU(cudaEventRecord(ctx->ev[5], st));
	CU(cudaStreamSynchronize(st));
	ctx->timing_dirty = true;   // elapsed times are read from the events on demand (nwb_timing)
	ctx->timing_ncat = nc;
	ctx->nrows = R;
	ctx->matched = true;
	ctx->finalized = fuse_final != 0;
	if (nrows) *nrows = R;
	if (R == 0) return fail(ctx, NWB_ERR_EMPTY, "No matches.");
	return NWB_OK;
}

// ---- shard mode: streaming split by secondary rows, matches scattered to the owners over peer memory --------------
int nwb_shard_setup(nwb_ctx *ctx, int rank, int world, int64_t spill_capacity, void *ipc_handle_out, int64_t *exchange_bytes)
{
	if (!ctx) return NWB_ERR_ARG;
	if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(ctx, NWB_ERR_ARG, "bad rank / world (at most 16 ranks)");
	if (world > 1 && !ipc_handle_out) return fail(ctx, NWB_ERR_ARG, "ipc_handle_out is NULL");
	const int nc = ctx->ncat;
	if (nc < 2 || !ctx->params_set) return fail(ctx, NWB_ERR_STATE, "set the catalogues and nwb_set_params first");
	for (int c = 0; c < nc; c++)
		if (!ctx->cat[c].set) return fail(ctx, NWB_ERR_STATE, "catalogue " + std::to_string(c) + " not set");
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	shard_release(ctx);
	nwb_ctx::Shard &S = ctx->shard;
	S.rank = rank; S.world = world;
	const int64_t n0 = ctx->cat[0].n;
	S.block = std::max<int64_t>(1, (n0 + world - 1) / world);
	S.spill_cap = (unsigned long long) std::max<int64_t>(spill_capacity, 4096);
	// layout, identical on every rank: spill counters | match counters per catalogue | slots per catalogue | spill records
	size_t off = 0;
	auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
	const size_t sc = take(MAXC * sizeof(unsigned long long));
	for (int c = 0; c < MAXC; c++) S.off_spillcnt[c] = sc + (size_t) c * sizeof(unsigned long long);
	for (int c = 1; c < nc; c++) S.off_cnt[c] = take((size_t) (S.block + 4) * sizeof(int));
	S.zero_bytes = off;
	for (int c = 1; c < nc; c++) {
		S.C[c] = slots_per_primary(ctx, c, S.block);
		S.off_slot[c] = take((size_t) S.block * S.C[c] * sizeof(Slot16));
	}
	for (int c = 1; c < nc; c++) S.off_spill[c] = take((size_t) S.spill_cap * sizeof(SpillRec));
	S.set_bytes = off;
	S.off_bar = 2 * off;
	S.bytes = 2 * off + 256;
	S.epoch = 0;
	S.bar_tag = 0;
	// a plain cudaMalloc of its own: IPC handles cover whole allocations
	{ int r = ensure(ctx, S.xch, S.bytes); if (r) return r; }
	CU(cudaMemsetAsync(S.xch.p, 0, S.zero_bytes, ctx->stream));
	CU(cudaMemsetAsync((char *) S.xch.p + S.set_bytes, 0, S.zero_bytes, ctx->stream));
	CU(cudaMemsetAsync((char *) S.xch.p + S.off_bar, 0, 256, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (ipc_handle_out) {
		cudaIpcMemHandle_t h;
		memset(&h, 0, sizeof(h));
		if (world > 1) CU(cudaIpcGetMemHandle(&h, S.xch.p));
		memcpy(ipc_handle_out, &h, sizeof(h));
	}
	if (exchange_bytes) *exchange_bytes = (int64_t) S.bytes;
	S.on = true;
	S.connected = false;
	ctx->geom_valid = false;
	ctx->matched = ctx->finalized = false;
	return NWB_OK;
}

int nwb_shard_connect(nwb_ctx *ctx, const void *ipc_handles)
{
	if (!ctx) return NWB_ERR_ARG;
	nwb_ctx::Shard &S = ctx->shard;
	if (!S.on) return fail(ctx, NWB_ERR_STATE, "nwb_shard_setup first");
	if (S.world > 1 && !ipc_handles) return fail(ctx, NWB_ERR_ARG, "ipc_handles is NULL");
	CU(cudaSetDevice(ctx->device));
	for (int r = 0; r < S.world; r++) {
		if (r == S.rank) { S.peer[r] = S.xch.p; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *) ipc_handles + (size_t) r * sizeof(h), sizeof(h));
		void *p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			cudaGetLastError();
			return fail(ctx, NWB_ERR_CUDA, "cudaIpcOpenMemHandle for rank " + std::to_string(r) + ": " + cudaGetErrorString(e) +
				" (shard mode needs peer access between the GPUs of one node)");
		}
		S.peer[r] = p;
	}
	{ int r = ensure(ctx, S.d_peers, 16 * sizeof(void *)); if (r) return r; }
	CU(cudaMemcpyAsync(S.d_peers.p, S.peer, 16 * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	S.connected = true;
	return NWB_OK;
}

int nwb_shard_close(nwb_ctx *ctx)
{
	if (!ctx) return NWB_ERR_ARG;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	shard_release(ctx);
	ctx->geom_valid = false;
	return NWB_OK;
}

int nwb_shard_match(nwb_ctx *ctx, int phase, int fuse_final, int64_t *nrows)
{
	if (!ctx) return NWB_ERR_ARG;
	nwb_ctx::Shard &S = ctx->shard;
	if (!S.on || !S.connected) return fail(ctx, NWB_ERR_STATE, "nwb_shard_setup / nwb_shard_connect first");
	CU(cudaSetDevice(ctx->device));
	if (phase == 0) {
		// both sets of counters back to zero (after an aborted match; a barrier between the ranks must follow).  Not part
		// of the steady state: every match zeroes the OTHER set for its successor
		LAUNCH(ctx, k_zero, (int) std::min<size_t>((S.zero_bytes / 16 + 255) / 256, 148 * 4), 256, (int4 *) S.xch.p, (long long) (S.zero_bytes / 16),
			(unsigned long long *) S.xch.p, 0);
		LAUNCH(ctx, k_zero, (int) std::min<size_t>((S.zero_bytes / 16 + 255) / 256, 148 * 4), 256, (int4 *) ((char *) S.xch.p + S.set_bytes), (long long) (S.zero_bytes / 16),
			(unsigned long long *) S.xch.p, 0);
		ctx->matched = ctx->finalized = false;
		return NWB_OK;
	}
	if (phase == 1) {
		// the set the NEXT match will use: nobody reads it any more (this rank's rows of the previous match are stream-ordered
		// before this launch) and no peer writes it before the barrier that follows this phase
		char *next = (char *) S.xch.p + (size_t) ((S.epoch + 1) & 1) * S.set_bytes;
		LAUNCH(ctx, k_zero, (int) std::min<size_t>((S.zero_bytes / 16 + 255) / 256, 148 * 4), 256, (int4 *) next, (long long) (S.zero_bytes / 16),
			(unsigned long long *) next, 0);
		return match_impl(ctx, fuse_final, nullptr, true, false, 1);
	}
	if (phase == 2) {
		const int r = match_impl(ctx, fuse_final, nrows, true, false, 2);   // 1 = every rank has to repeat the match (phases 1, 2)
		S.epoch++;   // the next match (or the repetition) uses the other set
		return r;
	}
	if (phase == 3) {
		// both halves with the barrier between them as flags in peer memory (k_peer_sync): no collective library, no
		// return to the caller in between
		int r = nwb_shard_match(ctx, 1, fuse_final, nullptr);
		if (r) return r;
		if (!ctx->h_status) CU(cudaHostAlloc((void **) &ctx->h_status, 64 * sizeof(long long), cudaHostAllocMapped));
		ctx->h_status[45] = 0;
		PeerSync B;
		for (int k = 0; k < 16; k++) B.peers[k] = k < S.world ? (char *) S.peer[k] : nullptr;
		B.tag_off = (long long) S.off_bar; B.payload_off = -1; B.tag = ++S.bar_tag; B.payload = 0;
		B.collect_dev = nullptr; B.collect_host = ctx->h_status + 45 - 16;   // the error word lands in h_status[45]
		B.timeout_ns = PEER_TIMEOUT_NS; B.world = S.world; B.rank = S.rank;
		LAUNCH(ctx, k_peer_sync, 1, 32, B);
		r = nwb_shard_match(ctx, 2, fuse_final, nrows);
		if (ctx->h_status[45] != 0) return fail(ctx, NWB_ERR_STATE, "shard mode: a rank did not reach the barrier within 10 s");
		return r;
	}
	return fail(ctx, NWB_ERR_ARG, "phase must be 0, 1, 2 or 3");
}

// ---- table all-gather over peer memory ---------------------------------------------------------------------------
int nwb_gather_setup(nwb_ctx *ctx, int rank, int world, int64_t capacity_rows, int ncols, void *ipc_handle_out)
{
	if (!ctx) return NWB_ERR_ARG;
	if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(ctx, NWB_ERR_ARG, "bad rank / world (at most 16 ranks)");
	if (capacity_rows < 0 || ncols < 1) return fail(ctx, NWB_ERR_ARG, "capacity_rows < 0 or ncols < 1");
	if (world > 1 && !ipc_handle_out) return fail(ctx, NWB_ERR_ARG, "ipc_handle_out is NULL");
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	gather_release(ctx);
	nwb_ctx::Gather &T = ctx->gather;
	T.rank = rank; T.world = world;
	T.ncols = ncols;
	T.cap_rows = std::max<int64_t>(capacity_rows, 32);
	T.stride = ((size_t) T.cap_rows * 8 + 255) / 256 * 256;
	T.set_bytes = T.stride * ncols;
	T.epoch = 0;
	{ int r = ensure(ctx, T.buf, GATHER_HEADER + 2 * T.set_bytes); if (r) return r; }   // a cudaMalloc of its own: IPC handles cover whole allocations
	{ int r = ensure(ctx, T.d_counts, 17 * sizeof(long long)); if (r) return r; }
	CU(cudaMemsetAsync(T.buf.p, 0, GATHER_HEADER, ctx->stream));
	CU(cudaMemsetAsync(T.d_counts.p, 0, 17 * sizeof(long long), ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (!T.h_words) CU(cudaHostAlloc((void **) &T.h_words, 32 * sizeof(long long), cudaHostAllocMapped));
	memset(T.h_words, 0, 32 * sizeof(long long));
	if (ipc_handle_out) {
		cudaIpcMemHandle_t h;
		memset(&h, 0, sizeof(h));
		if (world > 1) CU(cudaIpcGetMemHandle(&h, T.buf.p));
		memcpy(ipc_handle_out, &h, sizeof(h));
	}
	T.on = true;
	T.connected = false;
	return NWB_OK;
}

int nwb_gather_connect(nwb_ctx *ctx, const void *ipc_handles)
{
	if (!ctx) return NWB_ERR_ARG;
	nwb_ctx::Gather &T = ctx->gather;
	if (!T.on) return fail(ctx, NWB_ERR_STATE, "nwb_gather_setup first");
	if (T.world > 1 && !ipc_handles) return fail(ctx, NWB_ERR_ARG, "ipc_handles is NULL");
	CU(cudaSetDevice(ctx->device));
	for (int r = 0; r < T.world; r++) {
		if (r == T.rank) { T.peer[r] = T.buf.p; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *) ipc_handles + (size_t) r * sizeof(h), sizeof(h));
		void *p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			cudaGetLastError();
			return fail(ctx, NWB_ERR_CUDA, "cudaIpcOpenMemHandle for rank " + std::to_string(r) + ": " + cudaGetErrorString(e) +
				" (the peer-memory all-gather needs peer access between the GPUs of one node)");
		}
		T.peer[r] = p;
	}
	T.connected = true;
	return NWB_OK;
}

int nwb_gather_close(nwb_ctx *ctx)
{
	if (!ctx) return NWB_ERR_ARG;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	gather_release(ctx);
	return NWB_OK;
}

int nwb_gather_push(nwb_ctx *ctx, const int64_t *counts, int counts_on_device, int engine, void **table, int64_t *stride_bytes)
{
	if (!ctx || (!counts && engine != 2)) return NWB_ERR_ARG;
	nwb_ctx::Gather &T = ctx->gather;
	if (!T.on || !T.connected) return fail(ctx, NWB_ERR_STATE, "nwb_gather_setup / nwb_gather_connect first");
	if (!ctx->matched && !ctx->pending) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	if (ctx->ncols != T.ncols) return fail(ctx, NWB_ERR_ARG, "the table has " + std::to_string(ctx->ncols) + " columns, nwb_gather_setup was told " + std::to_string(T.ncols));
	if (counts_on_device && engine != 0) return fail(ctx, NWB_ERR_ARG, "the copy engines need the counts on the host");
	if (engine == 2 && ctx->pending) return fail(ctx, NWB_ERR_STATE, "engine 2 publishes the row count of a finished match: nwb_match_wait first");
	int64_t total = 0, off = 0;
	if (engine == 2) {
		counts = (const int64_t *) T.d_counts.p;   // filled by the count exchange below
		counts_on_device = 1;
	} else if (!counts_on_device) {
		for (int r = 0; r < T.world; r++) {
			if (counts[r] < 0) return fail(ctx, NWB_ERR_ARG, "negative row count");
			if (r < T.rank) off += counts[r];
			total += counts[r];
		}
		if (!ctx->pending && counts[T.rank] != ctx->nrows) return fail(ctx, NWB_ERR_ARG, "counts[rank] is not the row count of the last match");
		if (total > T.cap_rows)
			return fail(ctx, NWB_ERR_NOMEM, "the gathered table (" + std::to_string(total) + " rows) exceeds the capacity given to nwb_gather_setup");
	}
	CU(cudaSetDevice(ctx->device));
	const size_t set = GATHER_HEADER + (size_t) (T.epoch & 1) * T.set_bytes;
	T.epoch++;
	if (table) *table = (char *) T.buf.p + set;
	if (stride_bytes) *stride_bytes = (int64_t) T.stride;
	const int ncols = T.ncols;
	const long long src_stride = (long long) (((size_t) std::max<int64_t>(ctx->cols_cap_rows, 1) * 8 + 255) / 256 * 256);
	PeerSync Y;
	if (engine == 2) {
		// the row counts travel as flag words: every rank writes its count into every rank's header and waits for the others'
		for (int k = 0; k < 16; k++) Y.peers[k] = k < T.world ? (char *) T.peer[k] : nullptr;
		Y.tag_off = (long long) GATHER_OFF_COUNT_TAG; Y.payload_off = (long long) GATHER_OFF_COUNT;
		Y.tag = (unsigned long long) T.epoch; Y.payload = ctx->nrows;   // T.epoch was incremented above: tags start at 1
		Y.collect_dev = (long long *) T.d_counts.p; Y.collect_host = T.h_words;
		Y.timeout_ns = PEER_TIMEOUT_NS; Y.world = T.world; Y.rank = T.rank;
		LAUNCH(ctx, k_peer_sync, 1, 32, Y);
	}
	if (engine == 0 || engine == 2) {
		PushArgs A;
		A.src = (const char *) ctx->d_cols.p; A.src_stride = src_stride;
		A.counts = counts_on_device ? (const long long *) counts : nullptr;
		for (int r = 0; r < 16; r++) A.counts_val[r] = (!counts_on_device && r < T.world) ? counts[r] : 0;
		A.cap_rows = T.cap_rows;
		const char *envc = getenv("NWB_PUSH_CHUNK");   // measurement knob
		A.chunk = envc && atoll(envc) >= 512 ? atoll(envc) / 16 * 16 : 0;
		A.ncols = ncols; A.world = T.world; A.rank = T.rank;
		for (int r = 0; r < 16; r++) A.dst[r] = r < T.world ? (char *) T.peer[r] + set : nullptr;
		A.dst_stride = (long long) T.stride;
		A.abort = engine == 2 ? (const long long *) T.d_counts.p + 16 : nullptr;
		if (ctx->num_sms <= 0) { int nsm = 0; CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device)); ctx->num_sms = std::max(nsm, 1); }
		// enough blocks for every link to stay busy, few enough to leave after a small table at once
		const char *env = getenv("NWB_PUSH_BLOCKS_PER_SM");   // measurement knob (tools/bench_push.py)
		const int per_sm = env && atoi(env) > 0 && atoi(env) <= 4 ? atoi(env) : 4;
		long long grid = (long long) ctx->num_sms * per_sm;
		if (!counts_on_device) grid = std::max<long long>(1, std::min<long long>(grid, (counts[T.rank] + 2047) / 2048 * ncols * T.world));
		const char *envs = getenv("NWB_PUSH_STORE");
		const int flavour = envs ? atoi(envs) : 1;
		if (flavour == 1) LAUNCH(ctx, k_table_push<1>, (int) grid, PUSH_THREADS, A);
		else if (flavour == 2) LAUNCH(ctx, k_table_push<2>, (int) grid, PUSH_THREADS, A);
		else LAUNCH(ctx, k_table_push<0>, (int) grid, PUSH_THREADS, A);
		if (engine == 2) {
			// ... and so does "my push has landed": the table is complete when this kernel has seen every rank's tag
			Y.tag_off = (long long) GATHER_OFF_DONE_TAG; Y.payload_off = -1;
			Y.collect_dev = (long long *) T.d_counts.p;   // only its abort word is touched
			LAUNCH(ctx, k_peer_sync, 1, 32, Y);
		}
		return NWB_OK;
	}
	const int64_t rows = counts[T.rank];
	if (rows == 0) return NWB_OK;
	// copy engines: one stream per destination, ncols copies each, joined back into the context's stream
	if (!T.fork) {
		CU(cudaEventCreateWithFlags(&T.fork, cudaEventDisableTiming));
		for (int r = 0; r < T.world; r++) {
			CU(cudaStreamCreateWithFlags(&T.lane[r], cudaStreamNonBlocking));
			CU(cudaEventCreateWithFlags(&T.join[r], cudaEventDisableTiming));
		}
	}
	CU(cudaEventRecord(T.fork, ctx->stream));
	for (int i = 0; i < T.world; i++) {
		const int r = (T.rank + 1 + i) % T.world;
		CU(cudaStreamWaitEvent(T.lane[r], T.fork, 0));
		for (int k = 0; k < ncols; k++)
			CU(cudaMemcpyAsync((char *) T.peer[r] + set + (size_t) k * T.stride + (size_t) off * 8,
				(const char *) ctx->d_cols.p + (size_t) k * src_stride, (size_t) rows * 8, cudaMemcpyDeviceToDevice, T.lane[r]));
		CU(cudaEventRecord(T.join[r], T.lane[r]));
		CU(cudaStreamWaitEvent(ctx->stream, T.join[r], 0));
	}
	return NWB_OK;
}

int nwb_gather_counts(nwb_ctx *ctx, int64_t *counts)
{
	if (!ctx || !counts) return NWB_ERR_ARG;
	nwb_ctx::Gather &T = ctx->gather;
	if (!T.on || !T.connected || !T.h_words) return fail(ctx, NWB_ERR_STATE, "nwb_gather_setup / nwb_gather_connect first");
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	if (T.h_words[16] != 0) return fail(ctx, NWB_ERR_STATE, "table gather: a rank did not arrive within 10 s");
	int64_t total = 0;
	for (int r = 0; r < T.world; r++) { counts[r] = T.h_words[r]; total += counts[r]; }
	if (total > T.cap_rows)
		return fail(ctx, NWB_ERR_NOMEM, "the gathered table (" + std::to_string(total) + " rows) exceeds the capacity given to nwb_gather_setup");
	return NWB_OK;
}

int nwb_match(nwb_ctx *ctx, int fuse_final, int64_t *nrows)
{
	if (!ctx) return NWB_ERR_ARG;
	int r = match_impl(ctx, fuse_final, nrows, true, false);
	if (r == 1) r = match_impl(ctx, fuse_final, nrows, false, false);   // the cached grid geometry did not fit: rebuild it
	return r;
}

int nwb_match_async(nwb_ctx *ctx, int fuse_final)
{
	if (!ctx) return NWB_ERR_ARG;
	int r = match_impl(ctx, fuse_final, nullptr, true, true);
	if (r == 1) r = match_impl(ctx, fuse_final, nullptr, false, false);
	return r;
}

int nwb_match_wait(nwb_ctx *ctx, int64_t *nrows)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!ctx->pending) {   // nwb_match_async ran the match synchronously (first match of a context, N >= 3, ...)
		if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match_wait: no match in flight");
		if (nrows) *nrows = ctx->nrows;
		return ctx->nrows == 0 ? fail(ctx, NWB_ERR_EMPTY, "No matches.") : NWB_OK;
	}
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->pending = false;
	const long long *hs = ctx->h_status;
	const long long R = hs[8];
	// the conditions the row kernel checked on the device (guard_ok): geometry still valid, overflow entries fit,
	// nothing spilled, the table of the previous match is big enough
	const bool ok = hs[9] == 0 && (size_t) hs[0] <= ctx->entries_cap && hs[1] == 0 && R <= ctx->cols_cap_rows;
	if (!ok) return nwb_match(ctx, ctx->pending_fuse, nrows);
	ctx->stats[1] = R - ctx->np;
	ctx->stats[2] = ctx->geom_G.ncells;
	ctx->stats[3] = hs[10];
	ctx->res_ncat = ctx->ncat;
	ctx->rp.guard = nullptr;
	ctx->timing_dirty = true;
	ctx->timing_ncat = ctx->ncat;
	ctx->nrows = R;
	ctx->matched = true;
	ctx->finalized = ctx->pending_fuse != 0;
	if (nrows) *nrows = R;
	if (R == 0) return fail(ctx, NWB_ERR_EMPTY, "No matches.");
	return NWB_OK;
}

static int refresh_timings(nwb_ctx *ctx)
{
	if (!ctx->timing_dirty) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	for (int k = 0; k < 5; k++) CU(cudaEventElapsedTime(&ctx->ms[k], ctx->ev[k], ctx->ev[k + 1]));
	CU(cudaEventElapsedTime(&ctx->ms[NWB_T_TOTAL], ctx->ev[0], ctx->ev[5]));
	ctx->ms[NWB_T_KPAIRS] = 0;
	for (int c = 1; c < ctx->timing_ncat; c++) {
		float t = 0;
		if (ctx->cat[c].n > 0) CU(cudaEventElapsedTime(&t, ctx->kev[2 * c], ctx->kev[2 * c + 1]));
		ctx->ms[NWB_T_KPAIRS] += t;
	}
	CU(cudaEventElapsedTime(&ctx->ms[NWB_T_KROWS], ctx->kev[0], ctx->kev[1]));
	ctx->timing_dirty = false;
	return NWB_OK;
}

int nwb_finalize(nwb_ctx *ctx)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	CU(cudaSetDevice(ctx->device));
	{ int r = upload_tables(ctx); if (r) return r; }   // magnitude priors may have been set since nwb_match
	ctx->rp.nmag = ctx->res_nmag;
	CU(cudaEventRecord(ctx->ev[6], ctx->stream));
	{ int r = run_final(ctx); if (r) return r; }
	CU(cudaEventRecord(ctx->ev[7], ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	{ int r = refresh_timings(ctx); if (r) return r; }
	CU(cudaEventElapsedTime(&ctx->ms[NWB_T_FINAL], ctx->ev[6], ctx->ev[7]));
	ctx->finalized = true;
	return NWB_OK;
}

int nwb_truncate(nwb_ctx *ctx, double min_prob, int64_t *nrows)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!ctx->finalized) return fail(ctx, NWB_ERR_STATE, "nwb_finalize has not run");
	CU(cudaSetDevice(ctx->device));
	int64_t R = ctx->nrows;
	if (R == 0 || !(min_prob > 0)) { if (nrows) *nrows = R; return NWB_OK; }
	cudaStream_t st = ctx->stream;
	ENSURE(ctx->d_keep, (size_t) (R + 1) * sizeof(int));
	ENSURE(ctx->d_keeppos, (size_t) (R + 1) * sizeof(long long));
	int *keep = (int *) ctx->d_keep.p;
	long long *pos = (long long *) ctx->d_keeppos.p;
	CU(cudaMemsetAsync(keep + R, 0, sizeof(int), st));
	LAUNCH(ctx, k_keep_flags, grid_for(R, 256), 256, (long long) R, (const double *) ctx->cols.p_i, min_prob, keep);
	{ int r = scan_int_to_ll(ctx, keep, pos, R + 1); if (r) return r; }
	long long R2 = 0;
	CU(cudaMemcpyAsync(&R2, pos + R, sizeof(long long), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	Columns C2;
	memset(&C2, 0, sizeof(C2));
	int ncols2 = 0;
	{ int r = layout_columns(ctx, ctx->d_cols2, R2, ctx->res_ncat, ctx->res_nmag, C2, ncols2); if (r) return r; }
	size_t s1 = ((size_t) std::max<int64_t>(ctx->cols_cap_rows, 1) * 8 + 255) / 256 * 256, s2 = ((size_t) std::max<int64_t>(R2, 1) * 8 + 255) / 256 * 256;
	for (int k = 0; k < ctx->ncols; k++)
		LAUNCH(ctx, k_compact8, grid_for(R, 256), 256, (long long) R, (const int *) keep, (const long long *) pos,
			(const unsigned long long *) ((char *) ctx->d_cols.p + s1 * k), (unsigned long long *) ((char *) ctx->d_cols2.p + s2 * k));
	CU(cudaStreamSynchronize(st));
	std::swap(ctx->d_cols, ctx->d_cols2);
	ctx->cols_cap_rows = std::max<int64_t>(R2, 1);
	ctx->cols = C2;
	ctx->rp.C = C2;
	ctx->nrows = R2;
	if (nrows) *nrows = R2;
	return NWB_OK;
}

int nwb_column_ptr(nwb_ctx *ctx, int column, void **dev_ptr)
{
	if (!ctx || !dev_ptr) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	return column_lookup(ctx, column, dev_ptr);
}

int nwb_table_layout(nwb_ctx *ctx, void **base, int64_t *stride_bytes, int *ncols, int64_t *nrows)
{
	if (!ctx || !base || !stride_bytes || !ncols || !nrows) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	*base = ctx->d_cols.p;
	*stride_bytes = (int64_t) (((size_t) std::max<int64_t>(ctx->cols_cap_rows, 1) * 8 + 255) / 256 * 256);
	*ncols = ctx->ncols;
	*nrows = ctx->nrows;
	return NWB_OK;
}

int nwb_fetch(nwb_ctx *ctx, int column, void *dst_host)
{
	if (!ctx || !dst_host) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	void *p = nullptr;
	int r = column_lookup(ctx, column, &p);
	if (r) return r;
	CU(cudaSetDevice(ctx->device));
	if (ctx->nrows > 0)
		CU(cudaMemcpyAsync(dst_host, p, (size_t) ctx->nrows * 8, cudaMemcpyDeviceToHost, ctx->stream));
	return NWB_OK;
}

int nwb_nrows_device_ptr(nwb_ctx *ctx, void **dev_ptr)
{
	if (!ctx || !dev_ptr) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	*dev_ptr = (void *) ((long long *) ctx->d_rowoff.p + ctx->np);   // row_off[np] == R
	return NWB_OK;
}

int nwb_fetch_device(nwb_ctx *ctx, int column, void *dst_device)
{
	if (!ctx || !dst_device) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	void *p = nullptr;
	int r = column_lookup(ctx, column, &p);
	if (r) return r;
	CU(cudaSetDevice(ctx->device));
	if (ctx->nrows > 0)
		CU(cudaMemcpyAsync(dst_device, p, (size_t) ctx->nrows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
	return NWB_OK;
}

int nwb_timing(nwb_ctx *ctx, int stage, float *ms)
{
	if (!ctx || !ms || stage < 0 || stage >= NWB_T_COUNT) return NWB_ERR_ARG;
	{ int r = refresh_timings(ctx); if (r) return r; }
	*ms = ctx->ms[stage];
	return NWB_OK;
}

int nwb_bench_skeleton(nwb_ctx *ctx, int c, int reps, int blocks_per_sm, float *ms)
{
	if (!ctx || !ms) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	if (c < 1 || c >= ctx->res_ncat || ctx->cat[c].n == 0 || reps < 1) return fail(ctx, NWB_ERR_ARG, "bad catalogue / repetitions");
	CU(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const Grid G = ctx->geom_G;
	K1Args ka = ctx->last_k1[c];
	{
		// residency of the match kernel the skeleton is compared with (the instantiation the last match used)
		const int which = (ctx->last_k1_dense ? 1 : 0) | (ctx->flat_err > 0.0 ? 2 : 0);
		ctx->skel_blocks = blocks_per_sm > 0 ? blocks_per_sm : ctx->k1_occ[which];
	}
	// the skeleton takes its slots from a counter array of its own; the slot area it scribbles over is the match's
	ENSURE(ctx->d_misc, (size_t) (ctx->np + 1) * sizeof(int));
	ka.cnt = (int *) ctx->d_misc.p;
	ctx->matched = ctx->finalized = false;   // the pair store no longer holds the match
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0));
	CU(cudaEventCreate(&e1));
	float total = 0;
	for (int k = 0; k < reps + 1; k++) {
		CU(cudaMemsetAsync(ka.cnt, 0, (size_t) (ctx->np + 1) * sizeof(int), st));
		CU(cudaEventRecord(e0, st));
		{ int r = launch_pairs(ctx, ctx->last_k1_dense, false, true, false, (int) ctx->cat[c].n, ctx->cat[c].ra, ctx->cat[c].dec, G,
			(const int *) ctx->last_etotal, (const CellRec *) ctx->d_cells.p, (const Entry *) ctx->d_entries.p, (long long) ctx->entries_cap, ka); if (r) return r; }
		CU(cudaEventRecord(e1, st));
		CU(cudaStreamSynchronize(st));
		float t = 0;
		CU(cudaEventElapsedTime(&t, e0, e1));
		if (k > 0) total += t;   // the first repetition warms up
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*ms = total / reps;
	return NWB_OK;
}

int nwb_launch_count(nwb_ctx *ctx, int64_t *launches)
{
	if (!ctx || !launches) return NWB_ERR_ARG;
	*launches = ctx->launches;
	return NWB_OK;
}

int nwb_stats(nwb_ctx *ctx, int64_t *out4)
{
	if (!ctx || !out4) return NWB_ERR_ARG;
	for (int k = 0; k < 4; k++) out4[k] = ctx->stats[k];
	return NWB_OK;
}

// ---- N1: automatic magnitude histograms ---------------------------------------------------------------------
// the selection over the rows (res = index column of catalogue c, sepmax, dist_post: device columns of R rows) -- the
// context's own table, or the rows of ALL shards gathered by the caller (nwb_maghist_select_rows)
static int maghist_select_impl(nwb_ctx *ctx, int c, int k, int64_t R, const long long *res_col, const double *sepmax_col, const double *post_col,
	int by_radius, double thr_select, double thr_possible, int weights_cli, int64_t *nselected, int64_t *counts3, double *minmax2)
{
	if (c < 1 || c >= ctx->ncat) return fail(ctx, NWB_ERR_ARG, "catalogue index out of range");
	if (k < 0 || k >= ctx->cat[c].m) return fail(ctx, NWB_ERR_ARG, "magnitude column out of range");
	CU(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const int64_t n = ctx->cat[c].n;
	if (R >= 0x7f7f7f7fll) return fail(ctx, NWB_ERR_ARG, "table too long for the histogram selection");
	const double *mag = ctx->cat[c].mags + (size_t) k * n;
	// per-row scratch: flag bytes | isel | idef | selpos | defpos | W
	const size_t Rp = ((size_t) R + 63) / 64 * 64;
	ENSURE(ctx->d_hrow, Rp + 4 * Rp * sizeof(int) + Rp * sizeof(double) + 256);
	unsigned char *flag = (unsigned char *) ctx->d_hrow.p;
	int *isel = (int *) (flag + Rp), *idef = isel + Rp, *selpos = idef + Rp, *defpos = selpos + Rp;
	double *W = (double *) (defpos + Rp);
	// per-source scratch: first | selflag | selrank | possible bytes | stats
	const size_t np_ = ((size_t) n + 63) / 64 * 64;
	ENSURE(ctx->d_hsrc, 3 * (np_ + 64) * sizeof(int) + np_ + 64 + 8 * sizeof(unsigned long long));
	int *first = (int *) ctx->d_hsrc.p, *selflag = first + np_ + 64, *selrank = selflag + np_ + 64;
	unsigned char *possible = (unsigned char *) (selrank + np_ + 64);
	unsigned long long *stats = (unsigned long long *) (possible + np_ + 64);
	CU(cudaMemsetAsync(first, 0x7f, (np_ + 64) * sizeof(int), st));      // 0x7f7f7f7f: any real position is smaller
	CU(cudaMemsetAsync(possible, 0, np_ + 64, st));
	CU(cudaMemsetAsync(selflag, 0, (np_ + 64) * sizeof(int), st));
	unsigned long long init[5] = {0, 0, 0, ~0ull, 0};
	CU(cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
	const double *sw = by_radius ? nullptr : post_col;
	if (R > 0) {
		LAUNCH(ctx, k_hist_flags, grid_for(R, 256), 256, (long long) R, res_col, sepmax_col, post_col, by_radius, thr_select, thr_possible, flag, isel, idef);
		{ int r = scan_int(ctx, isel, selpos, R); if (r) return r; }
		{ int r = scan_int(ctx, idef, defpos, R); if (r) return r; }
		LAUNCH(ctx, k_hist_mark, grid_for(R, 256), 256, (long long) R, res_col, (const unsigned char *) flag,
			(const int *) selpos, (const int *) defpos, sw, weights_cli, first, possible, W);
	}
	if (n > 0) {
		LAUNCH(ctx, k_hist_sources, grid_for(n, 256), 256, (long long) n, mag, (const int *) first, (const unsigned char *) possible, selflag, stats);
		{ int r = scan_int(ctx, selflag, selrank, n + 1); if (r) return r; }   // selflag[n] == 0: selrank[n] is the total
	}
	if (!ctx->h_status) CU(cudaHostAlloc((void **) &ctx->h_status, 64 * sizeof(long long), cudaHostAllocMapped));
	long long *hs = ctx->h_status;
	CU(cudaMemcpyAsync(hs + 48, stats, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	hs[47] = 0;
	if (n > 0) CU(cudaMemcpyAsync(hs + 47, selrank + n, sizeof(int), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	const int64_t nsel = (int64_t) (int) hs[47];
	ENSURE(ctx->d_hout, (size_t) (2 * std::max<int64_t>(nsel, 1) + 2 * MAXB + 8) * sizeof(double));   // + edges / counts of nwb_maghist_count
	double *out = (double *) ctx->d_hout.p;
	if (nsel > 0)
		LAUNCH(ctx, k_hist_gather, grid_for(n, 256), 256, (long long) n, mag, (const int *) first, (const int *) selflag,
			(const int *) selrank, (const double *) (sw ? W : nullptr), out, out + nsel);
	ctx->hist_cat = c; ctx->hist_k = k;
	ctx->stats[0] = nsel;
	if (nselected) *nselected = nsel;
	if (counts3) { counts3[0] = hs[48]; counts3[1] = hs[49]; counts3[2] = hs[50]; }
	if (minmax2) {
		for (int j = 0; j < 2; j++) {
			unsigned long long u = (unsigned long long) hs[51 + j];
			u ^= (u >> 63) ? 0x8000000000000000ull : ~0ull;
			memcpy(&minmax2[j], &u, 8);
		}
		if (hs[49] == 0) minmax2[0] = minmax2[1] = NAN;
	}
	return NWB_OK;
}

int nwb_maghist_select(nwb_ctx *ctx, int c, int k, int by_radius, double thr_select, double thr_possible, int weights_cli,
	int64_t *nselected, int64_t *counts3, double *minmax2)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	if (c < 1 || c >= ctx->res_ncat) return fail(ctx, NWB_ERR_ARG, "catalogue index out of range");
	return maghist_select_impl(ctx, c, k, ctx->nrows, (const long long *) ctx->cols.idx[c], (const double *) ctx->cols.sepmax,
		(const double *) ctx->cols.dist_post, by_radius, thr_select, thr_possible, weights_cli, nselected, counts3, minmax2);
}

int nwb_maghist_select_rows(nwb_ctx *ctx, int c, int k, int64_t nrows, const int64_t *res_dev, const double *sepmax_dev, const double *dist_post_dev,
	int by_radius, double thr_select, double thr_possible, int weights_cli, int64_t *nselected, int64_t *counts3, double *minmax2)
{
	if (!ctx) return NWB_ERR_ARG;
	if (nrows < 0 || (nrows > 0 && (!res_dev || !sepmax_dev || !dist_post_dev))) return fail(ctx, NWB_ERR_ARG, "NULL column");
	return maghist_select_impl(ctx, c, k, nrows, (const long long *) res_dev, sepmax_dev, dist_post_dev, by_radius, thr_select, thr_possible,
		weights_cli, nselected, counts3, minmax2);
}

int nwb_maghist_sample(nwb_ctx *ctx, int64_t nselected, double *mag_host, double *weight_host)
{
	if (!ctx) return NWB_ERR_ARG;
	if (ctx->hist_cat < 0 || nselected != ctx->stats[0]) return fail(ctx, NWB_ERR_STATE, "nwb_maghist_select first");
	if (nselected <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	const double *out = (const double *) ctx->d_hout.p;
	CU(cudaMemcpyAsync(mag_host, out, nselected * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(weight_host, out + nselected, nselected * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

int nwb_maghist_count(nwb_ctx *ctx, int c, int k, int nbins, const double *edges, int64_t *counts)
{
	if (!ctx) return NWB_ERR_ARG;
	if (ctx->hist_cat != c || ctx->hist_k != k) return fail(ctx, NWB_ERR_STATE, "nwb_maghist_select for this column first");
	if (nbins < 1 || nbins > MAXB) return fail(ctx, NWB_ERR_ARG, "1.." + std::to_string(MAXB) + " bins");
	CU(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const int64_t n = ctx->cat[c].n;
	const size_t np_ = ((size_t) n + 63) / 64 * 64;
	const unsigned char *possible = (const unsigned char *) ((int *) ctx->d_hsrc.p + 3 * (np_ + 64));
	ENSURE(ctx->d_hout, (size_t) (2 * std::max<int64_t>(ctx->stats[0], 1) + 2 * MAXB + 8) * sizeof(double));
	// the compact sample may still be wanted: put the edges / counts behind it
	double *d_edges = (double *) ctx->d_hout.p + 2 * std::max<int64_t>(ctx->stats[0], 1);
	unsigned long long *d_counts = (unsigned long long *) (d_edges + MAXB + 1);
	CU(cudaMemcpyAsync(d_edges, edges, (size_t) (nbins + 1) * 8, cudaMemcpyHostToDevice, st));
	CU(cudaMemsetAsync(d_counts, 0, (size_t) nbins * 8, st));
	if (n > 0)
		LAUNCH(ctx, k_hist_count, (int) std::min<int64_t>((n + 255) / 256, 148 * 8), 256, (long long) n, ctx->cat[c].mags + (size_t) k * n,
			possible, nbins, (const double *) d_edges, d_counts);
	CU(cudaMemcpyAsync(counts, d_counts, (size_t) nbins * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return NWB_OK;
}

// ---- element-wise surface ------------------------------------------------------------------------------
static int elementwise_io(nwb_ctx *ctx, int64_t n, int nin, const double *const *in, const int64_t *in_len, double **d_in, double **d_out)
{
	size_t tot = 0;
	for (int k = 0; k < nin; k++) tot += (size_t) in_len[k];
	ENSURE(ctx->d_misc, (tot + (size_t) n + 8) * sizeof(double));
	double *b = (double *) ctx->d_misc.p;
	for (int k = 0; k < nin; k++) {
		d_in[k] = b;
		CU(cudaMemcpyAsync(b, in[k], (size_t) in_len[k] * 8, cudaMemcpyHostToDevice, ctx->stream));
		b += in_len[k];
	}
	*d_out = b;
	return NWB_OK;
}

int nwb_dist(nwb_ctx *ctx, int64_t n, const double *ra1, const double *dec1, const double *ra2,
	const double *dec2, double *out_deg)
{
	if (!ctx) return NWB_ERR_ARG;
	if (n <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	const double *in[4] = {ra1, dec1, ra2, dec2};
	int64_t len[4] = {n, n, n, n};
	double *d_in[4], *d_out;
	{ int r = elementwise_io(ctx, n, 4, in, len, d_in, &d_out); if (r) return r; }
	LAUNCH(ctx, k_dist, grid_for(n, 256), 256, (long long) n, d_in[0], d_in[1], d_in[2], d_in[3], d_out);
	CU(cudaMemcpyAsync(out_deg, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

int nwb_log_bf(nwb_ctx *ctx, int64_t n, int ncat, const double *sep, const double *err, double *out)
{
	if (!ctx) return NWB_ERR_ARG;
	if (ncat < 1 || ncat > MAXC) return fail(ctx, NWB_ERR_ARG, "ncat out of range");
	if (n <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	// norm / log10e only
	ConstTables T;
	memset(&T, 0, sizeof(T));
	const double log_arcsec2rad = std::log(3600 * 180 / M_PI);
	for (int k = 0; k <= MAXC; k++) T.norm[k] = (k - 1) * std::log(2.0) + 2 * (k - 1) * log_arcsec2rad;
	T.log10e = std::log10(M_E);
	const double *in[2] = {sep, err};
	int64_t len[2] = {(int64_t) ncat * ncat * n, (int64_t) ncat * n};
	double *d_in[2], *d_out;
	{ int r = elementwise_io(ctx, n + (int64_t) (sizeof(ConstTables) / 8 + 1), 2, in, len, d_in, &d_out); if (r) return r; }
	ConstTables *d_T = (ConstTables *) (d_out + n);
	CU(cudaMemcpyAsync(d_T, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	LAUNCH(ctx, k_log_bf, grid_for(n, 256), 256, (long long) n, ncat, d_in[0], d_in[1], (const ConstTables *) d_T, d_out);
	CU(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

int nwb_log_bf_elliptical(nwb_ctx *ctx, int64_t n, int ncat, const double *sep_ra, const double *sep_dec,
	const double *err, double *out)
{
	if (!ctx) return NWB_ERR_ARG;
	if (ncat < 1 || ncat > MAXC) return fail(ctx, NWB_ERR_ARG, "ncat out of range");
	if (n <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	ConstTables T;
	memset(&T, 0, sizeof(T));
	const double log_arcsec2rad = std::log(3600 * 180 / M_PI);
	for (int k = 0; k <= MAXC; k++) T.norm[k] = (k - 1) * std::log(2.0) + 2 * (k - 1) * log_arcsec2rad;
	T.log10e = std::log10(M_E);
	const double *in[3] = {sep_ra, sep_dec, err};
	int64_t len[3] = {(int64_t) ncat * ncat * n, (int64_t) ncat * ncat * n, (int64_t) ncat * 3 * n};
	double *d_in[3], *d_out;
	{ int r = elementwise_io(ctx, n + (int64_t) (sizeof(ConstTables) / 8 + 1), 3, in, len, d_in, &d_out); if (r) return r; }
	ConstTables *d_T = (ConstTables *) (d_out + n);
	CU(cudaMemcpyAsync(d_T, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	LAUNCH(ctx, k_log_bf_ell, grid_for(n, 256), 256, (long long) n, ncat, d_in[0], d_in[1], d_in[2], (const ConstTables *) d_T, d_out);
	CU(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

int nwb_row_offsets(nwb_ctx *ctx, int a, int b, double *dra_host, double *ddec_host)
{
	if (!ctx) return NWB_ERR_ARG;
	if (!ctx->matched) return fail(ctx, NWB_ERR_STATE, "nwb_match has not run");
	if (a < 0 || b <= a || b >= ctx->res_ncat) return fail(ctx, NWB_ERR_ARG, "need catalogue indices a < b");
	const int64_t R = ctx->nrows;
	if (R <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->d_misc, (size_t) (2 * R + 8) * sizeof(double));
	double *d = (double *) ctx->d_misc.p;
	LAUNCH(ctx, k_row_offsets, grid_for(R, 256), 256, (long long) R, (const long long *) ctx->cols.idx[a],
		(const long long *) ctx->cols.idx[b], ctx->cat[a].ra, ctx->cat[a].dec, ctx->cat[b].ra, ctx->cat[b].dec, d, d + R);
	CU(cudaMemcpyAsync(dra_host, d, R * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaMemcpyAsync(ddec_host, d + R, R * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

int nwb_score_rows(nwb_ctx *ctx, int64_t nrows, const int64_t *idx, double *sep, double *sepmax, int64_t *ncat_out, double *log_bf, double *dist_post)
{
	if (!ctx) return NWB_ERR_ARG;
	if (nrows < 0 || (nrows > 0 && (!idx || !sepmax || !ncat_out || !log_bf || !dist_post))) return fail(ctx, NWB_ERR_ARG, "NULL argument");
	const int nc = ctx->ncat;
	if (nc < 2 || !ctx->params_set) return fail(ctx, NWB_ERR_STATE, "set the catalogues and nwb_set_params first");
	bool ell = false;
	for (int c = 0; c < nc; c++) {
		if (!ctx->cat[c].set) return fail(ctx, NWB_ERR_STATE, "catalogue " + std::to_string(c) + " not set");
		ell = ell || ctx->cat[c].err_kind == NWB_ERR_ELLIPSE;
	}
	if (ell)
		for (int c = 0; c < nc; c++)
			if (ctx->cat[c].err_kind != NWB_ERR_ELLIPSE) return fail(ctx, NWB_ERR_ARG, "elliptical mode: every catalogue must carry (sigma_x, sigma_y, rho)");
	if (nrows == 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	if (!ctx->tables_set && ctx->tables_dirty) default_tables(ctx);
	{ int r = upload_tables(ctx); if (r) return r; }
	PairStore none[MAXC];
	memset(none, 0, sizeof(none));
	RowParams saved = ctx->rp;
	{ int r = fill_row_params(ctx, none, 0, 0, ell); if (r) return r; }
	const RowParams rp = ctx->rp;
	ctx->rp = saved;   // the parameter block of the last match stays what nwb_finalize expects
	const int npairs = nc * (nc - 1) / 2;
	const size_t R = (size_t) nrows;
	ENSURE(ctx->d_misc, R * (size_t) (nc + npairs + 4) * 8 + 64);
	long long *d_idx = (long long *) ctx->d_misc.p;
	double *d_sep = (double *) (d_idx + R * nc), *d_max = d_sep + R * npairs, *d_lbf = d_max + R, *d_post = d_lbf + R;
	long long *d_ncat = (long long *) (d_post + R);
	cudaStream_t st = ctx->stream;
	CU(cudaMemcpyAsync(d_idx, idx, R * nc * 8, cudaMemcpyHostToDevice, st));
	LAUNCH(ctx, k_score_rows, grid_for(nrows, 128), 128, rp, (long long) nrows, (const long long *) d_idx, d_sep, d_max, d_ncat, d_lbf, d_post);
	if (sep) CU(cudaMemcpyAsync(sep, d_sep, R * npairs * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(sepmax, d_max, R * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(ncat_out, d_ncat, R * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(log_bf, d_lbf, R * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(dist_post, d_post, R * 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return NWB_OK;
}

int nwb_posterior(nwb_ctx *ctx, int64_t n, const double *prior, const double *log_bf, double *out)
{
	if (!ctx) return NWB_ERR_ARG;
	if (n <= 0) return NWB_OK;
	CU(cudaSetDevice(ctx->device));
	const double *in[2] = {prior, log_bf};
	int64_t len[2] = {n, n};
	double *d_in[2], *d_out;
	{ int r = elementwise_io(ctx, n, 2, in, len, d_in, &d_out); if (r) return r; }
	LAUNCH(ctx, k_posterior, grid_for(n, 256), 256, (long long) n, d_in[0], d_in[1], d_out);
	CU(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return NWB_OK;
}

}  // extern "C"
