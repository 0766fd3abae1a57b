/*
 * nwayb200.h -- C ABI of the B200-native nway match-probability path.
 *
 * The reference (JohannesBuchner/nway 4.7.1) is pure Python and has no FFI of its own; the boundary it
 * offers is the Python function nwaylib.nway_match() (nwaylib/__init__.py:31-120).  This library is what a
 * ctypes binding of that function's numeric stages binds to: every entry point below names the reference
 * stage it replaces.  Plain pointers and sizes only; the caller owns every buffer it passes in; nothing is
 * retained past a call except device copies held inside the opaque context.  No exceptions, no exit():
 * every function returns 0 on success or a negative nwb_status; nwb_last_error() gives the text.
 *
 * Threading: one context per (process, device); calls on one context must be serialised by the caller;
 * different contexts are independent.  All work of a context runs on one private CUDA stream.
 *
 * Units follow the reference: ra/dec in degrees, positional errors and radii in arcsec, areas in deg^2.
 */
#ifndef NWAYB200_H
#define NWAYB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nwb_ctx nwb_ctx;

enum nwb_status {
	NWB_OK = 0,
	NWB_ERR_CUDA = -1,        /* a CUDA runtime call failed (text has the CUDA error string) */
	NWB_ERR_ARG = -2,         /* invalid argument / call order */
	NWB_ERR_EMPTY = -3,       /* no rows: maps to nwaylib.EmptyResultException (__init__.py:92-93) */
	NWB_ERR_NOMEM = -4,
	NWB_ERR_STATE = -5        /* results requested before nwb_match / nwb_finalize ran */
};

enum { NWB_MAX_CATS = 8, NWB_MAX_MAGS = 8, NWB_MAX_BINS = 64 };

/* error kinds for nwb_set_catalogue */
enum { NWB_ERR_CIRCULAR = 1,   /* err = n sigma values (arcsec)                 -> bayesdistance.log_bf           */
       NWB_ERR_ELLIPSE = 3 };  /* err = 3 columns of n: sigma_x, sigma_y, rho   -> bayesdistance.log_bf_elliptical */

/* unrelated-association correction */
enum { NWB_UNRELATED_API = 0,  /* nwaylib/__init__.py:262-301: inert in the reference -> no change */
       NWB_UNRELATED_CLI = 1 };/* nway.py:366-421: the live algorithm                                              */

/* stages for nwb_timing (device milliseconds of the last run, CUDA events on the context's stream) */
enum { NWB_T_GRID = 0,      /* primary prep + cell lists                       fastskymatch.py:119-133            */
       NWB_T_PAIRS = 1,     /* stream secondaries, exact separations, append   fastskymatch.py:134-160, 26-47     */
       NWB_T_LISTS = 2,     /* scan + scatter + per-primary sort               fastskymatch.py:178-181,217        */
       NWB_T_ROWS = 3,      /* enumerate + score (+ fused finalize)            __init__.py:123-259                */
       NWB_T_FINAL = 4,     /* bias lookup, p_single, group log-sum-exp        __init__.py:376-461                */
       NWB_T_TOTAL = 5,
       NWB_T_KPAIRS = 6,    /* k_pairs launches alone (summed over the secondary catalogues)                      */
       NWB_T_KROWS = 7,     /* the k_rows launch alone                                                            */
       NWB_T_COUNT = 8 };

/* ---- lifetime ------------------------------------------------------------------------------------------ */

int nwb_create(int device, nwb_ctx **out);
void nwb_destroy(nwb_ctx *ctx);
const char *nwb_last_error(nwb_ctx *ctx);   /* ctx may be NULL: error of the last failed nwb_create */
int nwb_version(void);
/* run all work of this context on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the private one; stream == NULL restores the private stream. */
int nwb_set_stream(nwb_ctx *ctx, void *cuda_stream);

/* ---- inputs -------------------------------------------------------------------------------------------- */

/* Catalogue c of ncat (c = 0 is the primary).  Replaces the match_tables[c] dict of nway_match
 * (__init__.py:38-48).  mags: m columns of n values, column after column; NaN or -99 = undefined
 * (__init__.py:319,384).  on_device != 0: the pointers are device pointers on the context's device and are
 * used in place (they must stay valid until the next nwb_set_catalogue for c or nwb_destroy); otherwise
 * they are host pointers and are copied (pinned host memory makes the copy asynchronous DMA).
 * Limits: at most NWB_MAX_CATS catalogues and NWB_MAX_MAGS magnitude columns in total; source indices travel as 32-bit
 * words inside the match (slots, lists), so a catalogue may hold just under 2^31 sources (2.1e9; 48 GB of coordinates and
 * errors) -- nwb_match returns NWB_ERR_ARG ("catalogue too large") beyond that; row counts and offsets are 64-bit. */
int nwb_set_catalogue(nwb_ctx *ctx, int c, int ncat, int64_t n, const double *ra, const double *dec,
	const double *err, int err_kind, const double *mags, int m, double area, int on_device);

/* Scalars of nway_match (__init__.py:31-36).  completeness: ncat values, [0] == 1 (__init__.py:224-229,
 * already expanded by the caller if scalar). */
int nwb_set_params(nwb_ctx *ctx, double match_radius_arcsec, const double *completeness,
	double prob_ratio_secondary, int unrelated_mode);

/* --prefilter-pair of nway.py (fastskymatch.crossproduct's pairwise_errs, fastskymatch.py:184-208), as intended: an
 * association that contains a source of catalogue cat_a[k] AND one of cat_b[k] is only formed if the two are closer
 * than radius_arcsec[k]; associations lacking either are unaffected.  The reference's code as written drops every
 * such association regardless of distance (its mask assignment writes to a temporary) -- that
 * behaviour is radius 0 here.  npairs = 0 clears the list.  Call after the catalogues are set. */
int nwb_set_prefilter(nwb_ctx *ctx, int npairs, const int *cat_a, const int *cat_b, const double *radius_arcsec);

/* Compatibility switches; default 0 = fp64 separations, complete enumeration on the whole sphere.
 * NWB_COMPAT_SEP_F32: nway.py stores every separation (and, for elliptical errors, every tangent-plane offset) in a
 * float32 FITS column and scores what it reads back (fastskymatch.py:328-331 -> nway.py:269,302-305):
 * round them to float32 after the (fp64) radius filter and before the Bayes factor.  The Separation columns then
 * hold float32-representable values.
 * NWB_COMPAT_FLAT_HASH: reproduce the ROW SET of the reference where its enumeration is incomplete.  When every
 * catalogue has |dec| < 45, ra in (10 r, 360 - 10 r) and r < 1 deg, crossproduct() hashes on square cells of r degrees
 * in (ra, dec) WITHOUT cos(dec) (fastskymatch.py:94-101,123-133): away from the equator pairs whose ra difference
 * spans more than two cells are never formed although they are within the radius.  With this flag a
 * pair / tuple is kept only if the reference's hash would have put all of its members into one bucket (cells
 * int(ra / err), int(dec / err), err = match_radius / 60. / 60, spanning <= 1 step each); when the reference would
 * take its HEALPix branch instead (fastskymatch.py:134-160, complete) nothing changes.  The catalogue bounds this
 * depends on are reduced on the device once per nwb_set_catalogue.  nwb_flat_hash_applied tells which it was. */
enum { NWB_COMPAT_SEP_F32 = 1, NWB_COMPAT_FLAT_HASH = 2 };
int nwb_set_compat(nwb_ctx *ctx, int flags);
/* after nwb_match: *applied = 1 if the last match applied the flat-sky bucket predicate (NWB_COMPAT_FLAT_HASH set and
 * the reference's flat-sky condition fastskymatch.py:94-98 true for these catalogues), else 0 */
int nwb_flat_hash_applied(nwb_ctx *ctx, int *applied);

/* Optional: the scalar tables the kernels consume, computed by the caller with the reference's own numpy
 * expressions (so they are bit-identical to what nwaylib computes on that host).  If not called, the library
 * derives the same tables itself in C double arithmetic.  norm: ncat+1 values indexed by the number of present
 * catalogues (bayesdistance.py:76); prior / log10prior: 2^(ncat-1) values indexed by the presence mask of the
 * secondaries, bit c-1 (__init__.py:254, bayesdistance.py:32,39); sub_log10prior: log10(nu[A0]/prod nu_plus[A])
 * for the CLI correction (nway.py:395). */
int nwb_set_tables(nwb_ctx *ctx, const double *norm, double log10e, const double *prior,
	const double *log10prior, const double *sub_log10prior);

/* Magnitude prior k of catalogue c as a step function (magnitudeweights.py:74-87): nbins+1 ascending edges,
 * and per bin the weight log10(ratio) (__init__.py:386) and the bias column value 10**weight (__init__.py:392).
 * The host computes both with the reference's own expressions so the table is bit-identical. */
int nwb_set_maghist(nwb_ctx *ctx, int c, int k, int nbins, const double *edges, const double *weight,
	const double *bias);

/* Shard: only primaries [first, first+count) are matched by this context.  Row indices in the
 * output stay global.  Default: all. */
int nwb_set_primary_range(nwb_ctx *ctx, int64_t first, int64_t count);

/* ---- multi-GPU, strong scaling: the streaming of the secondaries split between the ranks --------------------------
 * The shards of nwb_set_primary_range are independent but every one of them streams every secondary, so their time
 * does not shrink with the number of GPUs once the stream dominates (BASELINE.json configs[3], configs[4]: 1e5 .. 1e6
 * primaries against 1e7 .. 3e8 secondaries).  Shard mode divides the STREAM instead: every rank holds all catalogues in
 * HBM (the row stages gather errors, magnitudes and coordinates by global index) and builds the grid over ALL primaries,
 * but streams only its slice [n r / W, n (r + 1) / W) of every secondary catalogue; a match is written straight into the
 * pair store of the rank that OWNS the primary (contiguous blocks of ceil(n0 / W) primaries) -- one system-scope
 * atomicAdd and one 16-byte store into that rank's exchange buffer over NVLink peer memory (cudaIpc) from inside the
 * streaming kernel, no staging, no collective on the data path.  Each rank then produces the rows of its own primaries;
 * the table is reassembled as in the other mode (nway_b200.parallel.allgather_table).
 *   nwb_shard_setup    (after the catalogues and nwb_set_params) allocates this rank's exchange buffer and returns its
 *                      64-byte cudaIpcMemHandle; spill_capacity = matches beyond a primary's slots the buffer can hold
 *   nwb_shard_connect  takes all ranks' handles (world x 64 bytes, rank order; the caller all-gathers them) and opens them
 *   nwb_shard_match    one match in two phases with ONE barrier between the ranks in between (the caller's: any stream-ordered
 *                      collective on the context's stream): 1 = grid + streaming -- matches arrive from all ranks -- , 2 = rows
 *                      of the own primaries.  The exchange buffer holds two sets of counters and slots: match e uses set
 *                      e % 2 and zeroes the other one for its successor, which is what makes a single barrier enough.
 *                      Phase 2 returns 1 when a grid buffer turned out too small: every rank sees the same (the primaries
 *                      are replicated) and repeats phases 1 and 2.  Phase 0 resets both sets (after an aborted match; a
 *                      barrier must follow).  world = 1 degenerates to an ordinary match.
 *                      Phase 3 = phases 1 and 2 in one call with the barrier as flags in peer memory: every rank stores a
 *                      growing tag into every rank's exchange buffer behind its streaming kernel and waits, on the device,
 *                      for the others' (gives up after 10 s: NWB_ERR_STATE) -- no collective library on the path at all.
 *   nwb_shard_close    leaves shard mode. */
int nwb_shard_setup(nwb_ctx *ctx, int rank, int world, int64_t spill_capacity, void *ipc_handle_out, int64_t *exchange_bytes);
int nwb_shard_connect(nwb_ctx *ctx, const void *ipc_handles);
int nwb_shard_match(nwb_ctx *ctx, int phase, int fuse_final, int64_t *nrows);
int nwb_shard_close(nwb_ctx *ctx);

/* ---- multi-GPU: the reassembly of the sharded table over peer memory ------------------------------------------------
 * Either mode above leaves every rank with the rows of its own primaries; the reference's caller holds ONE table
 * (nwaylib/__init__.py:272-275 builds it in one process).  nwb_gather_* reassemble it on every rank without a
 * collective library on the data path: each rank's GPU stores its shard -- column by column, from where the row kernels
 * wrote it -- straight into its final position in the gathered table of every rank over NVLink peer memory (cudaIpc).
 * No padding to the largest shard, no staging buffer, no unpacking copy: the bytes cross the link once and land in place.
 *   nwb_gather_setup    allocates this rank's gathered-table buffer -- TWO sets of ncols columns of capacity_rows 8-byte
 *                       values, alternating between pushes, so that a rank may still read table e while its peers already
 *                       push e + 1 -- and returns its 64-byte cudaIpcMemHandle
 *   nwb_gather_connect  takes all ranks' handles (world x 64 bytes, rank order; the caller all-gathers them)
 *   nwb_gather_push     counts[world] = rows of every rank's shard, on the host (counts_on_device = 0) or in device memory
 *                       (1: e.g. the result of an all-gather of the ranks' nwb_nrows_device_ptr words enqueued on the same
 *                       stream -- no host round trip between the match and the push).
 *                       Enqueues the push of the own shard on the context's stream (engine 0: one kernel of 16-byte
 *                       stores from the SMs; 1: the copy engines, one stream per destination, host counts only; 2: as 0,
 *                       but self-contained -- counts may be NULL: the ranks' row counts travel as flag words through the
 *                       buffers' headers before the push, and a second exchange of flags behind it IS the barrier, so
 *                       the table is complete when the stream gets there; nwb_gather_counts then returns the counts) and
 *                       returns the device address of this rank's gathered table for this push: column k at
 *                       table + k * stride_bytes, sum(counts) rows, shards in rank order.  The table is complete once
 *                       every rank's push has finished: the caller places ONE stream-ordered barrier between the ranks
 *                       behind it (any collective on the context's stream).  A table beyond capacity_rows is an error
 *                       (host counts) or is not written at all (device counts: check sum(counts) afterwards).
 *   nwb_gather_close    closes the peer mappings and frees the buffer. */
int nwb_gather_setup(nwb_ctx *ctx, int rank, int world, int64_t capacity_rows, int ncols, void *ipc_handle_out);
int nwb_gather_connect(nwb_ctx *ctx, const void *ipc_handles);
int nwb_gather_push(nwb_ctx *ctx, const int64_t *counts, int counts_on_device, int engine, void **table, int64_t *stride_bytes);
int nwb_gather_counts(nwb_ctx *ctx, int64_t *counts /* [world] */);   /* engine 2: waits for the stream; error if a rank never arrived */
int nwb_gather_close(nwb_ctx *ctx);

/* ---- the path ------------------------------------------------------------------------------------------ */

/* H1+H2+H3: bin, enumerate, separations, radius filter, log Bayes factor, prior, dist_post
 * (fastskymatch.crossproduct + __init__._create_match_table + _compute_single_log_bf + posterior).
 * fuse_final != 0 also runs nwb_finalize's work inside the row kernel (allowed when every magnitude prior is
 * already set, i.e. no 'auto' histogram has to be built from dist_post on the host in between).
 * nrows receives R.  Rows stay in device memory. */
int nwb_match(nwb_ctx *ctx, int fuse_final, int64_t *nrows);

/* The same match without the host round trip at its end, for callers that run one match after another on resident
 * catalogues (a shard loop, the calibration loop of nway-calibrate-cutoff.py): nwb_match_async enqueues the whole
 * pipeline on the context's stream and returns; the row count stays on the device (nwb_nrows_device_ptr).
 * nwb_match_wait synchronises, checks the status words the kernels left (grid geometry of the previous match still
 * valid, buffers big enough) and, if they say no, redoes the match step by step -- the result is always that of
 * nwb_match.  When the pipeline cannot be enqueued blindly (first match of a context, three or more catalogues,
 * elliptical errors) nwb_match_async simply runs nwb_match.  Results may be read only after nwb_match_wait. */
int nwb_match_async(nwb_ctx *ctx, int fuse_final);
int nwb_match_wait(nwb_ctx *ctx, int64_t *nrows);

/* H3b+H4: magnitude bias lookup, p_single, per-primary log-sum-exp -> prob_has_match, prob_this_match,
 * match_flag (__init__._apply_magnitude_biasing lookup half + _compute_final_probabilities). */
int nwb_finalize(nwb_ctx *ctx);

/* Automatic magnitude histograms (nwaylib/__init__.py:324-366, nway.py:455-503) between nwb_match(fuse_final=0) and
 * nwb_finalize: the selection runs on the device, only the compact sample of secure counterparts comes back.
 *
 * nwb_maghist_select: over the rows of the last match, for magnitude column k of catalogue c (c >= 1):
 *   by_radius != 0: selected = Separation_max < thr_select, possible = Separation_max < thr_possible, weights 1;
 *   by_radius == 0: selected = dist_post > thr_select, possible = dist_post > thr_possible, weights = dist_post
 *   (both only where the catalogue has a counterpart).  The selected sources are made unique by first occurrence
 *   (numpy.unique(..., return_index=True)); weights_cli chooses the reference's weight indexing of the command-line
 *   program (nway.py:471) instead of the API's (__init__.py:337) -- SURVEY.md Q7.
 *   Out: nselected = unique selected sources; counts3 = {sources with a "possible" row, field sources (finite
 *   magnitude, not possible), sources with a finite magnitude}; minmax2 = min / max magnitude of the field sources.
 * nwb_maghist_sample: the magnitudes (NaN if undefined) and weights of the nselected sources, ascending source index.
 * nwb_maghist_count: numpy.histogram(magnitudes of the field sources, bins=edges) -- integer counts per bin, last bin
 *   closed on the right; the caller turns them into a density exactly as numpy does. */
int nwb_maghist_select(nwb_ctx *ctx, int c, int k, int by_radius, double thr_select, double thr_possible, int weights_cli,
	int64_t *nselected, int64_t *counts3, double *minmax2);
/* The same selection over caller-supplied device columns of nrows rows (index column of catalogue c, Separation_max,
 * dist_post): the rows of ALL shards of a multi-GPU match gathered in global row order -- the first occurrence of a source
 * and the reference's weight indexing are properties of the whole table, not of a shard.  Every rank runs it
 * on the same gathered columns and obtains the same histogram (nway_b200.parallel.nway_match_sharded). */
int nwb_maghist_select_rows(nwb_ctx *ctx, int c, int k, int64_t nrows, const int64_t *res_dev, const double *sepmax_dev,
	const double *dist_post_dev, int by_radius, double thr_select, double thr_possible, int weights_cli,
	int64_t *nselected, int64_t *counts3, double *minmax2);
int nwb_maghist_sample(nwb_ctx *ctx, int64_t nselected, double *mag_host, double *weight_host);
int nwb_maghist_count(nwb_ctx *ctx, int c, int k, int nbins, const double *edges, int64_t *counts);

/* a13 _truncate_table (__init__.py:464-471): keep rows with !(p_i < min_prob); compacts the device table in
 * place and returns the new row count. */
int nwb_truncate(nwb_ctx *ctx, double min_prob, int64_t *nrows);

/* ---- outputs ------------------------------------------------------------------------------------------- */

/* Column selectors for nwb_fetch / nwb_column_ptr.  Index columns: NWB_COL_IDX + c.  Separation of catalogues
 * a < b: NWB_COL_SEP + pair(a,b), pairs numbered (0,1),(0,2),...,(1,2),... as _create_match_table emits them.
 * Bias columns: NWB_COL_BIAS + j, j numbering the (catalogue, mag) pairs in catalogue order. */
enum { NWB_COL_IDX = 0, NWB_COL_SEP = 100, NWB_COL_BIAS = 200,
       NWB_COL_SEPMAX = 300, NWB_COL_NCAT = 301, NWB_COL_LOGBF_UNCORR = 302, NWB_COL_LOGBF = 303,
       NWB_COL_DIST_POST = 304, NWB_COL_P_SINGLE = 305, NWB_COL_MATCH_FLAG = 306, NWB_COL_P_ANY = 307,
       NWB_COL_P_I = 308 };

/* copy one column of the last result (R values of 8 bytes: int64 for IDX/NCAT/MATCH_FLAG, double otherwise)
 * to host memory; asynchronous on the context's stream when dst is pinned -- call nwb_sync() before reading. */
int nwb_fetch(nwb_ctx *ctx, int column, void *dst_host);
/* same, into a DEVICE buffer of the caller on the context's device (e.g. a torch tensor that is then handed to
 * NCCL); asynchronous on the context's stream */
int nwb_fetch_device(nwb_ctx *ctx, int column, void *dst_device);
/* device address of a column (valid until the next nwb_match on this context) */
int nwb_column_ptr(nwb_ctx *ctx, int column, void **dev_ptr);
/* The whole table at once: every column of the last match lives in ONE device allocation, column k (in the output order
 * of the reference: indices, separations, Separation_max, ncat, dist_bayesfactor_uncorrected, dist_bayesfactor,
 * dist_post, biases, p_single, match_flag, prob_has_match, prob_this_match) at base + k * stride_bytes, nrows 8-byte
 * values each.  A multi-GPU caller sends its shard of every column straight from here (nway_b200.parallel) -- no
 * per-column copy. */
int nwb_table_layout(nwb_ctx *ctx, void **base, int64_t *stride_bytes, int *ncols, int64_t *nrows);
/* (valid until nwb_truncate, which compacts the table but not the per-primary offsets this word belongs to)
 * device address of the int64 row count of the last match (lets a multi-GPU caller all-gather the per-rank row
 * counts with NCCL straight from device memory) */
int nwb_nrows_device_ptr(nwb_ctx *ctx, void **dev_ptr);
int nwb_sync(nwb_ctx *ctx);

/* Measurement only: run the memory-system skeleton of the streaming kernel k_pairs for secondary catalogue c on the
 * grid the last nwb_match left -- the same coalesced stream, cell-record gather, primary-record gather, slot atomicAdd
 * and slot store, none of the arithmetic -- `reps` times and return the mean duration.  blocks_per_sm:
 * resident blocks per SM to run it at (0 = as many as the match kernel of the last nwb_match had: the skeleton needs fewer
 * registers, and more resident warps make this access pattern slower, not faster).  bench.py reports it next to the kernel.
 * Invalidates the match result. */
int nwb_bench_skeleton(nwb_ctx *ctx, int c, int reps, int blocks_per_sm, float *ms);

/* per-stage device time of the last nwb_match (+ nwb_finalize), and how many kernels it launched */
int nwb_timing(nwb_ctx *ctx, int stage, float *ms);
int nwb_launch_count(nwb_ctx *ctx, int64_t *launches);
/* counters of the last run: [0] pairs tested exactly, [1] pairs kept, [2] grid cells, [3] cell entries */
int nwb_stats(nwb_ctx *ctx, int64_t *out4);

/* ---- element-wise surface (fastskymatch.dist, bayesdistance.log_bf / posterior) --------------------------
 * host pointers, n elements; used by the thin Python mirrors of those functions. */
int nwb_dist(nwb_ctx *ctx, int64_t n, const double *ra1, const double *dec1, const double *ra2,
	const double *dec2, double *out_deg);                        /* fastskymatch.py:26-47 */
int nwb_log_bf(nwb_ctx *ctx, int64_t n, int ncat, const double *sep /* ncat*ncat blocks of n, only i<j read */,
	const double *err /* ncat blocks of n */, double *out);     /* Score a caller-supplied candidate list with the catalogues, parameters and tables now set on the context: idx = nrows x
 * ncat row indices (row-major, host; -1 = the catalogue takes no part), as fastskymatch.crossproduct returns them
 * (fastskymatch.py:92-218).  Per row, UNFILTERED (the reference filters Separation_max < match_radius afterwards,
 * __init__.py:180): sep = the separations of every catalogue pair in arcsec, pair after pair in the order of the
 * Separation columns, NaN where a member is absent (may be NULL); sepmax (__init__.py:166); ncat (:177); log_bf = the log10
 * Bayes factor of the present catalogues (__init__.py:220-259, bayesdistance.py:64-86); dist_post = posterior(prior, log_bf)
 * (bayesdistance.py:26-32).  Lets a caller tell an arithmetic mismatch from an enumeration mismatch. */
int nwb_score_rows(nwb_ctx *ctx, int64_t nrows, const int64_t *idx, double *sep, double *sepmax, int64_t *ncat,
	double *log_bf, double *dist_post);

/* bayesdistance.py:64-86 */
int nwb_posterior(nwb_ctx *ctx, int64_t n, const double *prior, const double *log_bf, double *out); /* :26-32 */
int nwb_log_bf_elliptical(nwb_ctx *ctx, int64_t n, int ncat, const double *sep_ra, const double *sep_dec /* like sep */,
	const double *err /* ncat blocks of (sigma_x | sigma_y | rho), each n */, double *out);   /* bayesdistance.py:207-240 */

/* Tangent-plane offsets (arcsec) between the members a < b of every row of the last result, R values each, NaN where
 * either is absent: the Separation_<b>_<a>_ra / _dec columns match_multiple adds for non-circular errors
 * (fastskymatch.py:299-331, dist3d :50-74). */
int nwb_row_offsets(nwb_ctx *ctx, int a, int b, double *dra_host, double *ddec_host);

#ifdef __cplusplus
}
#endif
#endif
