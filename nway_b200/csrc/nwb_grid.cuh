// Grid over (ra, dec) and everything that decides WHICH (primary, secondary) pairs reach the exact test: the band / cell
// geometry, the registration of a primary in the cells its search box overlaps, the packed cell entries and the fp32
// pre-tests.  Kept apart from the kernels so that tests/emu/grid_emu.cpp can compile exactly these functions for the
// host and check, on millions of random geometries, that no pair within the radius is ever filtered out.
#pragma once
#include "nwb_device.cuh"

namespace nwb {

// ---------------------------------------------------------------------------------------------------------
// grid over (ra, dec): declination bands of height h, each cut into nra[b] cells along ra
// ---------------------------------------------------------------------------------------------------------
struct BandRec {   // 16 bytes, one load
	int base;       // first cell of the band
	int nra;        // cells along ra
	double inv_w;   // cells per degree of ra
};

struct Grid {
	double dec_lo, inv_h;
	double ra_org, ra_span;
	double ra_org_n;        // ra_org brought into [0, 360)
	int nbands;
	int full_circle;
	const BandRec *bands;   // [nbands]
	long long ncells;
	float rr2;              // squared radius (deg^2) of the fp32 flat pre-test, all margins included
	double tau_max;         // primaries with rb_rad * tan|dec| above this are tested on dec only
	const float *kx;        // [nbands] packed pre-test: degrees of true angle per cell width, rounded down; 0 = dec only
	float hdeg;             // band height in degrees
	float rr2p;             // squared radius of the packed pre-test (rr2's margins + the quantisation of the entries)
	const unsigned *bits2;  // sparse primaries, the pure stream (k_filter): a second, REGULAR and COARSE occupancy bitmap, small enough
	                        // for every block to keep a copy in shared memory -- nb2 declination strips of equal height times nr2
	                        // cells of equal ra width: cell = floor(t * band2_scale) * nr2 + floor(x * inv_w2).  A superset of `bits`
	                        // (a primary's search box is registered cell by cell, k_prim_bits2)
	int nr2, nb2;           // nr2: multiple of 32; nb2 * nr2 <= 32 * KF_SMEM_WORDS
	double inv_w2;          // regular cells per degree of ra
	double band2_scale;     // nb2 / nbands
	const unsigned *bits;   // one bit per cell: any primary registered?  nullptr when most cells are occupied anyway.  A sparse
	                        // primary catalogue leaves ~97 % of the cells empty; the bitmap (ncells / 8 bytes: L1 / L2 resident)
	                        // answers those without touching the 32-byte cell records
};

// 16 bytes, one load: one primary as seen from one cell, for the fp32 flat PRE-test
//     (dy)^2 + (clat * dx)^2 <= rr2,     dx, dy relative to the grid origin (small numbers: fp32 is accurate)
// which is a superset of the exact disc (derivation in DESIGN.md: hav(theta) = hav(ddec) + cos d1 cos d2 hav(dra),
// cos d2 >= cos d1 (1 - ddec^2/2 - |ddec| tan|d1|)); the exact fp64 separation decides.  Primaries too close to a
// pole for the flat metric get clat = 0, i.e. they are pre-tested on declination only.
struct Entry {
	float x, y, clat;
	int p;
};

// 8 bytes: the same primary in CELL units, for the entries that live inside the cell record.  x, y = position
// relative to the cell's origin (x in cell widths, y in band heights), fixed point over [-1.5, 2.5): 15 bits for x
// (bit 15 = "always pass": the encoding did not apply), 16 bits for y.  The pre-test then is
//     ((xr - px) kx)^2 + ((yr - py) h)^2 <= rr2p,     (xr, yr) = the secondary inside its cell, in [0, 1)^2,
// kx = cos(dec) x cell width for the band (rounded down, so the test only gets more permissive; 0 for bands near a
// pole).  No RA wrap logic: px is measured from the unwrapped cell index.
struct PEntry {
	unsigned xy;
	int p;
};
constexpr float PE_OFF = 1.5f;
constexpr float PE_XSTEP = 4.0f / 32768.0f, PE_YSTEP = 4.0f / 65536.0f;

__device__ __forceinline__ unsigned pe_encode(double px, double py, bool always)
{
	double fx = rint((px + 1.5) * 8192.0), fy = rint((py + 1.5) * 16384.0);
	if (!(fx >= 0.0 && fx <= 32767.0 && fy >= 0.0 && fy <= 65535.0)) { always = true; fx = 0.0; fy = 0.0; }
	return (unsigned) (int) fx | (always ? 0x8000u : 0u) | ((unsigned) (int) fy << 16);
}

__device__ __forceinline__ double wrap360(double x)
{
	if (x >= 0.0 && x < 360.0) return x;
	double y = x - 360.0 * floor(x * (1.0 / 360.0));
	return (y >= 360.0 || y < 0.0) ? 0.0 : y;
}

__device__ __forceinline__ int band_of(const Grid &G, double dec)
{
	double t = (dec - G.dec_lo) * G.inv_h;
	if (!(t >= 0.0)) return -1;
	if (t >= (double) G.nbands) return G.nbands;
	return __double2int_rd(t);
}

__device__ __forceinline__ int racell_of(const BandRec &B, double x /* wrap360(ra - ra_org) */)
{
	int i = __double2int_rd(x * B.inv_w);
	return i >= B.nra ? B.nra - 1 : (i < 0 ? 0 : i);
}

__device__ __forceinline__ BandRec load_band(const Grid &G, int b)
{
	const int4 v = __ldg(reinterpret_cast<const int4 *>(G.bands + b));
	BandRec B;
	B.base = v.x; B.nra = v.y;
	B.inv_w = __hiloint2double(v.w, v.z);
	return B;
}

// ---------------------------------------------------------------------------------------------------------
// K0: primaries
// ---------------------------------------------------------------------------------------------------------
struct PrimRec {   // 32 bytes = one L2 sector: what the exact formula needs of a primary
	double lon, slat, clat;
	long long ij;   // NWB_COMPAT_FLAT_HASH: the reference's flat-sky cell of the primary (flat_hash_pack), else 0
};

struct PrimArrays {
	PrimRec *rec;                // for the exact formula
	double *clat;                // cos(dec) for the fp32 entries of crowded cells: float-valued, rounded down, 0 near a pole
	double *ra_n, *dec, *dra;    // search box (degrees); dra >= 180 means "all ra"
};

// 32 bytes = one L2 sector per cell, fetched with one 256-bit load: how many primaries are registered here and the
// first three of them inline, packed (PEntry); entries beyond the third become work items.  A secondary in a cell
// with <= 3 primaries (90 % of the occupied cells at C3's densities) needs no second lookup.
//     q[0] = cnt | (overflow segment start - 3) << 32      q[1] = entry 0      q[2] = entry 1      q[3] = entry 2
// q[0] is written by k_cell_headers, q[1..3] by the fill pass (whichever three primaries asked first); entry k >= 3 of
// the cell is entries[(q[0] >> 32) + k].  Slots beyond cnt are never read.
struct CellRec {
	unsigned long long q[4];
};

// Half-width in ra (degrees) of the box that contains every point within rb degrees of a source at declination d
// (the small circle's extreme longitudes: sin(dra) = sin(rb) / cos(d)), slightly inflated; 360 = "all of ra" near a pole.
__device__ __forceinline__ double search_box_dra(double d, double rb)
{
	if (fabs(d) + rb >= 89.999) return 360.0;
	double s = sin(rb / 180 * NWB_PI) / cos((fabs(d)) / 180 * NWB_PI);
	return s >= 1.0 ? 360.0 : asin(s) * 180 / NWB_PI * (1 + 1e-9) + 1e-12;
}

// Register one primary in every grid cell its (slightly inflated) search box overlaps.
// MODE = REG_COUNT: cellcnt[cell] += 1.
// MODE = REG_FILL : take a slot of the cell by counting cellcnt back down (no second memset).  Slots 0..2 live INSIDE the
// cell record (packed, see PEntry) and are written there directly; later slots go to the cell's overflow segment of
// `entries`, whose start k_cell_headers put into the record.
// MODE = REG_COUNT_INLINE (grid geometry known beforehand): count AND place in one pass -- the slot is the value the
// counting atomicAdd returns; slots 0..2 are written into the cell record right away, the few later ones (5 % at the
// benchmark's densities) are noted in a work list (primary, cell, slot) and stored by k_fill_overflow once
// k_cell_headers has handed out the overflow segments.  No separate fill pass over all primaries.
// All passes enumerate the same (band, cell) pairs from the same doubles, whichever thread layout (bslot, bstride).
enum { REG_COUNT = 0, REG_FILL = 1, REG_COUNT_INLINE = 2 };

struct OverflowItem { int p, cell, slot, pad; };   // 16 bytes

template <int MODE>
__device__ __forceinline__ void prim_register(const Grid &G, const int i, const double d, const double rn, const double dra,
	const double clat_i, const double rb_ins, const double dra_eps, const int bslot, const int bstride,
	int *__restrict__ cellcnt, CellRec *cells, Entry *__restrict__ entries,
	OverflowItem *__restrict__ worklist = nullptr, int *__restrict__ worklist_n = nullptr, long long worklist_cap = 0)
{
	int b0 = band_of(G, d - rb_ins), b1 = band_of(G, d + rb_ins);
	b0 = max(b0, 0);
	b1 = min(b1, G.nbands - 1);
	Entry en;
	bool have_en = false;
	const double di = dra + dra_eps;
	const double xp = wrap360(rn - G.ra_org);          // the primary along ra, from the grid origin
	const double yp = (d - G.dec_lo) * G.inv_h;        // ... and in band heights
	for (int b = b0 + bslot; b <= b1; b += bstride) {
		BandRec B = load_band(G, b);
		int n = B.nra;
		int i0, cnt;
		const int ip = racell_of(B, xp);               // the primary's own cell in this band
		const double xc = xp * B.inv_w - (double) ip;  // its position inside that cell, in cell widths
		if (G.full_circle) {
			const double cellw = G.ra_span / n;
			if (2 * di + 2 * cellw >= 360.0) { i0 = 0; cnt = n; }
			else {
				i0 = racell_of(B, wrap360(rn - di - G.ra_org));
				int i1 = racell_of(B, wrap360(rn + di - G.ra_org));
				cnt = (i1 - i0 + n) % n + 1;
			}
		} else {
			// the grid's ra window was built from min(rn - dra) .. max(rn + dra) with a margin: no wrap inside
			double x0 = wrap360(rn - G.ra_org) - di, x1 = wrap360(rn - G.ra_org) + di;
			i0 = racell_of(B, fmax(x0, 0.0));
			int i1 = racell_of(B, fmin(x1, G.ra_span));
			cnt = i1 - i0 + 1;
		}
		int cb = B.base;
		if (MODE == REG_COUNT)
			for (int k = 0; k < cnt; k++) atomicAdd(&cellcnt[cb + (i0 + k) % n], 1);
		// four cells at a time: the slot requests (atomics with return) are independent, issued together so that their
		// latencies overlap; then the stores
		for (int k0 = 0; MODE != REG_COUNT && k0 < cnt; k0 += 4) {
			int sl[4], cell[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				sl[u] = -1;
				cell[u] = 0;
				if (k0 + u < cnt) {
					cell[u] = cb + (i0 + k0 + u) % n;
					sl[u] = MODE == REG_FILL ? atomicSub(&cellcnt[cell[u]], 1) - 1 : atomicAdd(&cellcnt[cell[u]], 1);
				}
			}
#pragma unroll
			for (int u = 0; u < 4; u++) {
				if (sl[u] < 0) continue;
				if (sl[u] < 3) {
					// packed: offset of this cell from the primary's own, unwrapped
					int m = i0 + k0 + u - ip;
					bool always = false;
					if (G.full_circle) {
						if (m > n / 2) m -= n;
						else if (m < -(n / 2)) m += n;
						always = n < 4 || cnt >= n;   // too few cells to tell which way round is the near one
					}
					const unsigned xy = pe_encode(xc - (double) m, yp - (double) b, always);
					cells[cell[u]].q[1 + sl[u]] = (unsigned long long) xy | ((unsigned long long) (unsigned) i << 32);
				} else if (MODE == REG_COUNT_INLINE) {
					const int pos = atomicAdd(worklist_n, 1);
					if ((long long) pos < worklist_cap) {
						int4 v;
						v.x = i; v.y = cell[u]; v.z = sl[u]; v.w = 0;
						*reinterpret_cast<int4 *>(worklist + pos) = v;
					}
				} else {
					if (!have_en) {   // the fp32 entry of the crowded cells' work items (rare: computed on demand)
						double x = rn - G.ra_org_n;
						if (x < 0.0) x += 360.0;
						en.x = (float) x;
						en.y = (float) (d - G.dec_lo);
						en.clat = (float) clat_i;   // k_prim_prep applied the pole rule and the rounding
						en.p = i;
						have_en = true;
					}
					const int est = (int) (cells[cell[u]].q[0] >> 32);   // written by k_cell_headers (an earlier launch)
					*reinterpret_cast<int4 *>(entries + est + sl[u]) = *reinterpret_cast<const int4 *>(&en);
				}
			}
		}
	}
}
// The reference's flat-sky hash (fastskymatch.py:125-132) puts a source into the buckets (i, j), (i+1, j), (i, j+1),
// (i+1, j+1) with i = int(ra / err), j = int(dec / err) (python's int(): truncation toward zero; err = the search radius
// in degrees, match_radius / 60. / 60), and forms tuples per bucket: a tuple exists iff the cells of its present members
// span at most one step in i and in j.  Off the equator that loses pairs a complete search finds.
// NWB_COMPAT_FLAT_HASH applies exactly this predicate after the exact search -- to every (primary, secondary) pair in
// k_pairs, to every tuple in k_count_rows / k_rows -- whenever the reference would take that branch
// (fastskymatch.py:94-98, decided on the host from the catalogues' bounds).
__device__ __forceinline__ long long flat_hash_cell(double coord_deg, double err_deg)
{
	return (long long) (coord_deg / err_deg);   // IEEE division, conversion truncates toward zero
}

__device__ __forceinline__ bool flat_hash_same_bucket(long long ia, long long ja, long long ib, long long jb)
{
	const long long di = ia - ib, dj = ja - jb;
	return di >= -1 && di <= 1 && dj >= -1 && dj <= 1;
}

// The bucket size and its reciprocal.  int(coord / err) needs the correctly rounded quotient (a cell boundary is
// exactly where the rounding decides); q0 = RN(coord * rerr), the FMA residual and one FMA correction give it without
// the division subroutine (Markstein, as div_const) -- coordinates are at most 360 and err is in (360 / 2^31, 1), so
// nothing over- or underflows on the way.  Checked against coord / err in tests/test_device_arithmetic_cpu.py.
struct FlatHash {
	double err;    // 0 = NWB_COMPAT_FLAT_HASH not in force
	double rerr;   // RN(1 / err)
};

__device__ __forceinline__ int flat_hash_cell_fast(double coord_deg, const FlatHash &F)
{
	const double q0 = coord_deg * F.rerr;
	const double r = fma(-q0, F.err, coord_deg);
	return (int) fma(r, F.rerr, q0);
}

// (i, j) of a source in one 64-bit word (the host refuses radii so small that 360 / err leaves the int range)
__device__ __forceinline__ long long flat_hash_pack(double ra_deg, double dec_deg, const FlatHash &F)
{
	const int i = flat_hash_cell_fast(ra_deg, F), j = flat_hash_cell_fast(dec_deg, F);
	return (long long) (((unsigned long long) (unsigned) i << 32) | (unsigned long long) (unsigned) j);
}
__device__ __forceinline__ int flat_hash_i(long long w) { return (int) (w >> 32); }
__device__ __forceinline__ int flat_hash_j(long long w) { return (int) (unsigned) (unsigned long long) w; }

// Where a secondary falls in the grid (first stage of k_pairs; the band record B is fetched in between): declination in
// band heights from the lower edge (valid if 0 <= t < nbands), ra in degrees from the grid origin in [0, 360) (valid if
// the grid spans the circle or x <= ra_span), the cell along ra (clamped like racell_of) and the position in cell widths.
__device__ __forceinline__ double k1_band_coord(const Grid &G, double dec) { return (dec - G.dec_lo) * G.inv_h; }

__device__ __forceinline__ double k1_ra_coord(const Grid &G, double ra)
{
	double x = wrap360(ra) - G.ra_org_n;
	if (x < 0.0) x += 360.0;
	return x;
}

__device__ __forceinline__ double k1_ra_cell(const BandRec &B, double x, int &ic)
{
	const double xcells = x * B.inv_w;
	ic = __double2int_rd(xcells);
	ic = ic >= B.nra ? B.nra - 1 : (ic < 0 ? 0 : ic);
	return xcells;
}

// the fp32 flat pre-test (see struct Entry); (x, y) = the secondary relative to the grid origin
__device__ __forceinline__ bool k1_pretest(const Grid &G, float x, float y, float ex, float ey, float eclat)
{
	float dx = x - ex;
	if (dx > 180.f) dx -= 360.f;
	else if (dx < -180.f) dx += 360.f;
	float u = dx * eclat;
	float dy = y - ey;
	return u * u + dy * dy <= G.rr2;
}

// the packed pre-test (see struct PEntry); (xr, yr) = the secondary inside its cell
__device__ __forceinline__ bool k1_pretest_packed(const Grid &G, float xr, float yr, float kx, unsigned xy)
{
	float px = (float) (xy & 0x7fffu) * PE_XSTEP - PE_OFF;
	float py = (float) (xy >> 16) * PE_YSTEP - PE_OFF;
	float fx = (xr - px) * kx;
	float fy = (yr - py) * G.hdeg;
	return (xy & 0x8000u) != 0u || fx * fx + fy * fy <= G.rr2p;
}

}  // namespace nwb
