// Device-side arithmetic of the nway match path.  fp64 throughout, compiled with --fmad=false so that the
// operation order below IS the rounding order (the reference is unfused numpy).
#pragma once
#ifdef NWB_HOST_EMU
#include "nwb_host_emu.h"   // tests/emu: the few CUDA names these headers use, for a plain host compiler
#else
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>
#include "nwb_sincos_tab.h"

#define NWB_FULL 0xffffffffu
#define NWB_PI 3.141592653589793

namespace nwb {

constexpr int MAXC = 8;                       // catalogues
constexpr int MAXP = MAXC * (MAXC - 1) / 2;   // catalogue pairs
constexpr int MAXM = 8;                       // magnitude columns (all catalogues together)
constexpr int MAXB = 64;                      // bins per magnitude prior

// One 32-byte sector per lane in ONE request (sm_100 has 256-bit global loads: LDG.E.256).  A 128-bit + a 64-bit load
// of the same record would walk the L1TEX tag / data stages twice, and k_pairs is bound by exactly those
// (l1tex__data_pipe_lsu_wavefronts at 90 % of peak, profiles/r01_step3_*).  Read-only data only (.nc).
struct Sector32 { unsigned long long q[4]; };
__device__ __forceinline__ Sector32 ldg_sector(const void *p /* 32-byte aligned */)
{
	Sector32 s;
#ifdef NWB_HOST_EMU
	s = *reinterpret_cast<const Sector32 *>(p);
#else
	asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s.q[0]), "=l"(s.q[1]), "=l"(s.q[2]), "=l"(s.q[3]) : "l"(p));
#endif
	return s;
}

// x / b for a compile-time constant b, correctly rounded like the IEEE division the reference performs, but
// without the division subroutine: q = RN(x * RN(1/b)), one FMA residual, one FMA correction (Markstein).
// Checked against x / b on 4e8 random arguments for b = 180 and b = pi (0 mismatches).
__device__ __forceinline__ double div_const(double x, double b, double rb)
{
	double q = x * rb;
	double r = fma(-q, b, x);
	return fma(r, rb, q);
}
// a / b from a reciprocal y = RN(1 / b) that is already there (the row kernels memoise it per error value): the
// correctly rounded quotient (Markstein: q0 = RN(a y), r = a - b q0 exactly by FMA, RN(q0 + r y) = RN(a / b)) as long as
// nothing under- or overflows on the way -- anything else (a huge or tiny, b denormal-ish) takes the division.
__device__ __forceinline__ double quotient_by_reciprocal(double a, double b, double y)
{
	if (fabs(a) < 1e280 && fabs(a) > 1e-280 && b > 1e-280 && b < 1e280) {
		double q0 = a * y;
		double rr = fma(-q0, b, a);
		return fma(rr, y, q0);
	}
	return a / b;
}
#define NWB_INV180 (1.0 / 180.0)
#define NWB_INVPI (1.0 / NWB_PI)

// ---------------------------------------------------------------------------------------------------------
// sin / cos with the bits of the reference's.  The reference evaluates fastskymatch.py:36-41 with numpy, whose
// float64 sin / cos are glibc's (numpy 2.x ships no SIMD double-precision sin / cos).  The Vincenty numerator
// (fastskymatch.py:44) cancels: a 1-ulp difference in sin(lat) or cos(dlon) is an ABSOLUTE ~1e-16 rad in the
// separation, which the Bayes factor multiplies by separation / sigma^2 -- up to a few 1e-10 in the posteriors.  So
// the path needs glibc's bits, not merely an accurate sine.  What follows restates the algorithm of glibc 2.39's
// double-precision sin / cos (sysdeps/ieee754/dbl-64/s_sin.c as built for x86-64 with FMA -- the variant every
// AVX2-class host selects): |x| < 0.126: odd polynomial; below 0.855469: table of sin / cos(k/128) in double-double
// plus degree-5/6 corrections; below 2.426265: the same around pi/2 - |x|; below 105414350: Cody-Waite reduction by
// pi/2 in four pieces.  Every rounding -- which products are fused into an FMA and which are not -- follows that build
// (the TU is compiled --fmad=false, so fma() below is the only fusion).  tests/test_device_arithmetic_cpu.py builds
// this header for the host and compares with libm on tens of millions of arguments: identical bits.
// ---------------------------------------------------------------------------------------------------------
#ifdef NWB_HOST_EMU
static const double nwb_sincostab[4 * NWB_SINCOS_ROWS] = {NWB_SINCOS_TABLE};
#else
__device__ __align__(32) const double nwb_sincostab[4 * NWB_SINCOS_ROWS] = {NWB_SINCOS_TABLE};
#endif

struct SinCosRow { double sn, ssn, cs, ccs; };

__device__ __forceinline__ SinCosRow sincos_row(int k)
{
	const Sector32 r = ldg_sector(nwb_sincostab + 4 * k);   // one 32-byte row, one request
	SinCosRow t;
	t.sn = __longlong_as_double((long long) r.q[0]); t.ssn = __longlong_as_double((long long) r.q[1]);
	t.cs = __longlong_as_double((long long) r.q[2]); t.ccs = __longlong_as_double((long long) r.q[3]);
	return t;
}

namespace gl {
constexpr double big = 0x1.8p45, toint = 0x1.8p52;
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
constexpr double cs2 = 0.5, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
	s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
constexpr double hpinv = 0x1.45f306dc9c883p-1, mp1 = 0x1.921fb58p+0, mp2 = -0x1.dde973cp-27, pp3 = -0x1.cb3b398p-55,
	pp4 = -0x1.d747f23e32ed7p-83;

// x + dx -> sin, |x| < 0.126
__device__ __forceinline__ double taylor_sin(double x, double dx)
{
	const double xx = x * x;
	const double poly = fma(fma(fma(fma(s5, xx, s4), xx, s3), xx, s2), xx, s1);
	const double t = fma(xx, fma(poly, x, -(0.5 * dx)), dx);
	return x + t;
}

// where |x| falls in the table: row k = rint(|x| * 128), remainder xr = |x| - k/128 (exact)
__device__ __forceinline__ int table_split(double ax, double &xr)
{
	const double u = big + ax;
	xr = ax - (u - big);
	return __double2loint(u);
}

__device__ __forceinline__ double do_sin(double x, double dx)
{
	if (fabs(x) < 0.126) return taylor_sin(x, dx);
	if (x <= 0) dx = -dx;
	double xr;
	const SinCosRow T = sincos_row(table_split(fabs(x), xr));
	const double xx = xr * xr;
	const double s = xr + fma(xr * xx, fma(xx, sn5, sn3), dx);
	const double c = fma(xr, dx, xx * fma(xx, fma(xx, cs6, cs4), cs2));
	const double cor = fma(s, T.cs, fma(-c, T.sn, fma(s, T.ccs, T.ssn)));
	return copysign(T.sn + cor, x);
}

__device__ __forceinline__ double do_cos(double x, double dx)
{
	if (x < 0) dx = -dx;
	double xr;
	const SinCosRow T = sincos_row(table_split(fabs(x), xr));
	xr = xr + dx;
	const double xx = xr * xr;
	const double s = fma(xr * xx, fma(xx, sn5, sn3), xr);
	const double c = xx * fma(xx, fma(xx, cs6, cs4), cs2);
	const double cor = fma(-s, T.sn, fma(-c, T.cs, fma(-s, T.ssn, T.ccs)));
	return T.cs + cor;
}

// x = n pi/2 + (a + da), |a| <= pi/4 (+ rounding); 136 bits of pi/2
__device__ __forceinline__ int reduce(double x, double &a, double &da)
{
	const double t = fma(x, hpinv, toint);
	const double xn = t - toint;
	const int n = __double2loint(t) & 3;
	const double y = fma(-xn, mp2, fma(-xn, mp1, x));
	const double t2 = fma(-xn, pp3, y);
	const double d1 = fma(-pp3, xn, y - t2);
	const double b = fma(-xn, pp4, t2);
	const double d2 = fma(-xn, pp4, t2 - b);
	a = b;
	da = d1 + d2;
	return n;
}

__device__ __forceinline__ double do_sincos(double a, double da, int n)
{
	const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
	return (n & 2) ? -r : r;
}
}  // namespace gl

// the complete functions, every branch.  Out of line: the hot paths below take the one or two branches their arguments
// can reach inline and come here only for the rest.
#ifndef NWB_SINCOS_ANY_INLINE
#define NWB_SINCOS_ANY_INLINE 0
#endif
#if NWB_SINCOS_ANY_INLINE
#define NWB_ANY_ATTR __forceinline__
#else
#define NWB_ANY_ATTR __noinline__
#endif
__device__ NWB_ANY_ATTR double sin_ref_any(double x)
{
	const int k = __double2hiint(x) & 0x7fffffff;
	if (k < 0x3e500000) return x;
	if (k < 0x3feb6000) return gl::do_sin(x, 0.0);
	if (k < 0x400368fd) return copysign(gl::do_cos(gl::hp0 - fabs(x), gl::hp1), x);
	if (k < 0x419921fb) {
		double a, da;
		const int n = gl::reduce(x, a, da);
		return gl::do_sincos(a, da, n);
	}
	return sin(x);   // |x| > 1e8 rad: not an angle this path produces
}

__device__ NWB_ANY_ATTR double cos_ref_any(double x)
{
	const int k = __double2hiint(x) & 0x7fffffff;
	if (k < 0x3e400000) return 1.0;
	if (k < 0x3feb6000) return gl::do_cos(x, 0.0);
	if (k < 0x400368fd) {
		const double y = gl::hp0 - fabs(x);
		const double a = y + gl::hp1;
		const double da = (y - a) + gl::hp1;
		return gl::do_sin(a, da);
	}
	if (k < 0x419921fb) {
		double a, da;
		const int n = gl::reduce(x, a, da);
		return gl::do_sincos(a, da, n + 1);
	}
	return cos(x);
}

// sin and cos of a small angle (a longitude difference of a close pair): below 1/256 rad the sine is the odd
// polynomial and the cosine sits on row 0 of the table (sin 0 = 0, cos 0 = 1: the corrections collapse to 1 - c) --
// the same roundings as do_sin / do_cos, no table access
__device__ __forceinline__ void sincos_ref_small(double x, double *s, double *c)
{
#if defined(NWB_AB_LIBRARY_SINCOS) && !defined(NWB_HOST_EMU)   // A/B measurement only: what the reference's bits cost (round-1 arithmetic)
	sincos(x, s, c);
	return;
#endif
	const double ax = fabs(x);
	if (ax <= 0x1p-8 && ax >= 0x1p-26) {
		*s = gl::taylor_sin(x, 0.0);
		const double xx = ax * ax;
		*c = 1.0 - xx * fma(xx, fma(xx, gl::cs6, gl::cs4), gl::cs2);
	} else {
		*s = sin_ref_any(x);
		*c = cos_ref_any(x);
	}
}

// sin and cos of one angle in [-pi/2, pi/2] (a latitude): the table row, the remainder and the two polynomials are
// shared between the sine and the cosine wherever glibc's two functions take the same branch
__device__ __forceinline__ void sincos_ref(double x, double *s, double *c)
{
#if defined(NWB_AB_LIBRARY_SINCOS) && !defined(NWB_HOST_EMU)
	sincos(x, s, c);
	return;
#endif
	const double ax = fabs(x);
	const int hx = __double2hiint(ax);   // glibc branches on the high word
	if (hx < 0x3feb6000 && hx >= 0x3e500000) {
		double xr;
		const int k = gl::table_split(ax, xr);
		const double xx = xr * xr;
		const double x3p = (xr * xx);
		const double p = fma(xx, gl::sn5, gl::sn3);
		const double cc = xx * fma(xx, fma(xx, gl::cs6, gl::cs4), gl::cs2);
		if (k == 0) {   // |x| <= 1/256: row 0 is (0, 0, 1, 0)
			*s = gl::taylor_sin(x, 0.0);
			*c = 1.0 - cc;
			return;
		}
		const SinCosRow T = sincos_row(k);
		const double sc = fma(x3p, p, xr);                 // do_cos(x, 0): s
		*c = T.cs + fma(-sc, T.sn, fma(-cc, T.cs, fma(-sc, T.ssn, T.ccs)));
		if (ax < 0.126) {
			*s = gl::taylor_sin(x, 0.0);
		} else {
			const double ss = xr + x3p * p;                  // do_sin(x, 0): s (dx = +-0 adds nothing), c = cc
			*s = copysign(T.sn + fma(ss, T.cs, fma(-cc, T.sn, fma(ss, T.ccs, T.ssn))), x);
		}
		return;
	}
	if (hx >= 0x3feb6000 && hx < 0x400368fd) {   // around pi/2
		const double y = gl::hp0 - ax;
		*s = copysign(gl::do_cos(y, gl::hp1), x);
		const double a = y + gl::hp1;
		const double da = (y - a) + gl::hp1;
		*c = gl::do_sin(a, da);
		return;
	}
	*s = sin_ref_any(x);
	*c = cos_ref_any(x);
}

__device__ __forceinline__ double sin_ref(double x) { return sin_ref_any(x); }
__device__ __forceinline__ double cos_ref(double x) { return cos_ref_any(x); }

// reference: nwaylib/fastskymatch.py:31-34 -- divide by 180 first, then multiply by pi
__device__ __forceinline__ double deg2rad_ref(double x) { return div_const(x, 180.0, NWB_INV180) * NWB_PI; }

// Great-circle separation in ARCSEC between a (lower catalogue index) and b, from the precomputed
// lon = ra/180*pi, sin/cos(dec/180*pi) (sincos_ref).  fastskymatch.py:36-47 then *60*60 (__init__.py:163).
// The operation order is the reference's and sin / cos of the longitude difference carry the reference's bits
// (sin_ref / cos_ref above): num2 and den are O(1) sums that cancel down to the separation, so those bits ARE the
// result.  What follows the cancellation only has to be accurate in the relative sense: for separations below
// 2^-7 rad hypot / atan2 are a square root and atan's Taylor polynomial (truncation < 2^-70); anything else (near
// the poles, huge radii) takes the library routines.
__device__ __forceinline__ double sep_arcsec_ref(double lon1, double slat1, double clat1,
	double lon2, double slat2, double clat2)
{
	const double dlon = lon2 - lon1;
	double sdlon, cdlon;
	sincos_ref_small(dlon, &sdlon, &cdlon);
	double num1 = clat2 * sdlon;
	double num2 = clat1 * slat2 - slat1 * clat2 * cdlon;
	double den = slat1 * slat2 + clat1 * clat2 * cdlon;
	double h2 = fma(num1, num1, num2 * num2);
	double ang;
	if (den > 0.5 && h2 < 0x1p-16 && h2 > 1e-280) {
		double t = sqrt(h2) / den;
		double t2 = t * t;
		double pa = fma(t2, fma(t2, fma(t2, 1.0 / 9, -1.0 / 7), 1.0 / 5), -1.0 / 3);
		ang = fma(t * t2, pa, t);
	} else {
		ang = atan2(hypot(num1, num2), den);
	}
	double deg = div_const(ang * 180, NWB_PI, NWB_INVPI);
	return deg * 60 * 60;
}

// index of the catalogue pair (a < b) in the order _create_match_table emits the Separation columns
// (__init__.py:143-168): (0,1),(0,2),...,(1,2),...
__host__ __device__ __forceinline__ int pair_index(int a, int b, int ncat)
{
	return a * (2 * ncat - a - 1) / 2 + (b - a - 1);
}

// one magnitude prior as a step function (magnitudeweights.py:74-87)
struct MagTable {
	int cat;            // catalogue the column belongs to
	int nbins;
	const double *mag;  // device column, n[cat] values
	double edges[MAXB + 1];
	double weight[MAXB];  // log10(ratio), NaN already replaced by 0 (__init__.py:388)
	double bias[MAXB];    // 10**weight     (__init__.py:392)
};

// host-computed scalar tables, all evaluated with the reference's own python/numpy expressions
struct ConstTables {
	double norm[MAXC + 1];             // (n-1)*log(2) + 2*(n-1)*log_arcsec2rad, by n   bayesdistance.py:76
	double log10e;                     // numpy.log10(numpy.e)
	double prior[1 << (MAXC - 1)];     // by presence mask of the secondaries (bit c-1)  __init__.py:254
	double log10prior[1 << (MAXC - 1)];
	double sub_log10prior[1 << (MAXC - 1)];  // log10(nu[A0]/prod(nu_plus[A]))           nway.py:395
	MagTable mag[MAXM];
};

// log10 Bayes factor of the present catalogues (bayesdistance.py:64-86).
// present: bit c set if catalogue c takes part (bit 0 = primary for a full row; sub-associations of the CLI
// correction pass a mask without bit 0).  sig[c]: sigma in arcsec.  sep[pair_index(a,b)]: arcsec.
// sq32: the separations are float32 values and are squared in float32, as numpy does for the command-line program's
// 'E' columns (`p[i][j]**2`, bayesdistance.py:83, on what nway.py:302-305 reads back); the products stay fp64.
template <int NC>
__device__ __forceinline__ double log_bf_ref(const ConstTables *__restrict__ T, int ncat_rt, unsigned present,
	const double *sig, const double *sep, bool sq32 = false)
{
	const int ncat = NC > 0 ? NC : ncat_rt;
	int n = __popc(present);
	if (n <= 1)
		return 0.0;   // norm = 0, slog = log(w) - log(w) = 0, q = 0  ->  exactly 0
	double w[MAXC];
	double wsum = 0.0, slog = 0.0;
	bool first = true;
#pragma unroll
	for (int c = 0; c < (NC > 0 ? NC : MAXC); c++) {
		if (c < ncat && (present >> c & 1u)) {
			double s = sig[c];
			w[c] = 1.0 / (s * s);
			double lw = log(w[c]);
			if (first) { wsum = w[c]; slog = lw; first = false; }
			else { wsum = wsum + w[c]; slog = slog + lw; }
		}
	}
	slog = slog - log(wsum);
	double q = 0.0;
	bool qfirst = true;
#pragma unroll
	for (int a = 0; a < (NC > 0 ? NC : MAXC); a++) {
#pragma unroll
		for (int b = a + 1; b < (NC > 0 ? NC : MAXC); b++) {
			if (b < ncat && (present >> a & 1u) && (present >> b & 1u)) {
				double p = sep[pair_index(a, b, ncat)];
				double p2 = sq32 ? (double) __fmul_rn((float) p, (float) p) : p * p;
				double term = w[a] * w[b] * p2;
				if (qfirst) { q = term; qfirst = false; } else q = q + term;   // 0 + term == term
			}
		}
	}
	double exponent = -q / 2 / wsum;
	return (T->norm[n] + slog + exponent) * T->log10e;
}

// ---- elliptical positional errors (CLI only in the reference): bayesdistance.py:164-240, fastskymatch.py:50-74 ----
// Offsets (arcsec) of the target t in the tangent frame centred on the origin o: what astropy's
// SkyOffsetFrame(origin=o) gives for t, with the sign of fastskymatch.py:65-66.
__device__ __forceinline__ void offsets_ref(double ra_o, double dec_o, double ra_t, double dec_t, double &dra, double &ddec)
{
	const double D2R = 0.017453292519943295, R2D = 57.29577951308232;   // numpy.radians / numpy.degrees constants
	double so, co, st, ct, sd, cd;
	sincos_ref(dec_o * D2R, &so, &co);
	sincos_ref(dec_t * D2R, &st, &ct);
	sincos_ref(ra_t * D2R - ra_o * D2R, &sd, &cd);
	double lon = atan2(ct * sd, co * ct * cd + so * st);
	double z = co * st - so * ct * cd;
	double lat = asin(fmin(fmax(z, -1.0), 1.0));
	dra = -(lon * R2D) * 60 * 60;
	ddec = -(lat * R2D) * 60 * 60;
}

// v^T Sigma^-1 v for the unit vector v (bayesdistance.py:150-161,183-187)
__device__ __forceinline__ double dir_precision(double vx, double vy, double sx, double sy, double rho)
{
	double f = 1.0 / (sx * sx * (sy * sy) * (1 - rho * rho));
	double m11 = f * (sy * sy), m12 = f * -rho * sx * sy, m22 = f * (sx * sx);
	double l1 = vx * m11 + vy * m12;
	double l2 = vx * m12 + vy * m22;
	return l1 * vx + l2 * vy;
}

// separation of one pair rescaled by the ratio of circular to directional error (bayesdistance.py:224-238)
__device__ __forceinline__ double ell_rescaled_sep(double vx, double vy, double sxa, double sya, double rhoa,
	double sxb, double syb, double rhob, double siga, double sigb)
{
	double d = sqrt(vx * vx + vy * vy);
	double ux = d == 0 ? 0.7071067811865476 : vx / (d + 1e-300);
	double uy = d == 0 ? 0.7071067811865476 : vy / (d + 1e-300);
	double wa = dir_precision(ux, uy, sxa, sya, rhoa);
	double wb = dir_precision(ux, uy, sxb, syb, rhob);
	double ratio = (siga * siga + sigb * sigb) / (1 / wa + 1 / wb);
	return d * (1.0 / sqrt(ratio));
}

// 10^x for the row kernels.  Same scheme as the CUDA library routine (k = rint(x log2 10), r = x - k log10 2 in two
// pieces, degree-13 polynomial of 10^r, scale by 2^k) but with the coefficients in constant memory, so a call is
// ~30 instructions instead of ~70 (the library materialises every 64-bit coefficient with two moves).  Checked
// against 80-bit powl on 4e7 arguments over [-320, 300]: max error 1.25 ulp (library: 1 ulp).
__constant__ double c_exp10_poly[13] = {
	2.3025850929940459, 2.6509490552391992, 2.034678592293476, 1.1712551489122669, 0.5393829291955814,
	0.2069958486968681, 0.068089365074437067, 0.019597694626478524, 0.0050139288337754401,
	0.0011544997789984348, 0.00024166672554424694, 4.6371516642572196e-05, 8.2134125354393867e-06};

#ifndef NWB_OWN_EXP10
#define NWB_OWN_EXP10 1
#endif
__device__ __forceinline__ double nwb_exp10(double x)
{
#if !NWB_OWN_EXP10
	return exp10(x);
#endif
	if (!(x >= -323.4)) return x != x ? x : 0.0;
	if (x > 308.3) return INFINITY;
	const double kd = rint(x * 3.321928094887362348);
	double r = fma(-kd, 0.3010299950838089, x);          // log10(2), 26 significant bits: k * hi is exact
	r = fma(-kd, 5.8017229629986518e-10, r);
	double p = c_exp10_poly[12];
#pragma unroll
	for (int n = 11; n >= 0; n--) p = fma(p, r, c_exp10_poly[n]);
	p = fma(p, r, 1.0);
	const int k = (int) kd;
	if (k > -1000 && k < 1000)
		return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
	const int k1 = k / 2, k2 = k - k1;   // subnormal / near-overflow results: scale in two steps
	return p * __hiloint2double((1023 + k1) << 20, 0) * __hiloint2double((1023 + k2) << 20, 0);
}

// Scalar pieces of the fused group normalisation of k_rows2 (__init__.py:423-457), with ONE exponential per row:
//   t_k = 10^(v_k - m_rest) (k >= 1, m_rest = max of them),  s_rest = sum t_k,  p_i = t_k / s_rest.
// p_any = 1 - 10^(v0 - bfsum), bfsum = log10(s_all) + m_all, s_all = sum 10^(v_k - m_all)   (__init__.py:428-439)
//       = 1 - 10^(v0 - m_all) / s_all = [sum over k >= 1 of 10^(v_k - m_all)] / s_all
// -- the same number without the logarithm, the second exponential and the cancellation of "1 -"; it differs from the
// reference's rounding of that expression by the reference's own ~1e-14 (parity metric).  One of the two
// exponents is zero: e0 = 10^(-|v0 - m_rest|).  rinv = 1 / s_rest (the largest t_k is exactly 1, so max p_i = rinv).
__device__ __forceinline__ void group_p_any(int rows, double v0, double m_rest, double s_rest, double &p_any, double &rinv)
{
	p_any = 0.0;
	rinv = 0.0;
	if (rows > 1) {
		const double m_all = fmax(v0, m_rest);
		const double e0 = nwb_exp10(fmin(v0, m_rest) - m_all);
		const double rest = v0 >= m_rest ? s_rest * e0 : s_rest;      // sum over k >= 1, scaled by 10^(-m_all)
		const double s_all = v0 >= m_rest ? 1.0 + rest : rest + e0;
		p_any = rest / s_all;
		rinv = 1.0 / s_rest;
	}
}

// dist_post = 1/(1 + (1 - prior) 10^(-v)) from the exponential the normalisation evaluated anyway: 10^(-v_k) =
// 10^(-m_rest) / t_k, i.e. t_k / (t_k + oscale) with oscale = (1 - prior) 10^(-m_rest); the direct form where that
// would leave the double range (|m_rest| > 250, t_k = 0)
__device__ __forceinline__ double shared_post(bool direct, double t, double oscale, double omp, double lbf, double l10p1)
{
	if (direct || !(t > 1e-290)) return 1. / (1 + omp * nwb_exp10(-lbf - l10p1));
	return t / (t + oscale);
}

// bayesdistance.py:26-32
__device__ __forceinline__ double posterior_ref(double prior, double log10prior, double log_bf)
{
	return 1. / (1 + (1 - prior) * exp10(-log_bf - log10prior));
}

// zero-order-hold lookup; returns the weight and sets bias.  m is already -99 for undefined
// (__init__.py:383-388).
__device__ __forceinline__ double mag_weight(const MagTable &M, double m, double &bias)
{
	if (!(m >= M.edges[0] && m <= M.edges[M.nbins])) {   // also true for NaN
		bias = 1.0;
		return 0.0;
	}
	int k = 0;
	for (int j = 1; j < M.nbins; j++)
		if (M.edges[j] <= m) k = j;
	bias = M.bias[k];
	return M.weight[k];
}

__device__ __forceinline__ double warp_max(double v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(NWB_FULL, v, o));
	return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NWB_FULL, v, o);
	return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NWB_FULL, v, o);
	return v;
}

}  // namespace nwb
