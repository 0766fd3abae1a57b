// Host stand-ins for the few CUDA names nway_b200/csrc/nwb_device.cuh and nwb_grid.cuh use, so that a plain host
// compiler can build exactly those device functions (tests/emu/grid_emu.cpp).  One "thread", sequential semantics:
// atomics are plain read-modify-writes, __ldg is a load, warp shuffles return their argument.
// Compile with -ffp-contract=off: the device code is built with --fmad=false.
#pragma once
#define _GNU_SOURCE 1
#include <math.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __constant__ static const
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
static inline double2 make_double2(double x, double y) { double2 v = {x, y}; return v; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { int old = *p; *p = old + v; return old; }
static inline int atomicSub(int *p, int v) { int old = *p; *p = old - v; return old; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long old = *p; *p = old + v; return old; }
static inline int __double2int_rd(double x) { return (int) floor(x); }
static inline float __double2float_rd(double x) { float f = (float) x; return (double) f > x ? nextafterf(f, -INFINITY) : f; }
static inline double __hiloint2double(int hi, int lo)
{
	unsigned long long u = ((unsigned long long) (unsigned) hi << 32) | (unsigned) lo;
	double d;
	memcpy(&d, &u, 8);
	return d;
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int __double2hiint(double d) { unsigned long long u; memcpy(&u, &d, 8); return (int) (u >> 32); }
static inline int __double2loint(double d) { unsigned long long u; memcpy(&u, &d, 8); return (int) (u & 0xffffffffu); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline long long __shfl_xor_sync(unsigned, long long v, int) { return v; }
