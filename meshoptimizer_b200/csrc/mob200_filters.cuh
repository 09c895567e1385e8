// mob200_filters.cuh -- decode filters as device functions (fused epilogue + standalone kernels).
//
// Arithmetic contract: the operation ORDER of the reference's
// x86 SSE2 kernels (reference src/vertexfilter.cpp:256-542), every step a separately rounded IEEE
// binary32 operation -- hence the explicit __f*_rn intrinsics (never contracted to FMA) and
// IEEE-compliant __fdiv_rn/__fsqrt_rn, and no fast-math anywhere.  float->int is round-half-even
// with the x86 "integer indefinite" result for NaN / out-of-range inputs, like cvtps2dq.
//   bit-exact vs reference : Exp, Oct (stride 8), Quat, Color (stride 8)
//   <= 1 LSB vs reference  : Oct (stride 4), Color (stride 4): the reference uses rsqrtps / rcpps
//                            hardware approximations there (:284, :467); we use the correctly
//                            rounded quotient, which is what every KAT in the reference tests expects.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mob200
{

__device__ __forceinline__ int cvt_rne(float x)
{
	int r = __float2int_rn(x);            // saturates; NaN -> 0
	return (x < 2147483648.f) ? r : (int)0x80000000; // NaN and x >= 2^31 -> 0x80000000 like cvtps2dq
}

__device__ __forceinline__ float xor_sign(float t, float x)
{
	return __int_as_float(__float_as_int(t) ^ (__float_as_int(x) & (int)0x80000000));
}

// shared tail of both octahedral variants: x,y already float, z = zf - (|x|+|y|)
__device__ __forceinline__ void oct_core(float& x, float& y, float z, float scale_num, int& xr, int& yr, int& zr)
{
	float t = (z < 0.f) ? z : 0.f; // minps(z, 0)
	x = __fadd_rn(x, xor_sign(t, x));
	y = __fadd_rn(y, xor_sign(t, y));
	float ll = __fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z)));
	float s = __fdiv_rn(scale_num, __fsqrt_rn(ll));
	xr = cvt_rne(__fmul_rn(x, s));
	yr = cvt_rne(__fmul_rn(y, s));
	zr = cvt_rne(__fmul_rn(z, s));
}

// reference src/vertexfilter.cpp:256-299; element = 4 bytes [x y z w]
__device__ __forceinline__ uint32_t filter_oct8(uint32_t v)
{
	float x = (float)(int)(signed char)(v & 0xff);
	float y = (float)(int)(signed char)((v >> 8) & 0xff);
	float zf = (float)(int)(signed char)((v >> 16) & 0xff);
	float z = __fsub_rn(zf, __fadd_rn(fabsf(x), fabsf(y)));
	int xr, yr, zr;
	oct_core(x, y, z, 127.f, xr, yr, zr);
	return (v & 0xff000000u) | ((uint32_t)xr & 0xffu) | (((uint32_t)yr & 0xffu) << 8) | (((uint32_t)zr & 0xffu) << 16);
}

// reference src/vertexfilter.cpp:301-357; element = 8 bytes, lo = x | y<<16, hi = z | w<<16
__device__ __forceinline__ uint2 filter_oct16(uint2 v)
{
	float x = (float)(int)(short)(v.x & 0xffff);
	float y = (float)((int)v.x >> 16);
	float zf = (float)(int)(v.y & 0x7fff);
	float z = __fsub_rn(zf, __fadd_rn(fabsf(x), fabsf(y)));
	int xr, yr, zr;
	oct_core(x, y, z, 32767.f, xr, yr, zr);
	uint2 r;
	r.x = ((uint32_t)xr & 0xffffu) | ((uint32_t)yr << 16);
	r.y = ((uint32_t)zr & 0xffffu) | (v.y & 0xffff0000u);
	return r;
}

// reference src/vertexfilter.cpp:359-423
__device__ __forceinline__ uint2 filter_quat(uint2 v)
{
	const float scale = __int_as_float(0x46b50389); // fl(32767.f / fl(sqrtf(2.f))), computed in binary32 as :361 does
	float x = (float)(int)(short)(v.x & 0xffff);
	float y = (float)((int)v.x >> 16);
	float z = (float)(int)(short)(v.y & 0xffff);
	int c = (int)v.y >> 16;
	float s = (float)(c | 3);

	float ws = __fmul_rn(s, __fadd_rn(s, s));
	float ww = __fsub_rn(ws, __fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
	float w = __fsqrt_rn((ww > 0.f) ? ww : 0.f); // maxps(ww, 0)
	float ss = __fdiv_rn(scale, s);

	uint32_t xr = (uint32_t)cvt_rne(__fmul_rn(x, ss)) & 0xffffu;
	uint32_t yr = (uint32_t)cvt_rne(__fmul_rn(y, ss)) & 0xffffu;
	uint32_t zr = (uint32_t)cvt_rne(__fmul_rn(z, ss)) & 0xffffu;
	uint32_t wr = (uint32_t)cvt_rne(__fmul_rn(w, ss)) & 0xffffu;

	// 16-bit lanes [w x y z] rotated left by 16*(c&3) bits
	unsigned long long packed = (unsigned long long)(wr | (xr << 16)) | ((unsigned long long)(yr | (zr << 16)) << 32);
	unsigned rot = ((unsigned)c & 3u) * 16u;
	unsigned long long res = rot ? (packed << rot) | (packed >> (64 - rot)) : packed;
	return make_uint2((uint32_t)res, (uint32_t)(res >> 32));
}

// reference src/vertexfilter.cpp:425-443; denormals are kept (no FTZ), inf*0 gives the x86 default NaN
__device__ __forceinline__ uint32_t filter_exp(uint32_t v)
{
	int e = (int)v >> 24;
	int m = (int)(v << 8) >> 8;
	float p = __uint_as_float((uint32_t)(e + 127) << 23);
	float r = __fmul_rn(p, (float)m);
	uint32_t bits = (uint32_t)__float_as_int(r);
	return (r != r) ? 0xffc00000u : bits;
}

__device__ __forceinline__ int smear_down(int a, bool wide)
{
	a |= a >> 1;
	a |= a >> 2;
	a |= a >> 4;
	if (wide)
		a |= a >> 8;
	return a;
}

// reference src/vertexfilter.cpp:445-488; element = 4 bytes [y co cg a]
__device__ __forceinline__ uint32_t filter_color8(uint32_t v)
{
	int y = (int)(v & 0xff);
	int co = (int)(signed char)((v >> 8) & 0xff);
	int cg = (int)(signed char)((v >> 16) & 0xff);
	int a = (int)(v >> 24);
	int as = smear_down(a, false);
	a = ((a << 1) & as) | (a & 1);
	float ss = __fdiv_rn(255.f, (float)as);

	int r = y + (co - cg);
	int g = y + cg;
	int b = y - (co + cg);

	// lanes are OR-ed together unmasked, as the reference does (:477-481)
	uint32_t res = (uint32_t)cvt_rne(__fmul_rn((float)r, ss));
	res |= (uint32_t)cvt_rne(__fmul_rn((float)g, ss)) << 8;
	res |= (uint32_t)cvt_rne(__fmul_rn((float)b, ss)) << 16;
	res |= (uint32_t)cvt_rne(__fmul_rn((float)a, ss)) << 24;
	return res;
}

// reference src/vertexfilter.cpp:490-542; element = 8 bytes, lo = y | co<<16, hi = cg | a<<16
__device__ __forceinline__ uint2 filter_color16(uint2 v)
{
	int y = (int)(v.x & 0xffff);
	int co = (int)v.x >> 16;
	int cg = (int)(short)(v.y & 0xffff);
	int a = (int)(v.y >> 16);
	int as = smear_down(a, true);
	a = ((a << 1) & as) | (a & 1);
	float ss = __fdiv_rn(65535.f, (float)as);

	int r = y + (co - cg);
	int g = y + cg;
	int b = y - (co + cg);

	uint32_t rr = (uint32_t)cvt_rne(__fmul_rn((float)r, ss));
	uint32_t gr = (uint32_t)cvt_rne(__fmul_rn((float)g, ss));
	uint32_t br = (uint32_t)cvt_rne(__fmul_rn((float)b, ss));
	uint32_t ar = (uint32_t)cvt_rne(__fmul_rn((float)a, ss));
	return make_uint2((rr & 0xffffu) | (gr << 16), (br & 0xffffu) | (ar << 16));
}

// element-wise dispatch on 4-byte elements / words (Exp works on every 32-bit word of any stride)
__device__ __noinline__ uint32_t apply_filter32(uint32_t v, int filter)
{
	switch (filter)
	{
	case 1: return filter_oct8(v);
	case 3: return filter_exp(v);
	case 4: return filter_color8(v);
	default: return v;
	}
}

__device__ __noinline__ uint2 apply_filter64(uint2 v, int filter)
{
	switch (filter)
	{
	case 1: return filter_oct16(v);
	case 2: return filter_quat(v);
	case 4: return filter_color16(v);
	default: return v;
	}
}

// a 16-byte piece at a time: the four (two) independent elements are interleaved by the compiler inside ONE call -- as
// single-element calls the decoders' store pass ran them back to back, each a dependent chain with a division and a
// square root (unit trace, 4-byte normals: 3.1 us of a round's 8 us)
__device__ __noinline__ uint4 apply_filter32x4(uint4 v, int filter)
{
	switch (filter)
	{
	case 1: return make_uint4(filter_oct8(v.x), filter_oct8(v.y), filter_oct8(v.z), filter_oct8(v.w));
	case 3: return make_uint4(filter_exp(v.x), filter_exp(v.y), filter_exp(v.z), filter_exp(v.w));
	case 4: return make_uint4(filter_color8(v.x), filter_color8(v.y), filter_color8(v.z), filter_color8(v.w));
	default: return v;
	}
}

__device__ __noinline__ uint4 apply_filter64x2(uint4 v, int filter)
{
	uint2 a = make_uint2(v.x, v.y), b = make_uint2(v.z, v.w);
	switch (filter)
	{
	case 1: a = filter_oct16(a), b = filter_oct16(b); break;
	case 2: a = filter_quat(a), b = filter_quat(b); break;
	case 4: a = filter_color16(a), b = filter_color16(b); break;
	default: break;
	}
	return make_uint4(a.x, a.y, b.x, b.y);
}

} // namespace mob200
