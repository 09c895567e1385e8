/*
 * oracle/vertexfilter_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the meshoptimizer decode filters (octahedral, quaternion, exponential,
 * colour).  Only tests/, __graft_entry__.smoke() and the cpu_baseline leg of bench.py may use it.
 *
 * The restatement follows the operation ORDER of the reference's x86 SSE2 kernels
 * (/root/reference/src/vertexfilter.cpp:256-542), because that is the code path tools/codecbench
 * runs and the one the parity contract names; every +,-,*,/ and sqrt below
 * is one separately rounded IEEE-754 binary32 operation.  BUILD WITH -ffp-contract=off.
 *
 * Parity status: PINNED against the reference's known-answer vectors (demo/tests.cpp:762-876,
 * js/meshopt_decoder.test.js:217-308) and, when oracle/_ref is present, against the reference
 * itself on encoder-produced inputs (tests/test_oracle_cpu.py):
 *   bit-exact lanes : Exp, Oct (stride 8), Quat, Color (stride 8)
 *   <=1 LSB lanes   : Oct (stride 4) and Color (stride 4) -- the reference uses the hardware
 *                     approximations rsqrtps / rcpps there (vertexfilter.cpp:284,467); this file
 *                     (and the CUDA path) use the correctly rounded 127/sqrt(ll) and 255/as.
 *
 * Conversions: cvtdq2ps is exact int->float (round to nearest even); cvtps2dq is round-half-even
 * with the x86 "integer indefinite" 0x80000000 for NaN and out-of-range inputs.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static int32_t cvt_rne(float x)
{
	if (!(x >= -2147483648.f && x < 2147483648.f))
		return INT32_MIN;
	return (int32_t)nearbyintf(x); /* default rounding mode: to nearest even */
}

static float flip_sign_like(float t, float x)
{
	/* t XOR signbit(x) */
	uint32_t tb, xb;
	memcpy(&tb, &t, 4);
	memcpy(&xb, &x, 4);
	tb ^= xb & 0x80000000u;
	memcpy(&t, &tb, 4);
	return t;
}

static float sse_min(float a, float b) { return a < b ? a : b; } /* minps: second operand on NaN */
static float sse_max(float a, float b) { return a > b ? a : b; }

/* shared octahedral core: x,y,z already converted to float, z already has (|x|+|y|) subtracted */
static void oct_unfold(float* x, float* y, float z)
{
	float t = sse_min(z, 0.f);
	*x = *x + flip_sign_like(t, *x);
	*y = *y + flip_sign_like(t, *y);
}

/* reference vertexfilter.cpp:256-299 (with exact 127/sqrt in place of 127*rsqrtps) */
static void filter_oct8(uint8_t* data, size_t count)
{
	for (size_t i = 0; i < count; ++i)
	{
		uint8_t* e = data + i * 4;
		float x = (float)(int8_t)e[0];
		float y = (float)(int8_t)e[1];
		float z = (float)(int8_t)e[2] - (fabsf(x) + fabsf(y));
		oct_unfold(&x, &y, z);
		float ll = x * x + (y * y + z * z);
		float s = 127.f / sqrtf(ll);
		e[0] = (uint8_t)cvt_rne(x * s);
		e[1] = (uint8_t)cvt_rne(y * s);
		e[2] = (uint8_t)cvt_rne(z * s);
	}
}

/* reference vertexfilter.cpp:301-357 */
static void filter_oct16(uint16_t* data, size_t count)
{
	for (size_t i = 0; i < count; ++i)
	{
		uint16_t* e = data + i * 4;
		float x = (float)(int16_t)e[0];
		float y = (float)(int16_t)e[1];
		float z = (float)(int32_t)(e[2] & 0x7fff) - (fabsf(x) + fabsf(y));
		oct_unfold(&x, &y, z);
		float ll = x * x + (y * y + z * z);
		float s = 32767.f / sqrtf(ll);
		e[0] = (uint16_t)cvt_rne(x * s);
		e[1] = (uint16_t)cvt_rne(y * s);
		e[2] = (uint16_t)cvt_rne(z * s);
	}
}

/* reference vertexfilter.cpp:359-423 */
static void filter_quat(uint16_t* data, size_t count)
{
	volatile float two = 2.f; /* the reference computes 32767/sqrt(2) in float at run time (:361) */
	const float scale = 32767.f / sqrtf(two);

	for (size_t i = 0; i < count; ++i)
	{
		uint16_t* e = data + i * 4;
		int32_t c = (int16_t)e[3];
		float x = (float)(int16_t)e[0];
		float y = (float)(int16_t)e[1];
		float z = (float)(int16_t)e[2];
		float s = (float)(c | 3);

		float ws = s * (s + s);
		float ww = ws - (x * x + (y * y + z * z));
		float w = sqrtf(sse_max(ww, 0.f));
		float ss = scale / s;

		uint64_t xr = (uint16_t)cvt_rne(x * ss);
		uint64_t yr = (uint16_t)cvt_rne(y * ss);
		uint64_t zr = (uint16_t)cvt_rne(z * ss);
		uint64_t wr = (uint16_t)cvt_rne(w * ss);

		/* lanes [w x y z], rotated left by 16*(c&3) bits */
		uint64_t packed = wr | (xr << 16) | (yr << 32) | (zr << 48);
		unsigned r = ((unsigned)c << 4) & 63;
		uint64_t rotated = r ? (packed << r) | (packed >> (64 - r)) : packed;

		e[0] = (uint16_t)rotated;
		e[1] = (uint16_t)(rotated >> 16);
		e[2] = (uint16_t)(rotated >> 32);
		e[3] = (uint16_t)(rotated >> 48);
	}
}

/* reference vertexfilter.cpp:425-443 */
static void filter_exp(uint32_t* data, size_t count)
{
	for (size_t i = 0; i < count; ++i)
	{
		uint32_t v = data[i];
		int32_t e = (int32_t)v >> 24;
		int32_t m = (int32_t)(v << 8) >> 8;
		uint32_t pow2 = (uint32_t)(e + 127) << 23;
		float p;
		memcpy(&p, &pow2, 4);
		float r = p * (float)m;
		uint32_t rb;
		memcpy(&rb, &r, 4);
		if (r != r)
			rb = 0xffc00000u; /* x86 default NaN for inf*0; only e == -128 with m == 0 gets here */
		data[i] = rb;
	}
}

static int32_t smear_down(int32_t a, int bits16)
{
	a |= a >> 1;
	a |= a >> 2;
	a |= a >> 4;
	if (bits16)
		a |= a >> 8;
	return a;
}

/* reference vertexfilter.cpp:445-488 (with exact 255/as in place of 255*rcpps) */
static void filter_color8(uint8_t* data, size_t count)
{
	for (size_t i = 0; i < count; ++i)
	{
		uint8_t* e = data + i * 4;
		int32_t y = e[0], co = (int8_t)e[1], cg = (int8_t)e[2], a = e[3];
		int32_t as = smear_down(a, 0);
		a = ((a << 1) & as) | (a & 1);
		float ss = 255.f / (float)as;

		int32_t r = y + (co - cg);
		int32_t g = y + cg;
		int32_t b = y - (co + cg);

		/* the reference ORs unmasked lanes together (:477-481) */
		uint32_t res = (uint32_t)cvt_rne((float)r * ss);
		res |= (uint32_t)cvt_rne((float)g * ss) << 8;
		res |= (uint32_t)cvt_rne((float)b * ss) << 16;
		res |= (uint32_t)cvt_rne((float)a * ss) << 24;

		e[0] = (uint8_t)res;
		e[1] = (uint8_t)(res >> 8);
		e[2] = (uint8_t)(res >> 16);
		e[3] = (uint8_t)(res >> 24);
	}
}

/* reference vertexfilter.cpp:490-542 */
static void filter_color16(uint16_t* data, size_t count)
{
	for (size_t i = 0; i < count; ++i)
	{
		uint16_t* e = data + i * 4;
		int32_t y = e[0], co = (int16_t)e[1], cg = (int16_t)e[2], a = e[3];
		int32_t as = smear_down(a, 1);
		a = ((a << 1) & as) | (a & 1);
		float ss = 65535.f / (float)as;

		int32_t r = y + (co - cg);
		int32_t g = y + cg;
		int32_t b = y - (co + cg);

		e[0] = (uint16_t)cvt_rne((float)r * ss);
		e[1] = (uint16_t)cvt_rne((float)g * ss);
		e[2] = (uint16_t)cvt_rne((float)b * ss);
		e[3] = (uint16_t)cvt_rne((float)a * ss);
	}
}

/* Public entry points: same contracts as meshopt_decodeFilter* (reference src/meshoptimizer.h:421-424,
 * dispatch in vertexfilter.cpp:1211-1274).  Results do not depend on count%4 (the reference's SIMD
 * tail path, :217-241, pads with zero elements whose results are discarded). */
ORACLE_API void oracle_decodeFilterOct(void* buffer, size_t count, size_t stride)
{
	if (stride == 4)
		filter_oct8((uint8_t*)buffer, count);
	else if (stride == 8)
		filter_oct16((uint16_t*)buffer, count);
}

ORACLE_API void oracle_decodeFilterQuat(void* buffer, size_t count, size_t stride)
{
	if (stride == 8)
		filter_quat((uint16_t*)buffer, count);
}

ORACLE_API void oracle_decodeFilterExp(void* buffer, size_t count, size_t stride)
{
	if (stride > 0 && stride % 4 == 0)
		filter_exp((uint32_t*)buffer, count * (stride / 4));
}

ORACLE_API void oracle_decodeFilterColor(void* buffer, size_t count, size_t stride)
{
	if (stride == 4)
		filter_color8((uint8_t*)buffer, count);
	else if (stride == 8)
		filter_color16((uint16_t*)buffer, count);
}
