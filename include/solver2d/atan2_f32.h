// solver2d-b200 — single-precision atan2 with the exact rounding behaviour of the C library the reference is built
// against.
//
// Why: the joint-limit rows of every solver variant read the relative joint angle through atan2f (reference
// include/solver2d/math.h:320-327 -> src/revolute_joint.c limit blocks). atan2f is NOT correctly rounded in glibc up
// to 2.40 — it is the classic fdlibm float algorithm (argument reduction against atan(0.5), atan(1), atan(1.5),
// atan(inf) plus an 11-term odd/even split polynomial, all in float arithmetic) — and CUDA's atan2f is a different
// approximation, so the two differ in the last bit on ~15 % of inputs. One differing bit in an active limit row is
// enough to break bit-exact parity of a whole joint chain. s2Atan2F32 therefore restates the published fdlibm
// algorithm (Sun Microsystems' e_atan2f / s_atanf, as shipped in glibc 2.39 sysdeps/ieee754/flt-32) in plain float
// operations; compiled with -fmad=false it returns the same bits as the host libm, which tests/test_host_cpu.py
// checks over tens of millions of inputs through the s2Atan2Device export.
//
// Used by the device code only; host code keeps calling atan2f like the reference does.
#pragma once

#include "solver2d/types.h"

#include <stdint.h>

S2_INLINE int32_t s2FloatBits(float f)
{
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	union
	{
		float f;
		int32_t i;
	} u;
	u.f = f;
	return u.i;
#endif
}

S2_INLINE float s2BitsFloat(int32_t i)
{
#if defined(__CUDA_ARCH__)
	return __int_as_float(i);
#else
	union
	{
		float f;
		int32_t i;
	} u;
	u.i = i;
	return u.f;
#endif
}

// atan(x), float: reduce to |t| < 7/16 around one of four break points, then x - x*P(x^2) with the break point's
// hi/lo parts added back.
S2_INLINE float s2AtanF32(float x)
{
	// atan(0.5), atan(1), atan(1.5), atan(inf): high parts, then the low-order remainders
	const float hi0 = s2BitsFloat(0x3eed6338), hi1 = s2BitsFloat(0x3f490fda), hi2 = s2BitsFloat(0x3f7b985e),
				hi3 = s2BitsFloat(0x3fc90fda);
	const float lo0 = s2BitsFloat(0x31ac3769), lo1 = s2BitsFloat(0x33222168), lo2 = s2BitsFloat(0x33140fb4),
				lo3 = s2BitsFloat(0x33a22168);
	// odd polynomial coefficients of atan(x)/x - 1 in x^2
	const float a0 = s2BitsFloat(0x3eaaaaab), a1 = s2BitsFloat((int32_t)0xbe4ccccd), a2 = s2BitsFloat(0x3e124925),
				a3 = s2BitsFloat((int32_t)0xbde38e38), a4 = s2BitsFloat(0x3dba2e6e), a5 = s2BitsFloat((int32_t)0xbd9d8795),
				a6 = s2BitsFloat(0x3d886b35), a7 = s2BitsFloat((int32_t)0xbd6ef16b), a8 = s2BitsFloat(0x3d4bda59),
				a9 = s2BitsFloat((int32_t)0xbd15a221), a10 = s2BitsFloat(0x3c8569d7);

	int32_t hx = s2FloatBits(x);
	int32_t ix = hx & 0x7fffffff;
	int id;
	float hi = 0.0f, lo = 0.0f;
	if (ix >= 0x4c000000) // |x| >= 2^25
	{
		if (ix > 0x7f800000)
		{
			return x + x; // NaN
		}
		return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
	}
	if (ix < 0x3ee00000) // |x| < 7/16
	{
		if (ix < 0x31000000) // |x| < 2^-29
		{
			return x;
		}
		id = -1;
	}
	else
	{
		x = s2BitsFloat(ix); // |x|
		if (ix < 0x3f980000) // |x| < 19/16
		{
			if (ix < 0x3f300000) // 7/16 <= |x| < 11/16
			{
				id = 0;
				hi = hi0;
				lo = lo0;
				x = (2.0f * x - 1.0f) / (2.0f + x);
			}
			else
			{
				id = 1;
				hi = hi1;
				lo = lo1;
				x = (x - 1.0f) / (x + 1.0f);
			}
		}
		else if (ix < 0x401c0000) // |x| < 39/16
		{
			id = 2;
			hi = hi2;
			lo = lo2;
			x = (x - 1.5f) / (1.0f + 1.5f * x);
		}
		else
		{
			id = 3;
			hi = hi3;
			lo = lo3;
			x = -1.0f / x;
		}
	}
	float z = x * x;
	float w = z * z;
	float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
	float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
	if (id < 0)
	{
		return x - x * (s1 + s2);
	}
	z = hi - ((x * (s1 + s2) - lo) - x);
	return hx < 0 ? -z : z;
}

// atan2(y, x), float: quadrant bookkeeping around s2AtanF32(|y / x|) with pi split into hi/lo parts.
S2_INLINE float s2Atan2F32(float y, float x)
{
	const float tiny = 1.0e-30f;
	const float pi_o_4 = s2BitsFloat(0x3f490fdb), pi_o_2 = s2BitsFloat(0x3fc90fdb), pi = s2BitsFloat(0x40490fdb),
				pi_lo = s2BitsFloat((int32_t)0xb3bbbd2e);
	int32_t hx = s2FloatBits(x), hy = s2FloatBits(y);
	int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
	if (ix > 0x7f800000 || iy > 0x7f800000)
	{
		return x + y; // NaN
	}
	if (hx == 0x3f800000)
	{
		return s2AtanF32(y); // x == 1
	}
	int m = ((hy >> 31) & 1) | ((hx >> 30) & 2); // 2 * sign(x) + sign(y)
	if (iy == 0)
	{
		switch (m)
		{
			case 0:
			case 1:
				return y;
			case 2:
				return pi + tiny;
			default:
				return -pi - tiny;
		}
	}
	if (ix == 0)
	{
		return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	}
	if (ix == 0x7f800000)
	{
		if (iy == 0x7f800000)
		{
			switch (m)
			{
				case 0:
					return pi_o_4 + tiny;
				case 1:
					return -pi_o_4 - tiny;
				case 2:
					return 3.0f * pi_o_4 + tiny;
				default:
					return -3.0f * pi_o_4 - tiny;
			}
		}
		switch (m)
		{
			case 0:
				return 0.0f;
			case 1:
				return -0.0f;
			case 2:
				return pi + tiny;
			default:
				return -pi - tiny;
		}
	}
	if (iy == 0x7f800000)
	{
		return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
	}
	int k = (iy - ix) >> 23;
	float z;
	if (k > 60)
	{
		z = pi_o_2 + 0.5f * pi_lo; // |y / x| > 2^60
	}
	else if (hx < 0 && k < -60)
	{
		z = 0.0f; // |y| / x < -2^60
	}
	else
	{
		float q = y / x;
		z = s2AtanF32(s2BitsFloat(s2FloatBits(q) & 0x7fffffff));
	}
	switch (m)
	{
		case 0:
			return z;
		case 1:
			return s2BitsFloat(s2FloatBits(z) ^ (int32_t)0x80000000);
		case 2:
			return pi - (z - pi_lo);
		default:
			return (z - pi_lo) - pi;
	}
}
