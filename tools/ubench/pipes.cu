// Micro-benchmark: issue throughput of the integer instructions the decoder is made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes
// Every warp runs 8 independent dependency chains of one instruction kind; 1 CTA of 1024 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kOp>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b)
{
	uint32_t r;
	if (kOp == 0) asm volatile("lop3.b32 %0, %1, %2, 0x0f0f0f0f, 0x6a;" : "=r"(r) : "r"(a), "r"(b));
	else if (kOp == 1) asm volatile("prmt.b32 %0, %1, %2, 0x6240;" : "=r"(r) : "r"(a), "r"(b));
	else if (kOp == 2) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (kOp == 3) asm volatile("vadd2.u32.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0u)); // (video form, may be emulated)
	else if (kOp == 4) r = __vadd2(a, b);
	else if (kOp == 5) asm volatile("mad.lo.u32 %0, %1, 0x01010101, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (kOp == 6) asm volatile("shf.r.wrap.b32 %0, %1, %2, 8;" : "=r"(r) : "r"(a), "r"(b));
	else if (kOp == 7) r = __dp4a(a, 0x01010101u, b);
	else if (kOp == 8) asm volatile("mul.hi.u32 %0, %1, 0x02040810;" : "=r"(r) : "r"(a + b));
	else if (kOp == 9) r = __popc(a) + b;
	else if (kOp == 10) asm volatile("shr.u32 %0, %1, 1;" : "=r"(r) : "r"(a ^ b));
	else r = a;
	return r;
}

template <int kOp>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, int iters)
{
	uint32_t x[8];
#pragma unroll
	for (int i = 0; i < 8; ++i)
		x[i] = threadIdx.x * 8 + i;
	const uint32_t y = blockIdx.x | 1;
	for (int it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int u = 0; u < 4; ++u)
#pragma unroll
			for (int i = 0; i < 8; ++i)
				x[i] = op<kOp>(x[i], y);
	}
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		s ^= x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kOp>
void run(const char* name, uint32_t* out, int sms)
{
	const int iters = 4096;
	cudaEvent_t a, b;
	cudaEventCreate(&a), cudaEventCreate(&b);
	k<kOp><<<sms, 1024>>>(out, 16);
	cudaEventRecord(a);
	k<kOp><<<sms, 1024>>>(out, iters);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms = 0;
	cudaEventElapsedTime(&ms, a, b);
	int clk = 0;
	cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const double warp_insts = (double)sms * 32 /*warps*/ * iters * 32.0;
	const double cycles = ms * 1e-3 * clk * 1e3;
	printf("%-28s %8.3f ms  %.3f warp-instructions / clk / SMSP (at the %d MHz attribute clock; source ops may expand to several SASS instructions)\n", name, ms, warp_insts / cycles / sms / 4, clk / 1000);
}

int main()
{
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	uint32_t* out;
	cudaMalloc(&out, (size_t)sms * 1024 * 4);
	run<0>("lop3", out, sms);
	run<1>("prmt", out, sms);
	run<2>("add.u32", out, sms);
	run<4>("__vadd2 (VIADD.16x2)", out, sms);
	run<5>("mad.lo (IMAD)", out, sms);
	run<6>("shf.r.wrap", out, sms);
	run<7>("dp4a (IDP.4A)", out, sms);
	run<8>("add + mul.hi (IMAD.HI)", out, sms);
	run<9>("popc + add", out, sms);
	run<10>("xor + shr", out, sms);
	return 0;
}
