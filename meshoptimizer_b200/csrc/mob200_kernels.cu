// mob200_kernels.cu -- sm_100a kernels of the vertex-buffer decode path.
//
// One persistent kernel, two warp roles per CTA (north_star phases 1-3):
//
//   walker warp    phase 1.  One LANE per stream walks the group headers (the stream stores no index:
//                  reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance a single
//                  pointer) and publishes, block by block, the byte offset of the block and of each of
//                  its 16-value groups, plus the reference return code of the stream (:1827-1869).
//                  Chains of different streams are independent, so 32 of them advance per warp
//                  instruction; a release store per block hands the block to the decoders.
//   decoder warps  phases 2+3.  The four decoder warps of a CTA take blocks from a ticket counter in
//                  level-major order (block b of every stream before block b+1 of any), i.e. in the
//                  order the walkers produce them.  Per block: one TMA bulk copy stages the encoded
//                  bytes in shared memory; thread-per-group unpack into byte planes (the work of
//                  decodeBytesGroup, :582-641); in-register 4x4 byte transposes back to interleaved
//                  vertices; un-zigzag / rotate and an in-block scan of the deltas (decodeDeltas1,
//                  :669-699); the cross-block carry (last_vertex, :1592,:1848-1849) is a single-pass
//                  decoupled look-back per 4-byte lane; the decode filter (src/vertexfilter.cpp) runs
//                  as an epilogue on the finished tile; 16-byte coalesced stores write the vertices.
//
//   filter_kernel  standalone meshopt_decodeFilter* on a device buffer.
//
// Everything is integer/byte work bound by HBM traffic; no tensor cores are involved.
#include "mob200_kernels.h"

#include "mob200_decoder.cuh"
#include "mob200_device.cuh"
#include "mob200_filters.cuh"
#include "mob200_walker.cuh"
#include "mob200_walker_wide.cuh"

namespace mob200
{

__global__ void __launch_bounds__(kCtaThreads, 8) decode_kernel(DevTables T)
{
	extern __shared__ __align__(128) uint8_t smem[];

	if (threadIdx.x < kDecodeThreads)
	{
		if (T.walker_lead != 0xffffffffu) // 0xffffffff: walk-only diagnostic mode (MOB200_WALKER_LEAD=4294967295)
			decoder_main(T, smem);
	}
	else if (T.wide_walk)
		walker_main_wide(T, smem + kSmemRing);
	else
		walker_main(T, smem + kSmemRing, smem + kSmemRows);

	// the last role to finish re-arms the counters for the next launch (stream order makes this visible)
	if ((threadIdx.x & 31u) == 0 && (threadIdx.x == 0 || threadIdx.x == kDecodeThreads))
	{
		__threadfence();
		uint32_t finished = atomicAdd(T.counters + 2, 1u);
		if (finished == 2 * gridDim.x - 1)
		{
			T.counters[0] = 0;
			T.counters[1] = 0;
			T.counters[2] = 0;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// standalone filters (meshopt_decodeFilter* on device memory)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) filter_kernel(uint8_t* data, size_t count, uint32_t stride, int filter)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t step = (size_t)gridDim.x * blockDim.x;
	const uintptr_t addr = reinterpret_cast<uintptr_t>(data);

	if (filter == MOB200_FILTER_EXP || stride == 4)
	{
		// independent 32-bit words
		const size_t words = filter == MOB200_FILTER_EXP ? count * (stride / 4) : count;
		if ((addr & 3) == 0)
		{
			uint32_t* p = reinterpret_cast<uint32_t*>(data);
			for (; i < words; i += step)
				p[i] = apply_filter32(p[i], filter);
		}
		else
		{
			for (; i < words; i += step)
			{
				uint8_t* b = data + i * 4;
				uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
				v = apply_filter32(v, filter);
				b[0] = (uint8_t)v;
				b[1] = (uint8_t)(v >> 8);
				b[2] = (uint8_t)(v >> 16);
				b[3] = (uint8_t)(v >> 24);
			}
		}
	}
	else
	{
		// 8-byte elements
		if ((addr & 7) == 0)
		{
			uint2* p = reinterpret_cast<uint2*>(data);
			for (; i < count; i += step)
				p[i] = apply_filter64(p[i], filter);
		}
		else
		{
			for (; i < count; i += step)
			{
				uint8_t* b = data + i * 8;
				uint2 v;
				v.x = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
				v.y = (uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24);
				v = apply_filter64(v, filter);
				for (int j = 0; j < 4; ++j)
				{
					b[j] = (uint8_t)(v.x >> (8 * j));
					b[4 + j] = (uint8_t)(v.y >> (8 * j));
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------

cudaError_t prepare_decode_kernel()
{
	return cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
}

cudaError_t decode_occupancy(int* ctas_per_sm)
{
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, decode_kernel, kCtaThreads, kSmemTotal);
}

cudaError_t launch_decode(const DevTables& T, uint32_t grid, cudaStream_t stream)
{
	if (T.n_streams == 0)
		return cudaSuccess;
	decode_kernel<<<grid, kCtaThreads, kSmemTotal, stream>>>(T);
	return cudaGetLastError();
}

cudaError_t launch_filter(int filter, void* data, size_t count, size_t stride, int sm_count, cudaStream_t stream)
{
	if (count == 0)
		return cudaSuccess;
	size_t work = filter == MOB200_FILTER_EXP ? count * (stride / 4) : count;
	size_t blocks = (work + 255) / 256;
	size_t cap = (size_t)sm_count * 8;
	uint32_t grid = (uint32_t)(blocks < cap ? blocks : cap);
	filter_kernel<<<grid, 256, 0, stream>>>(static_cast<uint8_t*>(data), count, (uint32_t)stride, filter);
	return cudaGetLastError();
}

} // namespace mob200
