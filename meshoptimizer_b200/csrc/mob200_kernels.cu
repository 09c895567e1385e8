// mob200_kernels.cu -- sm_100a kernels of the vertex-buffer decode path.
//
// One persistent kernel, one CTA per SM, five units of three warp roles each per CTA (SURVEY.md section 7.4,
// north_star phases 1-3):
//
//   walker warp    phase 1.  One LANE per stream walks the group headers (the stream stores no index:
//                  reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance a single
//                  pointer) and publishes, block by block, the byte offset of the block and of each of
//                  its 16-value groups, plus the reference return code of the stream (:1827-1869).
//                  Chains of different streams are independent, so 32 of them advance per warp
//                  instruction; a release store per block hands the block to the producers.
//   producer warp  stages the CTA's next blocks (level-major order: block b of every stream before block
//                  b+1 of any, i.e. the order the walkers produce them): TMA bulk copies of the encoded
//                  bytes and of the group-table rows into a shared-memory ring, and the cross-block carry
//                  (last_vertex, :1592,:1848-1849) by a single-pass decoupled look-back per 4-byte lane.
//   decoder warps  phases 2+3.  Thread-per-(4 byte-channels x 16 vertices): unpack of four 16-value groups
//                  in registers (decodeBytesGroup, :582-641), 4x4 byte transposes back to interleaved
//                  vertices, un-zigzag / rotate and scan of the deltas (decodeDeltas1, :669-699) with warp
//                  shuffles across the block, the decode filter (src/vertexfilter.cpp) as an epilogue on the
//                  finished words, and 16-byte coalesced stores from a padded tile.
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
#include "mob200_walk_team.cuh"

namespace mob200
{

static_assert(kSmemWalkerBytes >= kWalkSmemBytes && kSmemWalkerBytes >= kWideSmemBytes, "walker shared-memory region");

// kWideWalk: one walker warp per stream (few streams) instead of one lane per stream.
// kRounds:   blocks of small vertices (<= 16 bytes: one or two work quanta) are decoded up to four at a time by a unit's
//            four decoder warps (Lay<true>); otherwise one block at a time goes round the four warps.
// kBlock:    block mode (DevTables::block_mode): one walker lane per BLOCK from the block-offset table, look-back over
//            several predecessors per step.  A template parameter so that a kernel carries the code of one mode only.
template <bool kWideWalk, bool kRounds, bool kBlock>
__global__ void __launch_bounds__(kCtaThreads, 1) decode_kernel(DevTables T)
{
	using L = Lay<kRounds>;
	constexpr uint32_t kSlots = L::kSlots;
	extern __shared__ __align__(1024) uint8_t smem_all[];

	const bool decode_on = T.walker_lead != kWalkOnly;
	const bool walk_on = T.walker_lead != kDecodeOnly;

	// role and unit of this warp: decoder warps of unit 0..4, then the producers, then the walkers
	const uint32_t tid = threadIdx.x;
	uint32_t role, g, utid;
	if (tid < kFirstProducerThread)
		role = 0, g = tid / kDecodeThreads, utid = tid % kDecodeThreads;
	else if (tid < kFirstWalkerThread)
		role = 1, g = (tid - kFirstProducerThread) / kProducerThreads, utid = tid % kProducerThreads;
	else
		role = 2, g = (tid - kFirstWalkerThread) / kWalkerThreads, utid = tid % kWalkerThreads;
	// units of one CTA are far apart in the decode order (unit = g * gridDim.x + blockIdx.x): a batch with fewer
	// units than the device could hold still spreads over all SMs
	const uint32_t unit = g * gridDim.x + blockIdx.x;
	const bool unit_on = unit < T.units;
	uint8_t* smem = smem_all + g * L::kSmemTotal;

	if (decode_on)
	{
		if (role == 0 && utid >= 32 && utid < 48)
			reinterpret_cast<uint32_t*>(smem + L::kSmemPatch)[utid - 32] = patch_selector(utid - 32);
		if (role == 0 && utid == 0)
		{
			uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kSmemBars);
			for (uint32_t k = 0; k < kSlots; ++k)
			{
				mbar_init(bars + k, 1);                          // full: the producer's arrive.expect_tx
				mbar_init(bars + kSlots + k, kProducerThreads);  // carry: every lane of the producer warp
				mbar_init(bars + 2 * kSlots + k, kRounds ? kDecodeThreads : kDecodeThreads / 32); // empty: one arrival per decoder warp (plain form) or thread (rounds form)
			}
			mbar_init(bars + 3 * kSlots, kDecodeThreads);        // tile_free: one arrival per decoder thread
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
	}
	if (!unit_on)
		return;

#ifdef MOB200_DEBUG_ENDS
	const long long ends_t0 = clock64();
#endif
	if (role == 0)
	{
		if (decode_on)
			decoder_main<kRounds, kBlock>(T, smem, unit, utid, 1u + g);
	}
	else if (role == 1)
	{
		if (decode_on)
			producer_main<kRounds, kBlock>(T, smem, unit);
	}
	else if (walk_on && (g < 4 || (kBlock ? T.total_blocks : T.n_streams) > 4u * (kWideWalk ? 1u : 32u) * gridDim.x))
	{
		// (the fifth walker shares a scheduler with the first: it only runs when four per SM cannot take every
		// stream in one round -- one walker per scheduler is 9% faster alone and 3% faster fused)
		if (kWideWalk)
			walker_main_wide(T, smem + L::kSmemWalker);
		else
			walker_main<kBlock>(T, smem + L::kSmemWalker);
	}

#ifdef MOB200_DEBUG_ENDS
	// diagnostics: when did the roles finish (cycles since the start of the CTA): slots 13 walker sum, 14 decoder max, 15 walker max
	if (utid == 0 && role != 1)
	{
		unsigned long long* dbg = reinterpret_cast<unsigned long long*>(T.counters + 16);
		const unsigned long long dt = (unsigned long long)(clock64() - ends_t0);
		if (role == 2)
		{
			atomicAdd(dbg + 13, dt);
			atomicMax(dbg + 15, dt);
		}
		else
			atomicMax(dbg + 14, dt);
		if (unit == 0 && role == 0)
			dbg[12] = g_dbg_spins; // (cumulative; read it as a difference between runs)
	}
#endif
	// the last role to finish re-arms the counters for the next launch (stream order makes this visible)
	if (utid == 0 && role != 1)
	{
		__threadfence();
		uint32_t finished = atomicAdd(T.counters + 2, 1u);
		if (finished == 2 * T.units - 1)
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

	// 16-byte aligned buffers (every cudaMalloc'ed one): 16 bytes per thread and iteration -- four independent words or
	// two 8-byte elements -- with full-width loads and stores; the last few bytes go through the element paths below
	if ((addr & 15) == 0)
	{
		const bool words = filter == MOB200_FILTER_EXP || stride == 4;
		const size_t bytes = count * stride;
		const size_t vecs = bytes >> 4;
		uint4* p = reinterpret_cast<uint4*>(data);
		for (size_t j = i; j < vecs; j += step)
		{
			uint4 v = p[j];
			if (words)
			{
				v.x = apply_filter32(v.x, filter), v.y = apply_filter32(v.y, filter);
				v.z = apply_filter32(v.z, filter), v.w = apply_filter32(v.w, filter);
			}
			else
			{
				const uint2 lo = apply_filter64(make_uint2(v.x, v.y), filter), hi = apply_filter64(make_uint2(v.z, v.w), filter);
				v = make_uint4(lo.x, lo.y, hi.x, hi.y);
			}
			p[j] = v;
		}
		// tail: fewer than 16 bytes
		const size_t done = vecs << 4;
		if (words)
		{
			uint32_t* w = reinterpret_cast<uint32_t*>(data + done);
			const size_t rest = (bytes - done) >> 2;
			if (i < rest)
				w[i] = apply_filter32(w[i], filter);
		}
		else if (i == 0 && bytes - done >= 8)
		{
			uint2* e = reinterpret_cast<uint2*>(data + done);
			*e = apply_filter64(*e, filter);
		}
		return;
	}

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

template <bool kWideWalk, bool kRounds, bool kBlock>
static cudaError_t prepare_one()
{
	return cudaFuncSetAttribute(decode_kernel<kWideWalk, kRounds, kBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<kRounds>::kSmemCta);
}

cudaError_t prepare_decode_kernel()
{
	cudaError_t err = prepare_one<false, false, false>();
	if (err == cudaSuccess)
		err = prepare_one<true, false, false>();
	if (err == cudaSuccess)
		err = prepare_one<false, true, false>();
	if (err == cudaSuccess)
		err = prepare_one<true, true, false>();
	if (err == cudaSuccess)
		err = prepare_one<false, false, true>();
	if (err == cudaSuccess)
		err = prepare_one<false, true, true>();
	return err;
}

cudaError_t decode_occupancy(int* ctas_per_sm)
{
	int a = 0, b = 0, c = 0, d = 0;
	cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, decode_kernel<false, false, false>, kCtaThreads, Lay<false>::kSmemCta);
	if (err == cudaSuccess)
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, decode_kernel<false, true, false>, kCtaThreads, Lay<true>::kSmemCta);
	if (err == cudaSuccess)
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, decode_kernel<false, false, true>, kCtaThreads, Lay<false>::kSmemCta);
	if (err == cudaSuccess)
		err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, decode_kernel<false, true, true>, kCtaThreads, Lay<true>::kSmemCta);
	int m = a < b ? a : b;
	m = m < c ? m : c;
	*ctas_per_sm = m < d ? m : d;
	return err;
}

cudaError_t launch_decode(const DevTables& T, uint32_t grid, cudaStream_t stream)
{
	if (T.n_streams == 0)
		return cudaSuccess;
	if (T.block_mode)
	{
		// (one walker lane per block: the wide form has no part in it)
		if (T.rounds)
			decode_kernel<false, true, true><<<grid, kCtaThreads, Lay<true>::kSmemCta, stream>>>(T);
		else
			decode_kernel<false, false, true><<<grid, kCtaThreads, Lay<false>::kSmemCta, stream>>>(T);
	}
	else if (T.rounds)
	{
		if (T.wide_walk)
			decode_kernel<true, true, false><<<grid, kCtaThreads, Lay<true>::kSmemCta, stream>>>(T);
		else
			decode_kernel<false, true, false><<<grid, kCtaThreads, Lay<true>::kSmemCta, stream>>>(T);
	}
	else
	{
		if (T.wide_walk)
			decode_kernel<true, false, false><<<grid, kCtaThreads, Lay<false>::kSmemCta, stream>>>(T);
		else
			decode_kernel<false, false, false><<<grid, kCtaThreads, Lay<false>::kSmemCta, stream>>>(T);
	}
	return cudaGetLastError();
}

cudaError_t launch_walk_team(const DevTables& T, int sm_count, cudaStream_t stream)
{
	if (T.n_streams == 0)
		return cudaSuccess;
	// one CTA per stream, at most what the device holds at once (the rest by grid stride)
	const uint32_t cap = (uint32_t)sm_count * 10u;
	const uint32_t grid = T.n_streams < cap ? T.n_streams : cap;
	walk_team_kernel<<<grid, kTeamThreads, kTeamSmemBytes, stream>>>(T);
	return cudaGetLastError();
}

cudaError_t launch_filter(int filter, void* data, size_t count, size_t stride, int sm_count, cudaStream_t stream)
{
	if (count == 0)
		return cudaSuccess;
	size_t work = (count * stride + 15) / 16; // 16 bytes per thread and iteration on the aligned path
	size_t blocks = (work + 255) / 256;
	size_t cap = (size_t)sm_count * 8; // a multiple of the SM count: 8 resident CTAs of 256 threads per SM
	uint32_t grid = (uint32_t)(blocks < cap ? blocks : cap);
	filter_kernel<<<grid, 256, 0, stream>>>(static_cast<uint8_t*>(data), count, (uint32_t)stride, filter);
	return cudaGetLastError();
}

} // namespace mob200
