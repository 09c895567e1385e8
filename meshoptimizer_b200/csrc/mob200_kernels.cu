// mob200_kernels.cu -- sm_100a kernels of the vertex-buffer decode path.
//
//   walk_kernel    phase 1: one lane per stream walks the group headers and recovers the byte offset
//                  of every block and of every byte-channel inside it (the stream stores no index:
//                  reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance a single
//                  pointer).  Also produces the reference return code of the stream (:1827-1869).
//   decode_kernel  phases 2+3: persistent CTAs take blocks from a ticket counter.  Per block: one TMA
//                  bulk copy stages the encoded bytes in shared memory; thread-per-group unpack into
//                  byte planes (the work of decodeBytesGroup, :582-641); in-register 4x4 byte
//                  transposes back to interleaved vertices; un-zigzag / rotate and an in-block scan of
//                  the deltas (decodeDeltas1, :669-699); the cross-block carry (last_vertex, :1592,
//                  :1848-1849) is a single-pass decoupled look-back per 4-byte lane; the decode filter
//                  (src/vertexfilter.cpp) runs as an epilogue on the finished tile; 16-byte coalesced
//                  stores write the vertices.
//   filter_kernel  standalone meshopt_decodeFilter* on a device buffer.
//
// Everything is integer/byte work bound by HBM traffic; no tensor cores are involved.
#include "mob200_kernels.h"

#include "mob200_filters.cuh"

namespace mob200
{

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_addr(bar)), "r"(parity)
	             : "memory");
	return ok != 0;
}

// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 32 bits at an arbitrary byte offset of a 4-byte aligned shared-memory array
__device__ __forceinline__ uint32_t lds_u32_at(const uint8_t* base, uint32_t off)
{
	const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (off >> 2);
	return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

// up to 32 bits at an arbitrary global address, of which the first `needed` bytes matter: two aligned
// loads at most, and never a word that holds none of the needed bytes (memory safety on truncated input)
__device__ __forceinline__ uint32_t ldg_u32_at(const uint8_t* p, uint32_t needed)
{
	uintptr_t a = reinterpret_cast<uintptr_t>(p);
	const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
	uint32_t mis = (uint32_t)(a & 3u);
	uint32_t lo = __ldg(w);
	uint32_t hi = (mis + needed > 4u) ? __ldg(w + 1) : 0u;
	return __funnelshift_r(lo, hi, mis * 8u);
}

// group width in bits from version, v1 channel control (0/1) and the 2-bit selector
__device__ __forceinline__ uint32_t group_bits(uint32_t version, uint32_t ctrl, uint32_t sel)
{
	// v0: {0,2,4,8}; v1: {0,1,2,4,8}[ctrl + sel]
	uint32_t idx = version ? ctrl + sel : (sel ? sel + 1 : 0);
	return (0x84210u >> (idx * 4)) & 0xfu;
}

// number of all-ones fields among the 16 fields of a 1/2/4-bit group whose packed bytes are (w0, w1)
__device__ __forceinline__ uint32_t count_sentinels(uint32_t bits, uint32_t w0, uint32_t w1)
{
	if (bits == 1)
		return __popc(w0 & 0xffffu);
	if (bits == 2)
		return __popc(w0 & (w0 >> 1) & 0x55555555u);
	uint32_t a = w0 & (w0 >> 1);
	uint32_t b = w1 & (w1 >> 1);
	a &= a >> 2;
	b &= b >> 2;
	return __popc(a & 0x11111111u) + __popc(b & 0x11111111u);
}

// ------------------------------------------------------------------------------------------------
// phase 1: header walk
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kWalkThreads) walk_kernel(DevTables T)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s == 0)
		*T.ticket = 0; // the decode kernel of this run starts after this kernel (stream order)
	if (s >= T.n_streams)
		return;

	DevStream* d = T.streams + s;
	const uint8_t* src = d->src;
	const uint64_t size = d->src_size;
	const uint32_t vs = d->vertex_size;
	const uint32_t count = d->vertex_count;
	const uint32_t bv = block_vertices(vs);
	const uint32_t nblocks = (count + bv - 1) / bv;

	uint32_t* boff = T.block_offset + d->block_base + s;
	uint32_t* bstream = T.block_stream + d->block_base;
	uint16_t* coff = T.chan_offset + d->chan_base;

	int status = 0;
	uint32_t version = 0;
	uint32_t first_bad = 0; // blocks >= first_bad are not decodable

	// stream framing (reference src/vertexcodec.cpp:1827-1851)
	if (size < 1)
		status = -2;
	else
	{
		uint32_t h = __ldg(src);
		version = h & 0x0f;
		if ((h & 0xf0) != kMagic || version > 1)
			status = -1;
		else if (size - 1 < tail_padded(vs, version))
			status = -2;
	}

	if (status == 0 && version != 0 && nblocks > 0)
	{
		// a channel byte with mode 3 makes the first block fail (:1584-1585)
		const uint8_t* channels = src + size - vs / 4;
		for (uint32_t q = 0; q < vs / 4; ++q)
			if ((__ldg(channels + q) & 3u) == 3u)
				status = -2;
	}

	uint64_t p = 1;

	if (status == 0)
	{
		first_bad = nblocks;
		const uint32_t ctrl_bytes = version ? vs / 4 : 0;

		for (uint32_t b = 0; b < nblocks; ++b)
		{
			const uint32_t n = min(bv, count - b * bv);
			const uint32_t groups = (n + kGroup - 1) / kGroup;
			const uint32_t na = groups * kGroup;
			const uint32_t hdr = (groups + 3) / 4;
			const uint64_t start = p;
			bool ok = true;

			boff[b] = (uint32_t)start;
			bstream[b] = s;

			if (size - p < ctrl_bytes)
				ok = false;
			const uint8_t* control = src + p;
			p += ctrl_bytes;

			for (uint32_t k = 0; ok && k < vs; ++k)
			{
				coff[(uint64_t)b * vs + k] = (uint16_t)(p - start);
				uint32_t ctrl = version ? (__ldg(control + (k >> 2)) >> ((k & 3) * 2)) & 3u : 0u;

				if (ctrl == 3)
				{
					// literal bytes (:1546-1554): the 16-aligned count must be readable
					if (size - p < na)
						ok = false;
					p += n;
				}
				else if (ctrl != 2)
				{
					if (size - p < hdr)
					{
						ok = false;
						break;
					}
					uint32_t selectors = ldg_u32_at(src + p, hdr); // hdr <= 4 bytes, 2 bits per group
					p += hdr;

					for (uint32_t g = 0; g < groups; ++g)
					{
						if (size - p < kGroupReadLimit)
						{
							ok = false;
							break;
						}
						uint32_t bits = group_bits(version, ctrl, (selectors >> (g * 2)) & 3u);
						if (bits == 8)
							p += 16;
						else if (bits != 0)
						{
							uint32_t w0 = ldg_u32_at(src + p, 4);
							uint32_t w1 = bits == 4 ? ldg_u32_at(src + p + 4, 4) : 0u;
							p += 2 * bits + count_sentinels(bits, w0, w1);
						}
					}
				}
			}

			if (!ok)
			{
				first_bad = b;
				status = -2;
				break;
			}
		}

		if (status == 0)
		{
			boff[nblocks] = (uint32_t)p;
			if (size - p != tail_padded(vs, version))
				status = -3; // (:1868-1869) the blocks were decodable, the stream is still rejected
		}
	}

	for (uint32_t b = first_bad; b < nblocks; ++b)
	{
		boff[b] = kInvalidOffset;
		bstream[b] = s;
	}
	if (first_bad < nblocks || nblocks == 0)
		boff[nblocks] = kInvalidOffset;

	d->version = (uint8_t)version;
	d->status = status;
	T.status[s] = status;
}

// ------------------------------------------------------------------------------------------------
// phases 2+3: block decode
// ------------------------------------------------------------------------------------------------

struct BlockParams
{
	uint32_t ticket;
	uint32_t valid;
	uint32_t vs, n, groups, nq;
	uint32_t version, filter;
	uint32_t first_block; // block 0 of its stream: the carry is the tail's first vertex
	uint32_t cb_shift;    // position of the block's first byte inside the staging buffer
	uint32_t store_align; // 16, 4 or 1
	uint32_t pad;
	const uint8_t* tail;  // first vertex (vs bytes) then, for v1, vs/4 channel bytes
	uint8_t* out;
	const uint16_t* chan_offset;
	unsigned long long* lookback; // this block's entries (nq of them); predecessors lie nq entries lower each
};

// shared-memory map of one decode CTA (dynamic shared memory, 16-byte aligned pieces)
constexpr uint32_t kStageBytes = 12672; // >= kMaxEncodedBlock + 15 (alignment) + 16 (over-read slack), also holds the output tile
constexpr uint32_t kPlaneBytes = kBlockBytes;
constexpr uint32_t kSmemStage = 0;
constexpr uint32_t kSmemPlanes = kSmemStage + kStageBytes;
constexpr uint32_t kSmemGroupTab = kSmemPlanes + kPlaneBytes;   // u32[512]: offset | bits << 16
constexpr uint32_t kSmemTotals = kSmemGroupTab + 512 * 4;       // u32[128]: per (chunk, lane) scan totals
constexpr uint32_t kSmemCarry = kSmemTotals + 128 * 4;          // u32[64]: inclusive prefix of all previous blocks
constexpr uint32_t kSmemChanOff = kSmemCarry + 64 * 4;          // u16[256]
constexpr uint32_t kSmemChannels = kSmemChanOff + 256 * 2;      // u8[64] channel bytes + u8[64] control bytes
constexpr uint32_t kSmemParams = kSmemChannels + 128;           // BlockParams (<= 96 bytes)
constexpr uint32_t kSmemBarrier = kSmemParams + 96;             // mbarrier
constexpr uint32_t kSmemTotal = kSmemBarrier + 16;

static_assert(kStageBytes >= kMaxEncodedBlock + 31, "staging buffer too small");
static_assert(kStageBytes >= kBlockBytes + 512, "output tile (with per-chunk padding) must fit in the staging buffer");
static_assert(sizeof(BlockParams) <= 96, "BlockParams grew");

uint32_t decode_smem_bytes()
{
	return kSmemTotal;
}

// generic lane-wise "previous + delta" for the three channel modes with one code path:
//   H = 0x80808080 -> four byte lanes, 0x80008000 -> two 16-bit lanes, 0xffffffff -> xor
__device__ __forceinline__ uint32_t lane_combine(uint32_t a, uint32_t b, uint32_t H)
{
	return ((a & ~H) + (b & ~H)) ^ ((a ^ b) & H);
}

// output tile: row r (vertex) of vs bytes; every 16-row chunk is displaced by `pad` extra bytes
// (vs rounded up to 16) so that the 4-byte column writes of different chunks fall into different
// banks while 16-byte reads stay aligned
__device__ __forceinline__ uint32_t tile_pad(uint32_t vs)
{
	return (vs + 15u) & ~15u;
}

__device__ __forceinline__ uint32_t tile_offset(uint32_t r, uint32_t vs)
{
	return r * vs + (r >> 4) * tile_pad(vs);
}

// byte plane k, group g: 16-byte slots rotated by the channel quad so that the 128-bit reads of the
// transpose (same group, consecutive quads) hit different banks
__device__ __forceinline__ uint32_t plane_offset(uint32_t k, uint32_t g, uint32_t groups, uint32_t na)
{
	uint32_t slot = g + ((k >> 2) % groups);
	slot = slot >= groups ? slot - groups : slot;
	return k * na + slot * 16;
}

__global__ void __launch_bounds__(kDecodeThreads) decode_kernel(DevTables T)
{
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* stage = smem + kSmemStage;
	uint8_t* planes = smem + kSmemPlanes;
	uint32_t* group_tab = reinterpret_cast<uint32_t*>(smem + kSmemGroupTab);
	uint32_t* totals = reinterpret_cast<uint32_t*>(smem + kSmemTotals);
	uint32_t* carry = reinterpret_cast<uint32_t*>(smem + kSmemCarry);
	uint16_t* chan_off = reinterpret_cast<uint16_t*>(smem + kSmemChanOff);
	uint8_t* channels = smem + kSmemChannels;
	uint8_t* control = smem + kSmemChannels + 64;
	BlockParams& P = *reinterpret_cast<BlockParams*>(smem + kSmemParams);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmemBarrier);

	const uint32_t tid = threadIdx.x;
	uint32_t parity = 0;

	if (tid == 0)
	{
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	for (;;)
	{
		// ---- take a block ------------------------------------------------------------------------
		if (tid == 0)
		{
			uint32_t ticket = atomicAdd(T.ticket, 1u);
			P.ticket = ticket;
			P.valid = 0;
			if (ticket < T.total_blocks)
			{
				uint32_t s = T.block_stream[ticket];
				const DevStream* d = T.streams + s;
				uint32_t b = ticket - d->block_base;
				const uint32_t* boff = T.block_offset + d->block_base + s + b;
				uint32_t off = boff[0], end = boff[1];
				if (off != kInvalidOffset && end != kInvalidOffset)
				{
					uint32_t vs = d->vertex_size;
					uint32_t bv = block_vertices(vs);
					uint32_t n = min(bv, d->vertex_count - b * bv);
					uint32_t version = d->version;
					P.valid = 1;
					P.vs = vs;
					P.n = n;
					P.groups = (n + kGroup - 1) / kGroup;
					P.nq = vs / 4;
					P.version = version;
					P.filter = d->filter;
					P.first_block = b == 0;
					P.tail = d->src + d->src_size - tail_bytes(vs, version);
					uint8_t* out = d->dst + (uint64_t)b * bv * vs;
					P.out = out;
					uintptr_t oa = reinterpret_cast<uintptr_t>(out);
					P.store_align = (oa & 15) == 0 ? 16 : ((oa & 3) == 0 ? 4 : 1);
					P.chan_offset = T.chan_offset + d->chan_base + (uint64_t)b * vs;
					P.lookback = T.lookback + (d->chan_base >> 2) + (uint64_t)b * (vs / 4);

					// stage the encoded block: 16-byte aligned window around [off, end)
					uintptr_t a0 = reinterpret_cast<uintptr_t>(d->src) + off;
					uintptr_t a1 = reinterpret_cast<uintptr_t>(d->src) + end;
					uintptr_t lo = a0 & ~uintptr_t(15);
					uintptr_t hi = (a1 + 15) & ~uintptr_t(15);
					P.cb_shift = (uint32_t)(a0 - lo);
					uint32_t bytes = (uint32_t)(hi - lo);
					fence_proxy_async(); // earlier generic-proxy reads/writes of the staging buffer are ordered before the copy
					if (bytes > 0)
					{
						mbar_expect_tx(bar, bytes);
						tma_load_bulk(stage, reinterpret_cast<const void*>(lo), bytes, bar);
					}
					else
						mbar_expect_tx(bar, 0);
				}
			}
		}
		__syncthreads();

		if (P.ticket >= T.total_blocks)
			break;
		if (!P.valid)
		{
			__syncthreads(); // P is rewritten by thread 0 at the top of the loop
			continue;
		}

		const uint32_t vs = P.vs, n = P.n, groups = P.groups, nq = P.nq;
		const uint32_t na = groups * kGroup;
		const uint32_t version = P.version;
		const uint32_t hdr = (groups + 3) / 4;
		const uint32_t cb = P.cb_shift;

		// small per-block tables from global memory while the bulk copy is in flight
		for (uint32_t k = tid; k < vs; k += kDecodeThreads)
			chan_off[k] = P.chan_offset[k];
		if (tid < nq)
		{
			channels[tid] = version ? P.tail[vs + tid] : 0;
			if (P.first_block)
			{
				// carry into block 0 = first vertex stored in the tail (:1846-1849)
				const uint8_t* fv = P.tail + tid * 4;
				carry[tid] = (uint32_t)fv[0] | ((uint32_t)fv[1] << 8) | ((uint32_t)fv[2] << 16) | ((uint32_t)fv[3] << 24);
			}
		}

		while (!mbar_try_wait(bar, parity))
		{
		}
		parity ^= 1;

		if (tid < nq)
			control[tid] = version ? stage[cb + tid] : 0;
		__syncthreads();

		// ---- phase 2a: offsets of the groups inside every byte-channel (one lane per channel) -------
		for (uint32_t k = tid; k < vs; k += kDecodeThreads)
		{
			uint32_t ctrl = (control[k >> 2] >> ((k & 3) * 2)) & 3u;
			uint32_t o = cb + chan_off[k];
			uint32_t* tab = group_tab + k * groups;
			if (ctrl == 3)
			{
				for (uint32_t g = 0; g < groups; ++g)
					tab[g] = (o + g * 16) | (8u << 16);
			}
			else if (ctrl == 2)
			{
				for (uint32_t g = 0; g < groups; ++g)
					tab[g] = 0;
			}
			else
			{
				uint32_t selectors = lds_u32_at(stage, o);
				o += hdr;
				for (uint32_t g = 0; g < groups; ++g)
				{
					uint32_t bits = group_bits(version, ctrl, (selectors >> (g * 2)) & 3u);
					tab[g] = o | (bits << 16);
					if (bits == 8)
						o += 16;
					else if (bits != 0)
					{
						uint32_t w0 = lds_u32_at(stage, o);
						uint32_t w1 = bits == 4 ? lds_u32_at(stage, o + 4) : 0u;
						o += 2 * bits + count_sentinels(bits, w0, w1);
					}
				}
			}
		}
		__syncthreads();

		// ---- phase 2b: unpack, one thread per 16-value group ---------------------------------------------
		const uint32_t total_groups = vs * groups;
		for (uint32_t gi = tid; gi < total_groups; gi += kDecodeThreads)
		{
			uint32_t k = gi / groups;
			uint32_t g = gi - k * groups;
			uint32_t entry = group_tab[gi];
			uint32_t o = entry & 0xffffu;
			uint32_t bits = entry >> 16;
			uint32_t pofs = plane_offset(k, g, groups, na);
			uint4 r = make_uint4(0, 0, 0, 0);
			uint32_t m0 = 0, m1 = 0; // sentinel positions, most significant bit first
			uint32_t sh = 0;
			uint32_t esc = o;

			if (bits == 8)
			{
				const uint32_t* w = reinterpret_cast<const uint32_t*>(stage) + (o >> 2);
				uint32_t s8 = (o & 3u) * 8u;
				uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
				r.x = __funnelshift_r(a0, a1, s8);
				r.y = __funnelshift_r(a1, a2, s8);
				r.z = __funnelshift_r(a2, a3, s8);
				r.w = __funnelshift_r(a3, a4, s8);
			}
			else if (bits == 4)
			{
				uint32_t x0 = lds_u32_at(stage, o), x1 = lds_u32_at(stage, o + 4);
				uint32_t h0 = (x0 >> 4) & 0x0f0f0f0fu, l0 = x0 & 0x0f0f0f0fu;
				uint32_t h1 = (x1 >> 4) & 0x0f0f0f0fu, l1 = x1 & 0x0f0f0f0fu;
				r.x = __byte_perm(h0, l0, 0x5140);
				r.y = __byte_perm(h0, l0, 0x7362);
				r.z = __byte_perm(h1, l1, 0x5140);
				r.w = __byte_perm(h1, l1, 0x7362);
				uint32_t t0 = x0 & (x0 >> 1), t1 = x1 & (x1 >> 1);
				t0 &= t0 >> 2;
				t1 &= t1 >> 2;
				// byte-swap: value i of the word ends up at bit 28-4i, so clz enumerates values in order
				m0 = __byte_perm(t0 & 0x11111111u, 0, 0x0123);
				m1 = __byte_perm(t1 & 0x11111111u, 0, 0x0123);
				sh = 2;
				esc = o + 8;
			}
			else if (bits == 2)
			{
				uint32_t x = lds_u32_at(stage, o);
				uint32_t b0 = x & 0xff, b1 = (x >> 8) & 0xff, b2 = (x >> 16) & 0xff, b3 = x >> 24;
				r.x = ((b0 * 0x01004010u) & 0x03030300u) | (b0 >> 6);
				r.y = ((b1 * 0x01004010u) & 0x03030300u) | (b1 >> 6);
				r.z = ((b2 * 0x01004010u) & 0x03030300u) | (b2 >> 6);
				r.w = ((b3 * 0x01004010u) & 0x03030300u) | (b3 >> 6);
				m0 = __byte_perm(x & (x >> 1) & 0x55555555u, 0, 0x0123); // value i at bit 30-2i
				sh = 1;
				esc = o + 4;
			}
			else if (bits == 1)
			{
				uint32_t x = lds_u32_at(stage, o) & 0xffffu; // bit i = value i
				r.x = ((x & 15u) * 0x00204081u) & 0x01010101u;
				r.y = (((x >> 4) & 15u) * 0x00204081u) & 0x01010101u;
				r.z = (((x >> 8) & 15u) * 0x00204081u) & 0x01010101u;
				r.w = ((x >> 12) * 0x00204081u) & 0x01010101u;
				m0 = __brev(x); // value i at bit 31-i
				sh = 0;
				esc = o + 2;
			}

			*reinterpret_cast<uint4*>(planes + pofs) = r;

			// escape bytes replace the all-ones fields, in order
			uint32_t base = 0;
			for (uint32_t m = m0;;)
			{
				while (m)
				{
					uint32_t pz = __clz(m);
					m &= ~(0x80000000u >> pz);
					planes[pofs + base + (pz >> sh)] = stage[esc++];
				}
				if (base || m1 == 0)
					break;
				base = 8;
				m = m1;
			}
		}
		__syncthreads();

		// ---- phase 3a: transpose to vertex words, undo zigzag / rotation, scan 16 vertices ---------------------
		const uint32_t items = groups * nq;
		uint32_t w[16];
		uint32_t q = 0, c = 0;
		uint32_t H = 0x80808080u;
		const bool active = tid < items;
		if (active)
		{
			c = tid / nq;
			q = tid - c * nq;
			uint32_t channel = channels[q];
			uint32_t mode = channel & 3u;
			// per-lane constants of the generic transform r = ((t >> s1) & M) ^ ((t & L) * K), t = rotl(x, rot)
			uint32_t rot = mode == 2 ? (32u - (channel >> 4)) & 31u : 0u;
			uint32_t s1 = mode == 2 ? 0u : 1u;
			uint32_t M = mode == 0 ? 0x7f7f7f7fu : (mode == 1 ? 0x7fff7fffu : 0xffffffffu);
			uint32_t L = mode == 0 ? 0x01010101u : (mode == 1 ? 0x00010001u : 0u);
			uint32_t K = mode == 0 ? 0xffu : 0xffffu;
			H = mode == 0 ? 0x80808080u : (mode == 1 ? 0x80008000u : 0xffffffffu);

			uint4 pa = *reinterpret_cast<const uint4*>(planes + plane_offset(4 * q + 0, c, groups, na));
			uint4 pb = *reinterpret_cast<const uint4*>(planes + plane_offset(4 * q + 1, c, groups, na));
			uint4 pc = *reinterpret_cast<const uint4*>(planes + plane_offset(4 * q + 2, c, groups, na));
			uint4 pd = *reinterpret_cast<const uint4*>(planes + plane_offset(4 * q + 3, c, groups, na));
			const uint32_t A[4] = {pa.x, pa.y, pa.z, pa.w};
			const uint32_t B[4] = {pb.x, pb.y, pb.z, pb.w};
			const uint32_t C[4] = {pc.x, pc.y, pc.z, pc.w};
			const uint32_t D[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				uint32_t t0 = __byte_perm(A[j], B[j], 0x5140);
				uint32_t t1 = __byte_perm(A[j], B[j], 0x7362);
				uint32_t u0 = __byte_perm(C[j], D[j], 0x5140);
				uint32_t u1 = __byte_perm(C[j], D[j], 0x7362);
				w[4 * j + 0] = __byte_perm(t0, u0, 0x5410);
				w[4 * j + 1] = __byte_perm(t0, u0, 0x7632);
				w[4 * j + 2] = __byte_perm(t1, u1, 0x5410);
				w[4 * j + 3] = __byte_perm(t1, u1, 0x7632);
			}
#pragma unroll
			for (int i = 0; i < 16; ++i)
			{
				uint32_t t = __funnelshift_l(w[i], w[i], rot);
				w[i] = ((t >> s1) & M) ^ ((t & L) * K);
			}
#pragma unroll
			for (int i = 1; i < 16; ++i)
				w[i] = lane_combine(w[i - 1], w[i], H);
			totals[c * nq + q] = w[15];
		}
		__syncthreads();

		// ---- phase 3b: per 4-byte lane: in-block exclusive scan of the chunk totals + decoupled look-back ----
		if (tid < nq)
		{
			uint32_t channel = channels[tid];
			uint32_t mode = channel & 3u;
			uint32_t Hq = mode == 0 ? 0x80808080u : (mode == 1 ? 0x80008000u : 0xffffffffu);
			uint32_t run = 0;
			for (uint32_t cc = 0; cc < groups; ++cc)
			{
				uint32_t t = totals[cc * nq + tid];
				totals[cc * nq + tid] = run;
				run = lane_combine(run, t, Hq);
			}
			// run = aggregate of this block
			const unsigned long long tag = (unsigned long long)(T.epoch << 2) << 32;
			unsigned long long* mine = P.lookback + tid;
			uint32_t prefix;
			if (P.first_block)
				prefix = carry[tid];
			else
			{
				st_volatile_u64(mine, tag | (1ull << 32) | run); // state 1: aggregate only
				prefix = 0;
				const unsigned long long* prev = mine - nq;
				for (;;)
				{
					unsigned long long e = ld_volatile_u64(prev);
					uint32_t flag = (uint32_t)(e >> 32);
					if ((flag >> 2) != (T.epoch & 0x3fffffffu) || (flag & 3u) == 0)
						continue; // not published yet in this run
					prefix = lane_combine(prefix, (uint32_t)e, Hq);
					if ((flag & 3u) == 2)
						break;
					prev -= nq;
				}
				carry[tid] = prefix;
			}
			st_volatile_u64(mine, tag | (2ull << 32) | lane_combine(prefix, run, Hq)); // state 2: inclusive prefix
		}
		__syncthreads();

		// ---- phase 3c: add the carry, write the vertex tile (the staging buffer is free now) ----------------------
		uint8_t* tile = stage;
		if (active)
		{
			uint32_t startv = lane_combine(carry[q], totals[c * nq + q], H);
			const int filter = (int)P.filter;
			const bool word_filter = filter == MOB200_FILTER_EXP || ((filter == MOB200_FILTER_OCT || filter == MOB200_FILTER_COLOR) && vs == 4);
#pragma unroll
			for (int i = 0; i < 16; ++i)
			{
				uint32_t v = lane_combine(startv, w[i], H);
				if (word_filter)
					v = apply_filter32(v, filter);
				*reinterpret_cast<uint32_t*>(tile + tile_offset(c * 16 + i, vs) + q * 4) = v;
			}
		}
		__syncthreads();

		// ---- phase 3d: 8-byte filters on whole vertices ---------------------------------------------------------------
		if (P.filter != MOB200_FILTER_NONE && vs == 8 && P.filter != MOB200_FILTER_EXP)
		{
			for (uint32_t r = tid; r < n; r += kDecodeThreads)
			{
				uint2* e = reinterpret_cast<uint2*>(tile + tile_offset(r, vs));
				*e = apply_filter64(*e, (int)P.filter);
			}
			__syncthreads();
		}

		// ---- phase 3e: tile -> global memory -----------------------------------------------------------------------------
		{
			const uint32_t nbytes = n * vs;
			const uint32_t chunk_bytes = 16 * vs;
			const uint32_t pad = tile_pad(vs);
			uint8_t* out = P.out;
			if (P.store_align == 16)
			{
				const uint32_t pieces = nbytes >> 4;
				for (uint32_t j = tid; j < pieces; j += kDecodeThreads)
				{
					uint32_t o = j << 4;
					uint32_t ch = o / chunk_bytes;
					uint4 v = *reinterpret_cast<const uint4*>(tile + o + ch * pad);
					*reinterpret_cast<uint4*>(out + o) = v;
				}
				const uint32_t rem_words = (nbytes & 15u) >> 2;
				if (tid < rem_words)
				{
					uint32_t o = (pieces << 4) + tid * 4;
					uint32_t ch = o / chunk_bytes;
					*reinterpret_cast<uint32_t*>(out + o) = *reinterpret_cast<const uint32_t*>(tile + o + ch * pad);
				}
			}
			else if (P.store_align == 4)
			{
				for (uint32_t j = tid; j < (nbytes >> 2); j += kDecodeThreads)
				{
					uint32_t o = j << 2;
					uint32_t ch = o / chunk_bytes;
					*reinterpret_cast<uint32_t*>(out + o) = *reinterpret_cast<const uint32_t*>(tile + o + ch * pad);
				}
			}
			else
			{
				for (uint32_t o = tid; o < nbytes; o += kDecodeThreads)
				{
					uint32_t ch = o / chunk_bytes;
					out[o] = tile[o + ch * pad];
				}
			}
		}
		__syncthreads(); // the tile / tables are reused by the next block
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

cudaError_t launch_walk(const DevTables& T, cudaStream_t stream)
{
	if (T.n_streams == 0)
		return cudaSuccess;
	uint32_t grid = (T.n_streams + kWalkThreads - 1) / kWalkThreads;
	walk_kernel<<<grid, kWalkThreads, 0, stream>>>(T);
	return cudaGetLastError();
}

cudaError_t prepare_decode_kernel()
{
	return cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
}

cudaError_t decode_occupancy(int* ctas_per_sm)
{
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, decode_kernel, kDecodeThreads, kSmemTotal);
}

cudaError_t launch_decode(const DevTables& T, uint32_t grid, cudaStream_t stream)
{
	if (T.total_blocks == 0)
		return cudaSuccess;
	decode_kernel<<<grid, kDecodeThreads, kSmemTotal, stream>>>(T);
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
