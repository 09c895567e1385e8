// mob200_decoder.cuh -- phases 2+3 of the decode path: the decoder warps (one block per CTA iteration).
#pragma once

#include "mob200_device.cuh"
#include "mob200_filters.cuh"

namespace mob200
{

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
	uint32_t m_groups;    // ceil(2^32 / groups)          (x / groups    = umulhi(x, m_groups))
	uint32_t m_nq;        // ceil(2^32 / nq)
	uint32_t m_chunk;     // ceil(2^32 / (16 * vs))
	const uint8_t* tail;  // first vertex (vs bytes) then, for v1, vs/4 channel bytes
	uint8_t* out;
	const uint16_t* rows; // group table rows of this block
	unsigned long long* lookback; // this block's entries (nq of them); predecessors lie nq entries lower each
};

// shared-memory map of one CTA (dynamic shared memory, 16-byte aligned pieces)
constexpr uint32_t kStageBytes = 12672; // >= kMaxEncodedBlock + 15 (alignment) + 16 (over-read slack), also holds the output tile
constexpr uint32_t kPlaneBytes = kBlockBytes;
constexpr uint32_t kSmemStage = 0;
constexpr uint32_t kSmemPlanes = kSmemStage + kStageBytes;
constexpr uint32_t kSmemGroupTab = kSmemPlanes + kPlaneBytes;   // u16[vs][16] <= 8 KB only for vs = 256; see below
constexpr uint32_t kGroupTabBytes = 1024;                      // rows are compacted to `groups` entries: vs*groups*2 <= 1024
constexpr uint32_t kSmemTotals = kSmemGroupTab + kGroupTabBytes; // u32[128]: per (chunk, lane) scan totals
constexpr uint32_t kSmemCarry = kSmemTotals + 128 * 4;          // u32[64]: inclusive prefix of all previous blocks
constexpr uint32_t kSmemChannels = kSmemCarry + 64 * 4;         // u8[64] channel bytes
constexpr uint32_t kSmemParams = kSmemChannels + 64;            // BlockParams (<= 112 bytes)
constexpr uint32_t kSmemBarrier = kSmemParams + 112;            // mbarrier
constexpr uint32_t kSmemRing = (kSmemBarrier + 16 + 127) & ~127u; // walker warp: 32 lanes x 128-byte ring
constexpr uint32_t kSmemRows = kSmemRing + 32 * 128;               // walker warp: 32 lanes x one 32-byte table row
constexpr uint32_t kSmemTotal = kSmemRows + 32 * 32;

static_assert(kStageBytes >= kMaxEncodedBlock + 31, "staging buffer too small");
static_assert(kStageBytes >= kBlockBytes + 512, "output tile (with per-chunk padding) must fit in the staging buffer");
static_assert(sizeof(BlockParams) <= 112, "BlockParams grew");

uint32_t decode_smem_bytes()
{
	return kSmemTotal;
}

__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic, uint32_t d)
{
	return d == 1 ? x : __umulhi(x, magic);
}

__device__ __forceinline__ uint32_t magic_for(uint32_t d)
{
	return d <= 1 ? 0u : (uint32_t)(0xffffffffu / d) + 1u; // ceil(2^32 / d) for d >= 2 (exact for x*d < 2^32)
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
	uint32_t slot = g + ((k >> 2) & 15u);
	slot = groups == 16 ? (slot & 15u) : (slot % groups);
	return k * na + slot * 16;
}

__device__ void decoder_main(const DevTables& T, uint8_t* smem)
{
	uint8_t* stage = smem + kSmemStage;
	uint8_t* planes = smem + kSmemPlanes;
	uint16_t* group_tab = reinterpret_cast<uint16_t*>(smem + kSmemGroupTab);
	uint32_t* totals = reinterpret_cast<uint32_t*>(smem + kSmemTotals);
	uint32_t* carry = reinterpret_cast<uint32_t*>(smem + kSmemCarry);
	uint8_t* channels = smem + kSmemChannels;
	BlockParams& P = *reinterpret_cast<BlockParams*>(smem + kSmemParams);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSmemBarrier);

	const uint32_t tid = threadIdx.x;
	uint32_t parity = 0;

	if (tid == 0)
	{
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	decoder_sync();

	for (;;)
	{
		// ---- take a block ------------------------------------------------------------------------
		if (tid == 0)
		{
			uint32_t ticket = atomicAdd(T.counters, 1u);
			P.ticket = ticket;
			P.valid = 0;
			if (ticket < T.total_blocks)
			{
				uint2 info = __ldg(T.ticket_info + ticket);
				const uint32_t s = info.x, b = info.y;
				const DevStream* d = T.streams + s;

				// wait until the walker has published this block
				const unsigned long long* progress = T.progress + s;
				// (exponential back-off: a starved decoder must not take issue slots from the walker warps)
				for (uint32_t ns = 64;;)
				{
					unsigned long long v = ld_acquire_u64(progress);
					if ((uint32_t)(v >> 32) == T.epoch && (uint32_t)v > b)
						break;
					__nanosleep(ns);
					ns = ns < 4096 ? ns * 2 : ns;
				}

				const uint32_t* boff = T.block_offset + d->block_base + s + b;
				uint32_t off = __ldcg(boff), end = __ldcg(boff + 1);
				if (off != kInvalidOffset && end != kInvalidOffset)
				{
					uint32_t vs = d->vertex_size;
					uint32_t bv = block_vertices(vs);
					uint32_t n = min(bv, d->vertex_count - b * bv);
					uint32_t version = __ldg(d->src) & 0x0fu;
					uint32_t groups = (n + kGroup - 1) / kGroup;
					P.valid = 1;
					P.vs = vs;
					P.n = n;
					P.groups = groups;
					P.nq = vs / 4;
					P.version = version;
					P.filter = d->filter;
					P.first_block = b == 0;
					P.m_groups = magic_for(groups);
					P.m_nq = magic_for(vs / 4);
					P.m_chunk = magic_for(16 * vs);
					P.tail = d->src + d->src_size - tail_bytes(vs, version);
					uint8_t* out = d->dst + (uint64_t)b * bv * vs;
					P.out = out;
					uintptr_t oa = reinterpret_cast<uintptr_t>(out);
					P.store_align = (oa & 15) == 0 ? 16 : ((oa & 3) == 0 ? 4 : 1);
					P.rows = T.group_table + (d->chan_base + (uint64_t)b * vs) * 16;
					P.lookback = T.lookback + (d->chan_base >> 2) + (uint64_t)b * (vs / 4);

					// stage the encoded block: 16-byte aligned window around [off, end)
					uintptr_t a0 = reinterpret_cast<uintptr_t>(d->src) + off;
					uintptr_t a1 = reinterpret_cast<uintptr_t>(d->src) + end;
					uintptr_t lo = a0 & ~uintptr_t(15);
					uintptr_t hi = (a1 + 15) & ~uintptr_t(15);
					P.cb_shift = (uint32_t)(a0 - lo);
					uint32_t bytes = (uint32_t)(hi - lo);
					fence_proxy_async(); // earlier generic-proxy accesses of the staging buffer are ordered before the copy
					mbar_expect_tx(bar, bytes);
					if (bytes > 0)
						tma_load_bulk(stage, reinterpret_cast<const void*>(lo), bytes, bar);
				}
			}
		}
		decoder_sync();

		if (P.ticket >= T.total_blocks)
			break;
		if (!P.valid)
		{
			decoder_sync(); // P is rewritten by thread 0 at the top of the loop
			continue;
		}

		const uint32_t vs = P.vs, n = P.n, groups = P.groups, nq = P.nq;
		const uint32_t na = groups * kGroup;
		const uint32_t version = P.version;
		const uint32_t cb = P.cb_shift;

		// group table rows (written by a walker on another SM: read through L2) -> compact rows of
		// `groups` entries in shared memory, while the bulk copy is in flight
		{
			const uint2* rows = reinterpret_cast<const uint2*>(P.rows);
			const uint32_t quads = (groups + 3) >> 2; // 4 entries per 8-byte load
			for (uint32_t i = tid; i < vs * 4; i += kDecodeThreads)
			{
				uint32_t k = i >> 2, part = i & 3;
				if (part < quads)
				{
					uint2 v = __ldcg(rows + i);
					uint16_t* dstp = group_tab + k * groups + part * 4;
					if ((groups & 3u) == 0)
						*reinterpret_cast<uint2*>(dstp) = v;
					else
					{
						uint32_t left = groups - part * 4;
						dstp[0] = (uint16_t)v.x;
						if (left > 1)
							dstp[1] = (uint16_t)(v.x >> 16);
						if (left > 2)
							dstp[2] = (uint16_t)v.y;
						if (left > 3)
							dstp[3] = (uint16_t)(v.y >> 16);
					}
				}
			}
		}
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

		if (tid == 0)
		{
			while (!mbar_try_wait(bar, parity)) // one thread polls; the others sleep in the barrier below
			{
			}
		}
		parity ^= 1;
		decoder_sync();

		// ---- phase 2: unpack, one thread per 16-value group ------------------------------------------------
		const uint32_t total_groups = vs * groups;
		for (uint32_t gi = tid; gi < total_groups; gi += kDecodeThreads)
		{
			uint32_t k = groups == 16 ? gi >> 4 : fast_div(gi, P.m_groups, groups);
			uint32_t g = gi - k * groups;
			uint32_t entry = group_tab[gi];
			uint32_t o = cb + (entry >> 2);
			uint32_t bits = entry ? (1u << (entry & 3u)) : 0u;
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
		decoder_sync();

		// ---- phase 3a: transpose to vertex words, undo zigzag / rotation, scan 16 vertices ---------------------
		const uint32_t items = groups * nq;
		uint32_t w[16];
		uint32_t q = 0, c = 0;
		uint32_t H = 0x80808080u;
		const bool active = tid < items;
		if (active)
		{
			c = fast_div(tid, P.m_nq, nq);
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
		decoder_sync();

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
		decoder_sync();

		// ---- phase 3c: add the carry, write the vertex tile (the staging buffer is free now) ----------------------
		uint8_t* tile = stage;
		if (active)
		{
			uint32_t startv = lane_combine(carry[q], totals[c * nq + q], H);
			const int filter = (int)P.filter;
			const bool word_filter = filter == MOB200_FILTER_EXP || ((filter == MOB200_FILTER_OCT || filter == MOB200_FILTER_COLOR) && vs == 4);
			uint8_t* col = tile + tile_offset(c * 16, vs) + q * 4;
#pragma unroll
			for (int i = 0; i < 16; ++i)
			{
				uint32_t v = lane_combine(startv, w[i], H);
				if (word_filter)
					v = apply_filter32(v, filter);
				*reinterpret_cast<uint32_t*>(col + i * vs) = v;
			}
		}
		decoder_sync();

		// ---- phase 3d: 8-byte filters on whole vertices ---------------------------------------------------------------
		if (P.filter != MOB200_FILTER_NONE && vs == 8 && P.filter != MOB200_FILTER_EXP)
		{
			for (uint32_t r = tid; r < n; r += kDecodeThreads)
			{
				uint2* e = reinterpret_cast<uint2*>(tile + tile_offset(r, vs));
				*e = apply_filter64(*e, (int)P.filter);
			}
			decoder_sync();
		}

		// ---- phase 3e: tile -> global memory -----------------------------------------------------------------------------
		{
			const uint32_t nbytes = n * vs;
			const uint32_t chunk_bytes = 16 * vs;
			const uint32_t pad = tile_pad(vs);
			const uint32_t m_chunk = P.m_chunk;
			uint8_t* out = P.out;
			if (P.store_align == 16)
			{
				const uint32_t pieces = nbytes >> 4;
				for (uint32_t j = tid; j < pieces; j += kDecodeThreads)
				{
					uint32_t o = j << 4;
					uint32_t ch = __umulhi(o, m_chunk);
					uint4 v = *reinterpret_cast<const uint4*>(tile + o + ch * pad);
					*reinterpret_cast<uint4*>(out + o) = v;
				}
				const uint32_t rem_words = (nbytes & 15u) >> 2;
				if (tid < rem_words)
				{
					uint32_t o = (pieces << 4) + tid * 4;
					uint32_t ch = __umulhi(o, m_chunk);
					*reinterpret_cast<uint32_t*>(out + o) = *reinterpret_cast<const uint32_t*>(tile + o + ch * pad);
				}
			}
			else if (P.store_align == 4)
			{
				for (uint32_t j = tid; j < (nbytes >> 2); j += kDecodeThreads)
				{
					uint32_t o = j << 2;
					uint32_t ch = __umulhi(o, m_chunk);
					*reinterpret_cast<uint32_t*>(out + o) = *reinterpret_cast<const uint32_t*>(tile + o + ch * pad);
				}
			}
			else
			{
				for (uint32_t o = tid; o < nbytes; o += kDecodeThreads)
				{
					uint32_t ch = __umulhi(o, m_chunk);
					out[o] = tile[o + ch * pad];
				}
			}
		}
		decoder_sync(); // the tile / tables are reused by the next block
	}
}


} // namespace mob200
