// mob200_walker.cuh -- phase 1 of the decode path: the walker warp (one lane per stream).
//
// The stream stores no index: the offset of a block / byte-channel / group is the sum of the
// data-dependent sizes of everything before it (reference src/vertexcodec.cpp:1375-1425,1531-1568,
// 1857-1866 advance a single pointer).  That chain is serial inside a stream but independent across
// streams, so a walker warp advances 32 streams in lock step, one per lane:
//
//   * warp-synchronous, branch-free inner loops with uniform trip counts (a lane that has nothing to
//     do is predicated off), so that the lanes never diverge: a single diverged lane costs as much as 32;
//   * each lane streams its encoded bytes through a private 128-byte ring in shared memory fed by
//     cp.async (32-byte chunks, requested 3-4 chunks ahead of the read position), so the dependent chain
//     "read packed bits -> count all-ones fields -> next offset" waits for shared memory only;
//   * byte-channels that are zero or literal in every lane (warp vote) are skipped without touching data.
//
// Output per block: its byte range (block_offset), one table row of 16 group entries per byte-channel
// (group_table), a release store of the per-stream progress counter that hands the block to the
// decoders, and at the end the reference return code of the stream (:1827-1869).
#pragma once

#include "mob200_device.cuh"

namespace mob200
{

constexpr uint32_t kRingBytes = 128;  // per lane
constexpr uint32_t kChunkBytes = 32;  // cp.async granule: 2 x 16 bytes
constexpr uint32_t kRingChunks = kRingBytes / kChunkBytes;

// per-lane walker state
struct WalkLane
{
	// stream
	const uint8_t* src;
	uint32_t rel0;      // src & 31: all positions below are relative to src rounded down to 32 bytes
	uint32_t rel;       // read position
	uint32_t rel_end;   // rel0 + size
	uint32_t vs, count, bv, nblocks, version;
	int status;
	// ring
	uint32_t sbase;     // shared-space address of this lane's ring
	const uint8_t* org; // src - rel0
	uint32_t rel_first; // rel0 & ~15: 16-byte pieces before it are outside the readable range
	uint32_t rel_limit; // 16-byte pieces at or beyond it are never fetched
	uint32_t next_rel;  // next chunk to request (multiple of kChunkBytes); everything below it has been requested
	const uint8_t* next_ptr; // org + next_rel
};

// predicated cp.async / commit without branches (a branch would let the lanes of the walker warp diverge)
__device__ __forceinline__ void cp_async16_if(uint32_t dst, const void* src, bool p)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"((uint32_t)p) : "memory");
}

__device__ __forceinline__ void cp_async_commit_if(bool p)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q cp.async.commit_group;\n\t}" ::"r"((uint32_t)p) : "memory");
}

// request the chunk at next_rel (both 16-byte pieces, each only if it lies inside
// [src & ~15, (src + size + 15) & ~15)) and move next_rel on -- all predicated on p
template <bool kFirst>
__device__ __forceinline__ void ring_issue_next(WalkLane& L, bool p)
{
	const uint32_t dst = L.sbase + (L.next_rel & (kRingBytes - 1));
	bool p0 = p && L.next_rel < L.rel_limit;
	bool p1 = p && L.next_rel + 16 < L.rel_limit;
	if (kFirst) // only the first chunk of a stream can start before the readable range
	{
		p0 = p0 && L.next_rel >= L.rel_first;
		p1 = p1 && L.next_rel + 16 >= L.rel_first;
	}
	cp_async16_if(dst, L.next_ptr, p0);
	cp_async16_if(dst + 16, L.next_ptr + 16, p1);
	cp_async_commit_if(p);
	L.next_rel += p ? kChunkBytes : 0u;
	L.next_ptr += p ? kChunkBytes : 0u;
}

// Called before every read at L.rel: requests (at most) one more chunk and waits until the chunks that
// cover [rel, rel + 32 + 11] have landed.  Correct as long as rel advanced by less than one chunk since
// the previous call (groups advance by <= 24 bytes); larger moves go through ring_jump.  At most two
// requests are in flight and they are issued in order, so a request never targets a ring slot that an
// older in-flight request is still writing.
__device__ __forceinline__ void ring_step(WalkLane& L, bool p)
{
	const bool q = p && L.next_rel < (L.rel & ~(kChunkBytes - 1)) + kRingBytes;
	ring_issue_next<false>(L, q);
	asm volatile("cp.async.wait_group 2;" ::: "memory");
}

// Reposition the ring after a move of arbitrary size: drain, then request the four chunks at the new position.
__device__ __forceinline__ void ring_jump(WalkLane& L, bool p)
{
	const uint32_t cur = L.rel & ~(kChunkBytes - 1);
	p = p && L.next_rel < cur + kRingBytes;
	if (__any_sync(0xffffffffu, p))
	{
		asm volatile("cp.async.wait_all;" ::: "memory");
		if (p && L.next_rel < cur)
		{
			L.next_ptr += cur - L.next_rel;
			L.next_rel = cur;
		}
#pragma unroll
		for (uint32_t i = 0; i < kRingChunks; ++i)
			ring_issue_next<true>(L, p && L.next_rel < cur + kRingBytes);
		asm volatile("cp.async.wait_group 2;" ::: "memory");
	}
}

__device__ __forceinline__ uint32_t ring_word(const WalkLane& L, uint32_t rel)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(L.sbase + (rel & (kRingBytes - 4))));
	return v;
}

__device__ __forceinline__ uint32_t ring_u32_at(const WalkLane& L, uint32_t rel)
{
	return __funnelshift_r(ring_word(L, rel), ring_word(L, rel + 4), (rel & 3u) * 8u);
}

// The 16 group steps of one byte-channel, identical instruction stream for every lane.  selectors holds the
// 2-bit group selectors (literal lanes: all ones, zero lanes: all zero); `bias` is added to a selector to
// index the widths {0,1,2,4,8}.  kChecked re-checks the reference's 24-byte rule before each group
// (:1385,:1415) and is only used close to the end of a stream.  Returns false for a lane that ran out of input.
template <bool kChecked>
__device__ __forceinline__ bool walk_groups(WalkLane& L, bool gon0, bool packed, uint32_t selectors, uint32_t bias, uint32_t groups, uint32_t start, uint32_t rowbuf)
{
	const uint32_t version = L.version;
	bool bad = false;
#pragma unroll 4
	for (uint32_t g = 0; g < 16; ++g, selectors >>= 2)
	{
		bool gon = gon0 && g < groups;
		if (kChecked)
		{
			if (gon && packed && L.rel_end - L.rel < kGroupReadLimit)
				bad = true;
			gon = gon && !bad;
		}
		const uint32_t sel = selectors & 3u;
		uint32_t idx = sel + (version ? bias : (uint32_t)(sel != 0u)); // index into the widths {0,1,2,4,8}
		idx = gon ? idx : 0u;
		const uint32_t entry = idx ? (((L.rel - start) << 2) | (idx - 1u)) : 0u;
		asm volatile("st.shared.u16 [%0], %1;" ::"r"(rowbuf + g * 2), "r"(entry) : "memory");
		ring_step(L, idx != 0u);
		const uint32_t x0 = ring_word(L, L.rel), x1 = ring_word(L, L.rel + 4), x2 = ring_word(L, L.rel + 8);
		const uint32_t sh = (L.rel & 3u) * 8u;
		const uint32_t w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh);
		// all-ones fields: 1-bit -> the 16 bits themselves, 2-bit -> both bits of a pair, 4-bit -> all four of
		// a nibble; selected with arithmetic masks (no branches: the lanes must stay converged)
		const uint32_t a = w0 & (w0 >> 1), c = w1 & (w1 >> 1);
		const uint32_t f1 = 0u - (uint32_t)(idx == 1u), f2 = 0u - (uint32_t)(idx == 2u), f3 = 0u - (uint32_t)(idx == 3u);
		const uint32_t m = (w0 & 0xffffu & f1) | (a & 0x55555555u & f2) | (a & (a >> 2) & 0x11111111u & f3);
		const uint32_t m2 = c & (c >> 2) & 0x11111111u & f3;
		L.rel += ((1u << idx) & ~1u) + __popc(m) + __popc(m2); // fixed part {0,2,4,8,16} + escape bytes
	}
	return !bad;
}

// Walk block `b` of every lane's stream (lanes with on == false only keep the warp converged).
// Returns false for a lane whose block is malformed.
__device__ __forceinline__ bool walk_block(WalkLane& L, bool on, uint32_t n, uint16_t* rows, uint32_t rowbuf, uint32_t vs_max)
{
	const uint32_t groups = (n + kGroup - 1) / kGroup;
	const uint32_t na = groups * kGroup;
	const uint32_t hdr = (groups + 3) / 4;
	const uint32_t start = L.rel;
	const uint32_t version = L.version;
	const uint32_t ctrl_bytes = version ? L.vs / 4 : 0;
	bool bad = false;

	if (on && L.rel_end - L.rel < ctrl_bytes)
		bad = true;

	// control bytes of the first 64 byte-channels are kept in registers (the ring moves on);
	// wider vertices read the rest from global memory
	ring_jump(L, on && !bad);
	uint32_t cw0 = ring_u32_at(L, L.rel), cw1 = ring_u32_at(L, L.rel + 4), cw2 = ring_u32_at(L, L.rel + 8), cw3 = ring_u32_at(L, L.rel + 12);
	const uint8_t* control = L.src + (L.rel - L.rel0);
	L.rel += (on && !bad) ? ctrl_bytes : 0u;

	for (uint32_t k = 0; k < vs_max; ++k)
	{
		const bool kon = on && !bad && k < L.vs;
		uint32_t cbyte = 0;
		if (k < 64)
		{
			uint32_t wsel = k >> 4;
			uint32_t word = wsel == 0 ? cw0 : (wsel == 1 ? cw1 : (wsel == 2 ? cw2 : cw3));
			cbyte = (word >> (((k >> 2) & 3u) * 8)) & 0xffu;
		}
		else if (kon && version)
			cbyte = __ldg(control + (k >> 2));
		const uint32_t ctrl = version ? (cbyte >> ((k & 3) * 2)) & 3u : 0u;
		const bool packed = kon && ctrl < 2;
		const bool lit = kon && ctrl == 3;
		uint4* row = reinterpret_cast<uint4*>(rows + (size_t)k * 16);

		if (!__any_sync(0xffffffffu, packed))
		{
			// zero or literal in every lane: no data is looked at
			if (lit && L.rel_end - L.rel < na) // (:1546-1554): the 16-aligned count must be readable
				bad = true;
			if (kon && !bad)
			{
				// literal: group g is the 16 raw bytes at rel + 16 g; zero: all entries 0
				uint32_t e0 = lit ? (((L.rel - start) << 2) | 3u) : 0u;
				uint32_t step = lit ? (16u << 2) : 0u;
				uint32_t w[8];
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j)
				{
					uint32_t lo = 2 * j < groups ? e0 + step * (2 * j) : 0u;
					uint32_t hi = 2 * j + 1 < groups ? e0 + step * (2 * j + 1) : 0u;
					w[j] = lo | (hi << 16);
				}
				row[0] = make_uint4(w[0], w[1], w[2], w[3]);
				if (groups > 8)
					row[1] = make_uint4(w[4], w[5], w[6], w[7]);
				L.rel += lit ? n : 0u;
			}
			continue;
		}

		// general path: at least one lane has a bit-packed channel here; literal and zero lanes run the
		// same loop as "all groups 8-bit raw" / "all groups zero" without a header
		if (packed && L.rel_end - L.rel < hdr) // (:1376)
			bad = true;
		if (lit && L.rel_end - L.rel < na) // (:1546-1554)
			bad = true;
		const bool gon0 = kon && !bad;
		ring_jump(L, gon0); // literal skips of the fast path above may have moved rel arbitrarily
		uint32_t selectors = packed ? ring_u32_at(L, L.rel) : (lit ? 0xffffffffu : 0u);
		const uint32_t bias = version ? (lit ? 1u : (packed ? ctrl : 0u)) : 0u;
		L.rel += (gon0 && packed) ? hdr : 0u;
		const uint32_t chan_data = L.rel;

		// the per-group bounds checks are only needed within reach of the end of the input
		const bool near_end = gon0 && packed && L.rel_end - L.rel < kGroupReadLimit * groups;
		bool ok;
		if (__any_sync(0xffffffffu, near_end))
			ok = walk_groups<true>(L, gon0, packed, selectors, bias, groups, start, rowbuf);
		else
			ok = walk_groups<false>(L, gon0, packed, selectors, bias, groups, start, rowbuf);
		bad = bad || !ok;
		if (lit && !bad)
			L.rel = chan_data + n; // a literal channel holds n bytes, not 16 * groups (:1553)

		if (kon && !bad)
		{
			uint4 lo, hi;
			asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(rowbuf) : "memory");
			row[0] = lo;
			if (groups > 8)
			{
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(rowbuf + 16) : "memory");
				row[1] = hi;
			}
		}
	}
	return !bad;
}

// One pass of the walker warp over 32 streams (lane <-> stream base + lane).
__device__ void walk_stream_group(const DevTables& T, uint32_t base, uint32_t lane, uint32_t ring_smem, uint32_t rowbuf)
{
	const uint32_t s = base + lane;
	const bool have = s < T.n_streams;
	const DevStream* d = T.streams + (have ? s : 0);

	WalkLane L;
	L.src = d->src;
	const uint32_t size = have ? d->src_size : 0;
	L.vs = d->vertex_size;
	L.count = d->vertex_count;
	L.bv = block_vertices(L.vs);
	L.nblocks = have ? d->nblocks : 0;
	L.version = 0;
	L.status = 0;

	uint32_t* boff = T.block_offset + d->block_base + s;
	unsigned long long* progress = T.progress + s;
	const unsigned long long tag = (unsigned long long)T.epoch << 32;

	// stream framing (reference src/vertexcodec.cpp:1827-1851)
	if (size < 1)
		L.status = -2;
	else
	{
		uint32_t h = __ldg(L.src);
		L.version = h & 0x0f;
		if ((h & 0xf0) != kMagic || L.version > 1)
			L.status = -1;
		else if (size - 1 < tail_padded(L.vs, L.version))
			L.status = -2;
	}
	if (L.status == 0 && L.version != 0 && L.nblocks > 0)
	{
		// a channel byte with mode 3 makes the first block fail (:1584-1585)
		const uint8_t* channels = L.src + size - L.vs / 4;
		for (uint32_t q = 0; q < L.vs / 4; ++q)
			if ((__ldg(channels + q) & 3u) == 3u)
				L.status = -2;
	}
	if (L.status != 0)
		L.version = 0;

	L.rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(L.src) & (kChunkBytes - 1));
	L.rel = L.rel0 + 1;
	L.rel_end = L.rel0 + size;
	L.sbase = ring_smem;
	L.org = L.src - L.rel0;
	L.rel_first = L.rel0 & ~15u;
	L.rel_limit = (L.rel_end + 15u) & ~15u;
	L.next_rel = 0;
	L.next_ptr = L.org;

	const bool framed = have && L.status == 0;
	uint32_t done = 0; // blocks [0, done) are decodable
	bool alive = framed;
	if (framed && L.nblocks)
		boff[0] = 1;

	const uint32_t max_blocks = __reduce_max_sync(0xffffffffu, alive ? L.nblocks : 0u);
	for (uint32_t b = 0; b < max_blocks; ++b)
	{
		const bool on = alive && b < L.nblocks;
		const uint32_t n = on ? min(L.bv, L.count - b * L.bv) : 16u;
		const uint32_t vs_max = __reduce_max_sync(0xffffffffu, on ? L.vs : 0u);
		uint16_t* rows = T.group_table + (d->chan_base + (uint64_t)b * L.vs) * 16;
		const bool ok = walk_block(L, on, n, rows, rowbuf, vs_max);
		if (on)
		{
			if (ok)
			{
				boff[b + 1] = L.rel - L.rel0;
				done = b + 1;
				st_release_u64(progress, tag | done);
			}
			else
			{
				L.status = -2;
				alive = false;
			}
		}
	}
	asm volatile("cp.async.wait_all;" ::: "memory"); // the ring is reused by this lane's next stream

	if (have)
	{
		if (L.status == 0 && L.rel_end - L.rel != tail_padded(L.vs, L.version))
			L.status = -3; // (:1868-1869) the blocks were decodable, the stream is still rejected
		if (done < L.nblocks)
		{
			// block `done` failed (or the framing did): its end offset and everything after it is invalid
			for (uint32_t b = framed ? done + 1 : 0; b <= L.nblocks; ++b)
				boff[b] = kInvalidOffset;
			st_release_u64(progress, tag | L.nblocks); // nothing more will come: the decoders skip the rest
		}
		T.status[d->caller_index] = L.status;
	}
}

__device__ void walker_main(const DevTables& T, uint8_t* ring_base, uint8_t* row_base)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ring_smem = smem_addr(ring_base) + lane * kRingBytes;
	const uint32_t rowbuf = smem_addr(row_base) + lane * 32u;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(T.counters + 1, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= T.n_streams)
			break;
		walk_stream_group(T, base, lane, ring_smem, rowbuf);
		__syncwarp();
	}
}

} // namespace mob200
