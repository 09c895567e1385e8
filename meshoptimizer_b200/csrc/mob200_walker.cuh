// mob200_walker.cuh -- phase 1 of the decode path: the walker warp (one lane per stream).
//
// The stream stores no index: the offset of a block / byte-channel / group is the sum of the
// data-dependent sizes of everything before it (reference src/vertexcodec.cpp:1375-1425,1531-1568,
// 1857-1866 advance a single pointer).  That chain is serial inside a stream but independent across
// streams, so a walker warp advances 32 streams in lock step, one per lane:
//
//   * warp-synchronous, branch-free group steps, fully unrolled per byte-channel; everything that does not
//     depend on the running offset (field width, masks, shift amounts, fixed size) comes from a small
//     shared-memory table indexed by the group's 2-bit selector, so the dependent chain of a step is
//     "offset -> 3 shared-memory words -> funnel shift -> and/shift -> popc -> add" (~60 cycles);
//   * each lane streams its encoded bytes through a private 512-byte shared-memory ring of four 128-byte
//     chunks filled by per-lane TMA bulk copies (cp.async.bulk, completion on warp-shared mbarriers, one
//     per refill round); the ring is topped up every four groups, so the bytes a step needs were requested
//     at least one round earlier; after a jump (literal channels are skipped without being read) the
//     landing chunk is waited for; an L2 prefetch runs a few KB ahead of every lane;
//   * byte-channels that are zero or literal in every lane (warp vote) are skipped without touching data.
//
// Output per block: its byte range (block_offset), one table row of 16 group entries per byte-channel
// (group_table), a release store of the block's ready word that hands the block to the
// producers, and at the end the reference return code of the stream (:1827-1869).
#pragma once

#include "mob200_device.cuh"

namespace mob200
{

constexpr uint32_t kWalkRingBytes = 512;  // per lane
constexpr uint32_t kWalkChunkBytes = 128; // TMA granule
constexpr uint32_t kWalkChunks = kWalkRingBytes / kWalkChunkBytes;
constexpr uint32_t kWalkNeed = 208;       // bytes a lane may read between two refill points: header (4) + 8 groups (192) + over-read (12)
constexpr uint32_t kWalkPrefetch = 4096;  // L2 prefetch distance

// shared-memory layout of the walker warp: 32 rings, the step tables, the refill barrier
constexpr uint32_t kWalkSmemRings = 0;
constexpr uint32_t kWalkSmemLut = kWalkSmemRings + 32 * kWalkRingBytes; // 6 selector tables x 4 entries x 32 bytes (128-byte aligned)
constexpr uint32_t kWalkSmemBars = kWalkSmemLut + 6 * 4 * 32;
constexpr uint32_t kWalkSmemBytes = kWalkSmemBars + 16;

// step-table entry for one field width (index into {0,1,2,4,8} bits), two 16-byte halves:
//   a = { mask0, mask1, s1, s2 }: all-ones fields of data word j:  y = x & (x >> s1);  z = y & (y >> s2) & mask_j
//   b = { fixed bytes, table code, entry mask (0xffff for a group that stores bytes, else 0), 0 }
__device__ __forceinline__ void walk_lut_entry(uint32_t idx, uint4& a, uint4& b)
{
	switch (idx)
	{
	case 1: a = make_uint4(0x0000ffffu, 0u, 0u, 0u), b = make_uint4(2u, 0u, 0xffffu, 0u); break;
	case 2: a = make_uint4(0x55555555u, 0u, 1u, 0u), b = make_uint4(4u, 1u, 0xffffu, 0u); break;
	case 3: a = make_uint4(0x11111111u, 0x11111111u, 1u, 2u), b = make_uint4(8u, 2u, 0xffffu, 0u); break;
	case 4: a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(16u, 3u, 0xffffu, 0u); break;
	default: a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u); break;
	}
}

// selector tables: 0 inactive / zero / literal channel (no movement), 1 spare, 2 v0, 3 v1 ctrl 0, 4 v1 ctrl 1, 5 spare
__device__ __forceinline__ void walk_lut_init(uint8_t* lut, uint32_t lane)
{
	if (lane < 24)
	{
		const uint32_t table = lane >> 2, sel = lane & 3u;
		uint32_t idx = 0;
		switch (table)
		{
		case 2: idx = sel ? sel + 1 : 0; break;
		case 3: idx = sel; break;
		case 4: idx = sel + 1; break;
		default: idx = 0; break;
		}
		uint4 a, b;
		walk_lut_entry(idx, a, b);
		reinterpret_cast<uint4*>(lut)[2 * lane] = a;
		reinterpret_cast<uint4*>(lut)[2 * lane + 1] = b;
	}
	__syncwarp();
}

// per-lane walker state
struct WalkLane
{
	const uint8_t* src;
	const uint8_t* org; // src rounded down to 16 bytes: all positions below are relative to it
	uint32_t rel0;      // src & 15
	uint32_t rel;       // read position
	uint32_t rel_end;   // rel0 + size
	uint32_t limit;     // rel_end rounded up to 16: bytes at or beyond it are never fetched
	uint32_t vs, count, bv, nblocks, version;
	int status;
	uint32_t sbase;     // shared-space address of this lane's ring (512-byte aligned) XOR the lane's swizzle (lane % 8) << 4:
	                    // byte p of the ring lives at sbase ^ p.  The rings of a warp are 512 bytes apart, i.e. on the same
	                    // banks; the swizzle moves the 16-byte pieces that the lanes of a quarter-warp copy with one cp.async
	                    // instruction to eight different bank groups (4 shared-memory wavefronts per instruction instead of ~28)
	uint32_t issued;    // chunks [0, issued) have been requested or skipped
	uint32_t last;      // number of chunks of the stream
	uint32_t safe_end;  // bytes below it are in the ring (landed) as of the latest refill point
	uint32_t prefetched; // L2 prefetch position (multiple of 128)
};

// warp-uniform refill state: at most one round of copies is in flight
struct WalkWarp
{
	long long dbg_steps, dbg_refill; uint32_t dbg_general; // (temporary) cycles in the step loops / in refills, general channels
	long long dbg_wait, dbg_hard; // cycles spent waiting for the previous round / for a round that is needed at once
	uint32_t dbg_refills, dbg_hards;
	bool pending;     // a round is in flight
	uint32_t lut;     // shared-space address of the step tables
};

// every lane waits for its own copies (cp.async completion is tracked per thread, and a lane only reads its own ring)
__device__ __forceinline__ void walk_wait(WalkWarp& W)
{
	asm volatile("cp.async.wait_all;" ::: "memory");
	W.pending = false;
}

// Refill point: top up every lane's ring from the chunk that holds L.rel with 16-byte cp.async copies (all
// lanes at once: a per-lane TMA bulk copy would be issued lane by lane).  The round issued at the previous
// refill point is waited for first (it was issued a byte-channel ago and has normally landed), so copies
// never overtake each other into a ring slot; if a lane is about to read bytes that are only requested now
// (start of a stream, after a jump over literal channels) this round is waited for, too.
// `on` lanes take part; the others only keep the warp converged.
__device__ __forceinline__ void walk_refill(WalkLane& L, WalkWarp& W, bool on, uint32_t lane)
{
	const long long dbg_r0 = dbg_clock();
	W.dbg_refills++;
	if (W.pending)
	{
		const long long c0 = dbg_clock();
		walk_wait(W);
		W.dbg_wait += dbg_clock() - c0;
	}

	const uint32_t ci = L.rel / kWalkChunkBytes;
	if (L.issued < ci)
		L.issued = ci; // chunks that were jumped over are never read
	uint32_t safe = L.issued * kWalkChunkBytes; // everything requested so far has landed
	const uint32_t want = min(ci + kWalkChunks, L.last); // (topping up only three chunks ahead: walk alone 14 % slower, fused unchanged)
	const uint32_t fresh = on ? want - min(want, L.issued) : 0u; // 0..4 chunks to request
	const uint32_t fmax = __reduce_max_sync(0xffffffffu, fresh);

	if (fmax)
	{
#pragma unroll 1
		for (uint32_t k = 0; k < fmax; ++k)
		{
			if (k < fresh)
			{
				const uint32_t c = L.issued + k;
				const uint32_t b0 = c * kWalkChunkBytes;
				const uint32_t dst = (c & (kWalkChunks - 1)) * kWalkChunkBytes;
				const uint8_t* src = L.org + b0;
				if (L.limit - b0 >= kWalkChunkBytes)
				{
#pragma unroll
					for (uint32_t j = 0; j < kWalkChunkBytes; j += 16)
						asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(L.sbase ^ (dst + j)), "l"(src + j) : "memory");
				}
				else
				{
					// last chunk of the stream: only the 16-byte pieces inside [src & ~15, (src + size + 15) & ~15) are read
					for (uint32_t j = 0; b0 + j < L.limit; j += 16)
						asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(L.sbase ^ (dst + j)), "l"(src + j) : "memory");
				}
			}
		}
		L.issued += fresh;
		W.pending = true;
		const bool hard = on && min(L.rel + kWalkNeed, L.limit) > safe;
		if (__any_sync(0xffffffffu, hard))
		{
			const long long c0 = dbg_clock();
			walk_wait(W);
			W.dbg_hard += dbg_clock() - c0;
			W.dbg_hards++;
			safe = L.issued * kWalkChunkBytes;
		}
	}
	L.safe_end = safe;

	// L2 prefetch a few KB ahead (the decoders' copies and this ring's later refills then hit L2)
	if (L.prefetched < (L.rel & ~127u))
		L.prefetched = L.rel & ~127u;
	if (on && L.prefetched < L.rel + kWalkPrefetch && L.prefetched < L.limit)
	{
		asm volatile("prefetch.global.L2 [%0];" ::"l"(L.org + L.prefetched));
		L.prefetched += 128;
	}
	W.dbg_refill += dbg_clock() - dbg_r0;
}

__device__ __forceinline__ uint32_t ring_word(const WalkLane& L, uint32_t rel)
{
	uint32_t v;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(L.sbase ^ (rel & (kWalkRingBytes - 4))));
	return v;
}

__device__ __forceinline__ uint32_t ring_u32_at(const WalkLane& L, uint32_t rel)
{
	return __funnelshift_r(ring_word(L, rel), ring_word(L, rel + 4), (rel & 3u) * 8u);
}

// step-table entry of one group: the two halves of walk_lut_entry
struct WalkEntry
{
	uint4 a, b;
};

__device__ __forceinline__ WalkEntry walk_entry_load(uint32_t la)
{
	WalkEntry e;
	asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.a.x), "=r"(e.a.y), "=r"(e.a.z), "=r"(e.a.w) : "r"(la));
	asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.b.x), "=r"(e.b.y), "=r"(e.b.z), "=r"(e.b.w) : "r"(la + 16u));
	return e;
}

// One group step for every lane with the lane's step-table entry for this group (the all-zero entry for a
// lane that has no such group).  Returns the group-table entry.
__device__ __forceinline__ uint32_t walk_step(WalkLane& L, const WalkEntry& t, uint32_t start)
{
	const uint32_t entry = (((L.rel - start) << 2) + t.b.y) & t.b.z;

	const uint32_t w0 = ring_word(L, L.rel), w1 = ring_word(L, L.rel + 4), w2 = ring_word(L, L.rel + 8);
	const uint32_t sh = L.rel << 3;
	const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
	const uint32_t y0 = x0 & (x0 >> t.a.z), y1 = x1 & (x1 >> t.a.z);
	const uint32_t z0 = y0 & (y0 >> t.a.w) & t.a.x, z1 = y1 & (y1 >> t.a.w) & t.a.y;
	L.rel += t.b.x + __popc(z0 + 2u * z1); // fixed part {0,2,4,8,16} + escape bytes (the two masks never share a bit after the shift)
	return entry;
}

// Walk block `b` of every lane's stream (lanes with on == false only keep the warp converged).
// Returns false for a lane whose block is malformed.
__device__ __forceinline__ bool walk_block(WalkLane& L, WalkWarp& W, bool on, uint32_t n, uint16_t* rows, uint32_t vs_max, uint32_t lane)
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
	walk_refill(L, W, on && !bad, lane);
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
				const uint32_t e0 = lit ? (((L.rel - start) << 2) | 3u) : 0u;
				const uint32_t pair = lit ? e0 | ((e0 + (16u << 2)) << 16) : 0u; // entries of groups 0 and 1
				const uint32_t inc = lit ? (32u << 2) * 0x00010001u : 0u;        // two groups further (no carry between the halves: entries < 2^16)
				uint32_t w[8];
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j)
					w[j] = pair + inc * j;
				if (groups < 16)
				{
#pragma unroll
					for (uint32_t j = 0; j < 8; ++j)
						w[j] &= (2 * j < groups ? 0xffffu : 0u) | (2 * j + 1 < groups ? 0xffff0000u : 0u);
				}
				row[0] = make_uint4(w[0], w[1], w[2], w[3]);
				if (groups > 8)
					row[1] = make_uint4(w[4], w[5], w[6], w[7]);
				L.rel += lit ? n : 0u;
			}
			continue;
		}

		// general path: at least one lane has a bit-packed channel here; literal and zero lanes run the
		// same steps with the all-zero table entry (their position does not move)
		if (packed && L.rel_end - L.rel < hdr) // (:1376)
			bad = true;
		if (lit && L.rel_end - L.rel < na) // (:1546-1554)
			bad = true;
		const bool gon0 = kon && !bad;
		const bool pk = gon0 && packed;

		// The 16 group steps in four rounds of four (a rolled loop: the three warp roles share the instruction
		// cache), branch-free and independent of each other except through L.rel.  The reference requires 24
		// readable bytes in front of every group of a bit-packed channel (:1385,:1415); positions only grow, so the
		// rule holds for all groups iff it holds for the last one (a lane that runs past its input keeps stepping
		// over stale ring bytes -- memory-safe -- and is rejected below).
		uint32_t sels = 0, lutbase = W.lut, p_last = 0;
		WalkEntry t;
		const long long dbg_s0 = dbg_clock();
		W.dbg_general++;
#pragma unroll 1
		for (uint32_t gq = 0; gq < 16; gq += 4)
		{
			// refill at the start of the channel (literal skips may have moved rel arbitrarily), and in the middle if a
			// lane has come within reach of the end of what its ring is known to hold
			if (gq == 0 || (gq == 8 && __any_sync(0xffffffffu, gon0 && min(L.rel + kWalkNeed, L.limit) > L.safe_end)))
				walk_refill(L, W, gon0, lane);
			if (gq == 0)
			{
				sels = pk ? ring_u32_at(L, L.rel) : 0u;
				lutbase = W.lut + (pk ? (version ? 3u + ctrl : 2u) : 0u) * 128u;
				L.rel += pk ? hdr : 0u;
				p_last = L.rel;
				t = walk_entry_load(0 < groups ? lutbase | ((sels << 5) & 0x60u) : W.lut);
			}
			// (the shared-memory loads are volatile and stay in program order: the table entry of the next group is
			// requested before the ring words of the current one, so its latency is off the chain of the running offset)
			uint32_t e[4];
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j)
			{
				const uint32_t nsel = j < 3 ? ((sels >> (2 * (j + 1))) & 3u) : ((sels >> 8) & 3u);
				const WalkEntry nxt = walk_entry_load(gq + j + 1 < groups ? lutbase + nsel * 32u : W.lut);
				p_last = gq + j < groups ? L.rel : p_last;
				e[j] = walk_step(L, t, start);
				t = nxt;
			}
			sels >>= 8;
			if (gon0 && !lit) // (zero channels store zeros)
				*reinterpret_cast<uint2*>(rows + (size_t)k * 16 + gq) = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
		}
		W.dbg_steps += dbg_clock() - dbg_s0;
		if (pk && L.rel_end - min(p_last, L.rel_end) < kGroupReadLimit)
			bad = true;
		if (lit && !bad)
		{
			// literal channel: group g is the 16 raw bytes at rel + 16 g; the channel holds n bytes, not 16 * groups (:1553)
			const uint32_t e0 = ((L.rel - start) << 2) | 3u;
			uint32_t r[8];
#pragma unroll
			for (uint32_t j = 0; j < 8; ++j)
			{
				const uint32_t lo = 2 * j < groups ? e0 + (16u << 2) * (2 * j) : 0u;
				const uint32_t hi = 2 * j + 1 < groups ? e0 + (16u << 2) * (2 * j + 1) : 0u;
				r[j] = lo | (hi << 16);
			}
			row[0] = make_uint4(r[0], r[1], r[2], r[3]);
			row[1] = make_uint4(r[4], r[5], r[6], r[7]);
			L.rel += n;
		}
	}
	return !bad;
}

// Framing checks of one stream (reference src/vertexcodec.cpp:1827-1851, and the channel rule :1584-1585):
// returns the reference code (0, -1, -2) and the codec version.
__device__ __forceinline__ int walk_framing(const uint8_t* src, uint32_t size, uint32_t vs, uint32_t nblocks, uint32_t& version)
{
	int status = 0;
	version = 0;
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
	if (status != 0)
		version = 0;
	return status;
}

// One pass of the walker warp over 32 lane jobs.
//
//   serial walk (DevTables::block_mode == 0): lane <-> stream base + lane, walked from byte 1 to its tail; every
//     block's end offset goes to the block-offset table.
//   block mode: lane <-> decode ticket base + lane = ONE block whose start comes from the block-offset table (a
//     sidecar).  The lane walks that block -- group-table rows as usual -- and hands it to the producers only if the
//     walk ends exactly where the table says the next block starts (the padded tail for the last block) and block 0
//     starts at byte 1: the union of all verified blocks is then the chain the serial walk would have followed.
//     Anything else marks the stream kStatusSidecar.
template <bool kBlock>
__device__ void walk_group(const DevTables& T, WalkWarp& W, uint32_t base, uint32_t lane, uint32_t ring_smem)
{
	constexpr bool bm = kBlock; // (a template parameter: each kernel carries the code of one mode only -- the roles share the SM's instruction cache)
	const uint32_t job = base + lane;
	const bool have = job < (bm ? T.total_blocks : T.n_streams);
	uint32_t s = have ? job : 0, b_first = 0;
	if (bm)
	{
		const uint2 info = have ? __ldg(T.ticket_info + job) : make_uint2(0u, 0u);
		s = info.x, b_first = info.y;
	}
	const DevStream* d = T.streams + s;

	WalkLane L;
	L.src = d->src;
	const uint32_t size = have ? d->src_size : 0;
	L.vs = d->vertex_size;
	L.count = d->vertex_count;
	L.bv = d->block_groups * kGroup;
	L.nblocks = have ? d->nblocks : 0;
	L.status = walk_framing(L.src, size, L.vs, L.nblocks, L.version);

	uint32_t* boff = T.block_offset + d->block_base + s;
	uint32_t* ready = T.block_ready + d->block_base;

	uint32_t off = 1, want_end = 0;
	if (bm && have && L.status == 0)
	{
		off = __ldcg(boff + b_first);
		want_end = __ldcg(boff + b_first + 1);
		const uint32_t body_end = size - tail_padded(L.vs, L.version); // (:1868) the blocks end where the padded tail starts
		if (off < 1 || off > body_end || want_end < off || want_end > body_end || (b_first == 0 && off != 1) || (b_first + 1 == L.nblocks && want_end != body_end))
			L.status = kStatusSidecar;
	}

	L.rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(L.src) & 15u);
	L.org = L.src - L.rel0;
	L.rel = L.rel0 + off;
	L.rel_end = L.rel0 + size;
	L.limit = (L.rel_end + 15u) & ~15u;
	L.sbase = ring_smem | ((lane & 7u) << 4);
	L.issued = 0;
	L.safe_end = 0;
	L.prefetched = 0;

	const bool framed = have && L.status == 0;
	const uint32_t b_count = bm ? 1u : L.nblocks; // blocks this lane walks
	uint32_t done = 0;                            // blocks [b_first, b_first + done) are decodable
	bool alive = framed;
	if (!framed)
		L.limit = 0; // nothing is fetched for a job that is not walked
	L.last = (L.limit + kWalkChunkBytes - 1) / kWalkChunkBytes;
	if (!bm && framed && L.nblocks)
		boff[0] = 1;

	const uint32_t max_blocks = __reduce_max_sync(0xffffffffu, alive ? b_count : 0u);
	for (uint32_t i = 0; i < max_blocks; ++i)
	{
		const uint32_t b = b_first + i;
		const bool on = alive && i < b_count;
		const uint32_t n = on ? min(L.bv, L.count - b * L.bv) : 16u;
		const uint32_t vs_max = __reduce_max_sync(0xffffffffu, on ? L.vs : 0u);
		uint16_t* rows = T.group_table + (d->chan_base + (uint64_t)b * L.vs) * 16;
		const bool ok = walk_block(L, W, on, n, rows, vs_max, lane);
		if (on)
		{
			if (ok && (!bm || L.rel - L.rel0 == want_end))
			{
				if (!bm)
					boff[b + 1] = L.rel - L.rel0;
				done = i + 1;
				st_release_u32(ready + b, ready_word(T.epoch, L.version, true));
			}
			else
			{
				L.status = bm ? kStatusSidecar : -2;
				alive = false;
			}
		}
	}
	// every copy into the rings has landed before the lanes move on to their next jobs
	if (W.pending)
		walk_wait(W);

	if (have)
	{
		if (!bm && L.status == 0 && L.rel_end - L.rel != tail_padded(L.vs, L.version))
			L.status = -3; // (:1868-1869) the blocks were decodable, the stream is still rejected
		if (done < b_count)
		{
			// block `done` failed (or the framing did): its end offset and everything after it is invalid
			if (!bm)
				for (uint32_t b = framed ? done + 1 : 0; b <= L.nblocks; ++b)
					boff[b] = kInvalidOffset;
			for (uint32_t b = b_first + done; b < b_first + b_count; ++b)
				st_release_u32(ready + b, ready_word(T.epoch, 0, false)); // nothing more will come: the producers skip the rest
		}
		if (!bm)
			T.status[d->caller_index] = L.status;
		else if (L.status != 0)
		{
			// (block mode: the status words are cleared before the launch -- or, after the team walk, already hold
			// the reference codes, which a block that was left undecodable must not replace)
			if (T.keep_status)
				atomicCAS(reinterpret_cast<int*>(T.status + d->caller_index), 0, L.status);
			else
				T.status[d->caller_index] = L.status;
		}
	}
}

// Block mode: streams without blocks (vertex_count 0) only have their framing checked; they follow the streams
// with blocks in the sorted stream array.
__device__ void walk_frame_group(const DevTables& T, uint32_t base, uint32_t lane)
{
	const uint32_t s = base + lane;
	if (s >= T.n_streams)
		return;
	const DevStream* d = T.streams + s;
	uint32_t version;
	int status = walk_framing(d->src, d->src_size, d->vertex_size, 0, version);
	if (status == 0 && d->src_size - 1 != tail_padded(d->vertex_size, version))
		status = -3; // (:1868-1869)
	if (status != 0 && !T.keep_status)
		T.status[d->caller_index] = status;
}

template <bool kBlock>
__device__ void walker_main(const DevTables& T, uint8_t* region)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ring_smem = smem_addr(region + kWalkSmemRings) + lane * kWalkRingBytes;
	if (ring_smem & (kWalkRingBytes - 1))
		__trap(); // the ring addressing needs 512-byte aligned rings
	WalkWarp W;
	W.lut = smem_addr(region + kWalkSmemLut);
	W.pending = false;
	W.dbg_wait = W.dbg_hard = 0;
	W.dbg_steps = W.dbg_refill = 0;
	W.dbg_general = 0;
	W.dbg_refills = W.dbg_hards = 0;
	const long long dbg_t0 = dbg_clock();
	walk_lut_init(region + kWalkSmemLut, lane);
	__syncwarp();

	// block mode: tickets [0, total_blocks) in groups of 32, then the streams that have no blocks
	const uint32_t ticket_span = (T.total_blocks + 31u) & ~31u;
	for (;;)
	{
		uint32_t base = 0;
		if (lane == 0)
			base = atomicAdd(T.counters + 1, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (kBlock && base >= ticket_span)
		{
			if (T.n_with_blocks + (base - ticket_span) >= T.n_streams)
				break;
			walk_frame_group(T, T.n_with_blocks + (base - ticket_span), lane);
		}
		else
		{
			if (!kBlock && base >= T.n_streams)
				break;
			walk_group<kBlock>(T, W, base, lane, ring_smem);
		}
		__syncwarp();
	}
#ifdef MOB200_DEBUG_COUNTERS
	if (lane == 0)
	{
		unsigned long long* dbg = reinterpret_cast<unsigned long long*>(T.counters + 16);
		atomicAdd(dbg + 8, (unsigned long long)(dbg_clock() - dbg_t0));
		atomicAdd(dbg + 9, (unsigned long long)W.dbg_wait);
		atomicAdd(dbg + 10, (unsigned long long)W.dbg_hard);
		atomicAdd(dbg + 11, (unsigned long long)W.dbg_refills);
		atomicAdd(dbg + 12, (unsigned long long)W.dbg_hards);
		atomicAdd(dbg + 13, (unsigned long long)W.dbg_steps);
		atomicAdd(dbg + 14, (unsigned long long)W.dbg_refill);
		atomicAdd(dbg + 15, (unsigned long long)W.dbg_general);
	}
#endif
}

} // namespace mob200
