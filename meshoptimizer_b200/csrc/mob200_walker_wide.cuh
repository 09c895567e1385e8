// mob200_walker_wide.cuh -- phase 1, second form: one WARP per stream (used when a batch has too few
// streams to fill the lanes of the lane-per-stream walker, down to a single monolithic stream).
//
// The chain "offset of group g+1 = offset of group g + 2b + (all-ones fields of group g)" stays serial,
// but everything that does not depend on the running offset is done by the 32 lanes in parallel:
//
//   1. the warp streams the encoded bytes through a 2 KB shared-memory ring (coalesced cp.async, 512-byte
//      chunks, two chunks ahead);
//   2. per bit-packed byte-channel every lane takes 16 bytes of the 512-byte window that starts at the
//      channel and computes, for each of its byte positions p and each field width w in {1,2,4} bits, the
//      number of all-ones fields in the 2w bytes starting at p (SWAR per-byte counts, then sliding-window
//      sums by doubling) -- i.e. the answer to "how many escape bytes would a w-bit group starting at p
//      have" for EVERY p at once -- and stores 2w + that count, the bytes such a group takes, to one of three
//      512-entry byte tables in shared memory (two constant tables give 0 for a zero group and 16 for an
//      8-bit group);
//   3. the chain then costs one shared-memory byte load and one add per group (uniform across the warp);
//      the 24-byte rule is tested once per channel, outside the chain.
//
// Outputs are identical to the lane-per-stream walker (mob200_walker.cuh): block byte ranges, one table row
// of 16 group entries per byte-channel, a release store of the block's ready word, and the
// reference return code (src/vertexcodec.cpp:1827-1869).
#pragma once

#include "mob200_device.cuh"

namespace mob200
{

constexpr uint32_t kWideChunk = 512;                 // bytes per cp.async chunk (32 lanes x 16 bytes)
constexpr uint32_t kWideRingChunks = 4;
constexpr uint32_t kWideRingBytes = kWideChunk * kWideRingChunks; // 2 KB
constexpr uint32_t kWideTableBytes = 5 * 512;        // per width index {0,1,2,4,8 bits}: bytes a group starting at window position p takes (0 and 16 are constants)
constexpr uint32_t kWideSmemBytes = kWideRingBytes + kWideTableBytes;

struct WideRing
{
	uint32_t sbase;      // shared-space address of the ring
	const uint8_t* org;  // src rounded down to 16 bytes
	uint32_t rel_limit;  // 16-byte pieces at or beyond this relative offset are never fetched
	uint32_t issued;     // chunks [.., issued) have been requested
};

// make the bytes [rel & ~15, (rel & ~15) + 512 + 16) readable from the ring (rel relative to org)
__device__ __forceinline__ void wide_ring_ensure(WideRing& r, uint32_t rel, uint32_t lane)
{
	const uint32_t need = rel / kWideChunk;
	if (r.issued < need + 3)
	{
		if (r.issued + kWideRingChunks <= need) // a jump past everything in flight: restart at the new position
		{
			asm volatile("cp.async.wait_all;" ::: "memory");
			r.issued = need;
		}
		while (r.issued < need + 3)
		{
			const uint32_t piece = r.issued * kWideChunk + lane * 16;
			if (piece < r.rel_limit)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(r.sbase + (piece & (kWideRingBytes - 1))), "l"(r.org + piece) : "memory");
			asm volatile("cp.async.commit_group;" ::: "memory");
			++r.issued;
			// chunk c+4 reuses the slot of chunk c: never more than two requests in flight
			asm volatile("cp.async.wait_group 1;" ::: "memory");
		}
	}
	// chunks need and need+1 must have landed (issued >= need+3, at most one request pending: need+2)
	asm volatile("cp.async.wait_group 1;" ::: "memory");
	__syncwarp();
}

__device__ __forceinline__ uint32_t wide_ring_u32(const WideRing& r, uint32_t rel)
{
	uint32_t lo, hi;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(r.sbase + (rel & (kWideRingBytes - 4))));
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(r.sbase + ((rel + 4) & (kWideRingBytes - 4))));
	return __funnelshift_r(lo, hi, (rel & 3u) * 8u);
}

// per-byte counts of all-ones fields inside each byte of w: 4-bit fields (0..2), 2-bit (0..4), 1-bit (0..8)
__device__ __forceinline__ uint32_t bytes_n4(uint32_t w)
{
	uint32_t t = w & (w >> 1);
	t &= t >> 2;
	t &= 0x11111111u;
	return (t + (t >> 4)) & 0x0f0f0f0fu;
}

__device__ __forceinline__ uint32_t bytes_n2(uint32_t w)
{
	uint32_t t = w & (w >> 1) & 0x55555555u;
	t = (t + (t >> 2)) & 0x33333333u;
	return (t + (t >> 4)) & 0x0f0f0f0fu;
}

__device__ __forceinline__ uint32_t bytes_n1(uint32_t w)
{
	uint32_t t = w - ((w >> 1) & 0x55555555u);
	t = (t & 0x33333333u) + ((t >> 2) & 0x33333333u);
	return (t + (t >> 4)) & 0x0f0f0f0fu;
}

// n[0..5] hold per-byte counts of 24 consecutive bytes; returns in out[0..3] the sums over windows of
// `span` bytes (2, 4 or 8) starting at each of the first 16 byte positions (byte lanes never overflow: <= 16)
template <int kSpan>
__device__ __forceinline__ void window_sums(const uint32_t n[6], uint32_t out[4])
{
	uint32_t s2[6];
#pragma unroll
	for (int j = 0; j < 6; ++j)
		s2[j] = n[j] + __funnelshift_r(n[j], j < 5 ? n[j + 1] : 0u, 8);
	if (kSpan == 2)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
			out[j] = s2[j];
		return;
	}
	uint32_t s4[5];
#pragma unroll
	for (int j = 0; j < 5; ++j)
		s4[j] = s2[j] + __funnelshift_r(s2[j], s2[j + 1], 16);
	if (kSpan == 4)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
			out[j] = s4[j];
		return;
	}
#pragma unroll
	for (int j = 0; j < 4; ++j)
		out[j] = s4[j] + s4[j + 1];
}

// Walk one stream with the whole warp.  All lanes carry the same scalar state; lane l owns bytes
// [16 l, 16 l + 16) of the current window and, at the end of a channel, entry l of its table row.
__device__ void walk_stream_wide(const DevTables& T, uint32_t s, uint32_t lane, uint32_t smem_base)
{
	const DevStream* d = T.streams + s;
	const uint8_t* src = d->src;
	const uint32_t size = d->src_size;
	const uint32_t vs = d->vertex_size;
	const uint32_t count = d->vertex_count;
	const uint32_t bv = d->block_groups * kGroup;
	const uint32_t nblocks = d->nblocks;

	uint32_t* boff = T.block_offset + d->block_base + s;
	uint32_t* ready = T.block_ready + d->block_base;
	const uint32_t tab_base = smem_base + kWideRingBytes;

	int status = 0;
	uint32_t version = 0;

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

	const uint32_t rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
	const uint32_t rel_end = rel0 + size;
	uint32_t rel = rel0 + 1;
	uint32_t done = 0;
	uint32_t published_blocks = 0;
	const bool framed = status == 0;

	if (framed && nblocks)
	{
		WideRing ring;
		ring.sbase = smem_base;
		ring.org = src - rel0;
		ring.rel_limit = (rel_end + 15u) & ~15u;
		ring.issued = 0;

		if (lane == 0)
			boff[0] = 1;
		const uint32_t publish_every = vs <= 8 ? 4u : (vs <= 16 ? 2u : 1u);
		uint32_t published = 0; // blocks [0, published) have been handed to the producers
		for (uint32_t b = 0; b < nblocks && status == 0; ++b)
		{
			const uint32_t n = min(bv, count - b * bv);
			const uint32_t groups = (n + kGroup - 1) / kGroup;
			const uint32_t na = groups * kGroup;
			const uint32_t hdr = (groups + 3) / 4;
			const uint32_t start = rel;
			const uint32_t ctrl_bytes = version ? vs / 4 : 0;
			uint16_t* rows = T.group_table + (d->chan_base + (uint64_t)b * vs) * 16;
			bool bad = rel_end - rel < ctrl_bytes;

			// control bytes of the first 64 byte-channels are kept in registers; wider vertices read the
			// rest from global memory
			const uint8_t* control = src + (rel - rel0);
			uint32_t cw0 = 0, cw1 = 0, cw2 = 0, cw3 = 0;
			if (!bad && version)
			{
				wide_ring_ensure(ring, rel, lane);
				cw0 = wide_ring_u32(ring, rel);
				cw1 = wide_ring_u32(ring, rel + 4);
				cw2 = wide_ring_u32(ring, rel + 8);
				cw3 = wide_ring_u32(ring, rel + 12);
			}
			if (!bad)
				rel += ctrl_bytes;

			for (uint32_t k = 0; k < vs && !bad; ++k)
			{
				uint32_t cbyte;
				if (k < 64)
				{
					uint32_t wsel = k >> 4;
					uint32_t word = wsel == 0 ? cw0 : (wsel == 1 ? cw1 : (wsel == 2 ? cw2 : cw3));
					cbyte = (word >> (((k >> 2) & 3u) * 8)) & 0xffu;
				}
				else
					cbyte = version ? __ldg(control + (k >> 2)) : 0u;
				const uint32_t ctrl = (cbyte >> ((k & 3) * 2)) & 3u;
				uint32_t my_entry = 0; // entry of group `lane` (lanes >= 16 unused)

				if (ctrl == 3)
				{
					// literal bytes (:1546-1554): the 16-aligned count must be readable
					if (rel_end - rel < na)
					{
						bad = true;
						break;
					}
					my_entry = lane < groups ? ((((rel - start) + lane * 16) << 2) | 3u) : 0u;
					rel += n;
				}
				else if (ctrl != 2)
				{
					if (rel_end - rel < hdr) // (:1376)
					{
						bad = true;
						break;
					}
					wide_ring_ensure(ring, rel, lane);
					const uint32_t selectors = wide_ring_u32(ring, rel);
					rel += hdr;

					// which widths occur (uniform): bit i set <=> index i of {0,1,2,4,8} is used
					uint32_t used = 0;
					{
						const uint32_t gm = (groups >= 16 ? 0xffffffffu : ((1u << (2 * groups)) - 1u)) & 0x55555555u;
						const uint32_t lo = selectors & gm, hi = (selectors >> 1) & gm;
						const uint32_t b1 = version ? ctrl : 1u; // selector s > 0 maps to index s + b1 (v0) / s + ctrl (v1)
						if (~lo & ~hi & gm)
							used |= 1u << (version ? ctrl : 0u);
						if (lo & ~hi)
							used |= 1u << (1 + b1);
						if (~lo & hi)
							used |= 1u << (2 + b1);
						if (lo & hi)
							used |= 1u << (3 + b1);
					}

					// window = 512 bytes from the 16-byte boundary at or below the first group
					const uint32_t win = rel & ~15u;
					if (used & 0xeu)
					{
						// (the call above made the two chunks from rel's own readable: enough for the 528-byte window unless
						// the header ended within 48 bytes of a chunk boundary)
						if ((rel & (kWideChunk - 1)) > kWideChunk - 48 || (rel & (kWideChunk - 1)) < hdr)
							wide_ring_ensure(ring, rel, lane);
						uint32_t w[6];
						{
							uint32_t a = ring.sbase + ((win + lane * 16) & (kWideRingBytes - 1));
							asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
							uint32_t a2 = ring.sbase + ((win + lane * 16 + 16) & (kWideRingBytes - 1));
							asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[4]), "=r"(w[5]) : "r"(a2));
						}
						uint32_t n[6], out[4];
						if (used & 0x8u) // 4-bit fields: windows of 8 bytes
						{
#pragma unroll
							for (int j = 0; j < 6; ++j)
								n[j] = bytes_n4(w[j]);
							window_sums<8>(n, out);
							asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tab_base + 1536 + lane * 16), "r"(out[0] + 0x08080808u), "r"(out[1] + 0x08080808u), "r"(out[2] + 0x08080808u), "r"(out[3] + 0x08080808u) : "memory");
						}
						if (used & 0x4u) // 2-bit fields: windows of 4 bytes
						{
#pragma unroll
							for (int j = 0; j < 6; ++j)
								n[j] = bytes_n2(w[j]);
							window_sums<4>(n, out);
							asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tab_base + 1024 + lane * 16), "r"(out[0] + 0x04040404u), "r"(out[1] + 0x04040404u), "r"(out[2] + 0x04040404u), "r"(out[3] + 0x04040404u) : "memory");
						}
						if (used & 0x2u) // 1-bit fields: windows of 2 bytes
						{
#pragma unroll
							for (int j = 0; j < 6; ++j)
								n[j] = bytes_n1(w[j]);
							window_sums<2>(n, out);
							asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tab_base + 512 + lane * 16), "r"(out[0] + 0x02020202u), "r"(out[1] + 0x02020202u), "r"(out[2] + 0x02020202u), "r"(out[3] + 0x02020202u) : "memory");
						}
						__syncwarp();
					}

					// the chain: one table lookup per group.  The reference's 24-byte rule (:1385,:1415) is tested once, on the
					// position of the last group: positions only grow, so it fails there if it fails anywhere, and the test
					// (a branch on the running offset) stays out of the dependent chain  offset -> table byte -> offset.
					// (Table indices stay below 16 + 16 * 24 whatever the bytes are; ring reads wrap inside the ring.)
					// (Software-pipelined by hand: the table of group g+1 is worked out between the load of group g's step and
					// its first use, so the chain is  add -> LDS -> add  and nothing else waits behind the load.)
					uint32_t sel_bits = selectors;
					uint32_t rel_last = rel;
					uint32_t idx = (sel_bits & 3u) + (version ? ctrl : (uint32_t)((sel_bits & 3u) != 0u));
					uint32_t slot_base = tab_base + idx * 512u - win; // table of this width: bytes a group at position p takes (fixed part + escape bytes)
#pragma unroll 4
					for (uint32_t g = 0; g < groups; ++g)
					{
						uint32_t step;
						asm volatile("ld.shared.u8 %0, [%1];" : "=r"(step) : "r"(slot_base + rel));
						rel_last = rel;
						if (lane == g)
							my_entry = idx ? (((rel - start) << 2) | (idx - 1u)) : 0u;
						sel_bits >>= 2;
						idx = (sel_bits & 3u) + (version ? ctrl : (uint32_t)((sel_bits & 3u) != 0u));
						slot_base = tab_base + idx * 512u - win;
						asm volatile("" : "+r"(slot_base)); // (keeps idx * 512 out of the sum with the running offset)
						rel += step;
					}
					if (rel_end - rel_last < kGroupReadLimit || rel_last > rel_end)
					{
						bad = true;
						break;
					}
				}

				if (lane < 16 && (lane < 8 || groups > 8))
					rows[(size_t)k * 16 + lane] = (uint16_t)my_entry;
			}

			if (bad)
			{
				status = -2;
				break;
			}
			done = b + 1;
			__syncwarp();
			if (lane == 0)
				boff[b + 1] = rel - rel0;
			// blocks of small vertices are published in groups (a release is a MEMBAR.GPU: several hundred cycles next to
			// the ~850 per byte-channel of a block that has only 4 ... 16 of them)
			if (lane == 0 && ((done & (publish_every - 1u)) == 0 || done == nblocks))
			{
				// (the table rows were written by other lanes before the __syncwarp above: the
				// fence is cumulative over what this lane has synchronised with, and every strong store that follows it in
				// program order is a release: one MEMBAR.GPU per group of blocks -- a second one per block was a fifth of
				// the walk of a 4-byte-vertex stream)
				const uint32_t word = ready_word(T.epoch, version, true);
				asm volatile("fence.acq_rel.gpu;" ::: "memory");
				for (uint32_t p = published; p < done; ++p)
					st_volatile_u32(ready + p, word);
				published = done;
				published_blocks = done;
			}
		}
		asm volatile("cp.async.wait_all;" ::: "memory");
		__syncwarp();
	}

	if (framed && status == 0 && rel_end - rel != tail_padded(vs, version))
		status = -3; // (:1868-1869) the blocks were decodable, the stream is still rejected

	if (lane == 0)
	{
		if (done < nblocks)
		{
			for (uint32_t b = framed ? done + 1 : 0; b <= nblocks; ++b)
				boff[b] = kInvalidOffset;
			// blocks walked before the failure may still be waiting for their release; the rest is not decodable
			__threadfence();
			for (uint32_t b = 0; b < nblocks; ++b)
				if (b >= done || b >= published_blocks)
					st_volatile_u32(ready + b, ready_word(T.epoch, version, b < done));
		}
		T.status[d->caller_index] = status;
	}
}

__device__ void walker_main_wide(const DevTables& T, uint8_t* smem_region)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t base = smem_addr(smem_region);
	// the two constant step tables: a zero group takes no bytes, an 8-bit group 16 (the chain reads every step from a table)
	asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + kWideRingBytes + lane * 16), "r"(0u) : "memory");
	asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(base + kWideRingBytes + 2048 + lane * 16), "r"(0x10101010u) : "memory");
	__syncwarp();
	for (;;)
	{
		uint32_t s = 0;
		if (lane == 0)
			s = atomicAdd(T.counters + 1, 1u);
		s = __shfl_sync(0xffffffffu, s, 0);
		if (s >= T.n_streams)
			break;
		walk_stream_wide(T, s, lane, base);
		__syncwarp();
	}
}

} // namespace mob200
