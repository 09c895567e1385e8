// mob200_walk_team.cuh -- phase 1 for batches of FEW LONG streams: the offsets-only team walk.
//
// A stream stores no index (reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance one pointer), so the
// offset chain of one stream is serial whatever the hardware.  What can be taken OFF that chain is everything that
// does not depend on the running offset.  One CTA walks one stream as a team:
//
//   helper warps   stream the encoded bytes through a shared-memory ring in 512-byte windows and, for EVERY byte
//                  position p of a window and every field width w in {1, 2, 4} bits, work out how many bytes a group of
//                  that width starting at p would take (2w + its all-ones fields: SWAR per-byte counts, sliding-window
//                  sums by doubling) into position-indexed step tables -- ahead of the chain, window after window;
//   chain warp     follows the stream: control bytes, per bit-packed byte-channel the header, then ONE table byte and one
//                  add per group (add -> LDS -> add, ~33 cycles), literal and zero channels in O(1); it writes nothing but
//                  the end offset of every block (the block-offset table, DevTables::block_offset) and the stream's
//                  reference return code (:1827-1869).
//
// The decode then runs in block mode (mob200_walker.cuh walk_group<true>): every block is walked again by its own lane
// -- this time in parallel, group-table rows and all -- and decoded.  Against the fused one-warp-per-stream walker this
// replaces, the serial part per bit-packed channel drops from ~2400 cycles (tables built on the chain's own warp) to
// ~700.
#pragma once

#include "mob200_device.cuh"
#include "mob200_walker.cuh"      // walk_framing
#include "mob200_walker_wide.cuh" // bytes_n1 / bytes_n2 / bytes_n4, window_sums

namespace mob200
{

constexpr uint32_t kTeamHelpers = 2;
constexpr uint32_t kTeamThreads = 32 * (1 + kTeamHelpers);
constexpr uint32_t kTeamWindow = 512;                          // bytes per window
constexpr uint32_t kTeamWindows = 8;                           // windows in the ring
constexpr uint32_t kTeamRing = kTeamWindow * kTeamWindows;     // 4 KB
constexpr uint32_t kTeamMirror = 512;                          // the first 512 entries again behind the end: a channel (<= 16 + 16 * 24 bytes) never wraps
constexpr uint32_t kTeamTable = kTeamRing + kTeamMirror;
// shared memory: data ring (+ mirror), five step tables indexed by ring position (widths {0,1,2,4,8}: 0 and 8 are constant), flags
constexpr uint32_t kTeamSmemData = 0;
constexpr uint32_t kTeamSmemTables = kTeamSmemData + kTeamTable;
constexpr uint32_t kTeamSmemFlags = kTeamSmemTables + 5 * kTeamTable; // ready[kTeamWindows], consumed
constexpr uint32_t kTeamSmemBytes = kTeamSmemFlags + (kTeamWindows + 1) * 4 + 12;
constexpr uint32_t kTeamStop = 0x7fffffffu; // "consumed" value that tells the helpers the chain is done

__device__ __forceinline__ uint32_t lds_acquire_u32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
	return v;
}

__device__ __forceinline__ void sts_release_u32(uint32_t a, uint32_t v)
{
	asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// ---- helper warps: window k of the stream -> data ring + step tables --------------------------------------------------
__device__ void team_helper(uint32_t helper, uint32_t lane, uint32_t smem, const uint8_t* org, uint32_t rel_limit, uint32_t n_windows)
{
	const uint32_t flags = smem + kTeamSmemFlags;
	for (uint32_t k = helper; k < n_windows; k += kTeamHelpers)
	{
		// the slot's previous window (k - kTeamWindows) must not be needed by the chain any more
		uint32_t consumed = lds_acquire_u32(flags + kTeamWindows * 4);
		while (k >= kTeamWindows && consumed != kTeamStop && consumed + kTeamWindows <= k)
		{
			__nanosleep(40);
			consumed = lds_acquire_u32(flags + kTeamWindows * 4);
		}
		if (consumed == kTeamStop)
			break; // the chain has finished (or given up on) the stream

		// 24 bytes from the lane's 16-byte piece on (the sliding windows look up to 7 bytes past a position, and the last
		// lane's run into the next window: its bytes are read straight from global memory, never beyond rel_limit)
		const uint32_t piece = k * kTeamWindow + lane * 16;
		uint32_t w[6] = {0, 0, 0, 0, 0, 0};
		if (piece < rel_limit)
		{
			const uint4 a = __ldg(reinterpret_cast<const uint4*>(org + piece));
			w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w;
		}
		if (piece + 16 < rel_limit)
		{
			const uint2 b = __ldg(reinterpret_cast<const uint2*>(org + piece + 16));
			w[4] = b.x, w[5] = b.y;
		}
		const uint32_t pos = (k % kTeamWindows) * kTeamWindow + lane * 16;
		const bool mirror = (k % kTeamWindows) == 0; // the first window of the ring is kept twice
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem + kTeamSmemData + pos), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
		if (mirror)
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem + kTeamSmemData + kTeamRing + pos), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");

		uint32_t n[6], out[4];
		// width index 1: 1-bit fields, windows of 2 bytes; 2: 2-bit, 4 bytes; 3: 4-bit, 8 bytes.  Entry = bytes the group takes.
#pragma unroll
		for (int j = 0; j < 6; ++j)
			n[j] = bytes_n1(w[j]);
		window_sums<2>(n, out);
		{
			const uint32_t t = smem + kTeamSmemTables + 1 * kTeamTable + pos;
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(out[0] + 0x02020202u), "r"(out[1] + 0x02020202u), "r"(out[2] + 0x02020202u), "r"(out[3] + 0x02020202u) : "memory");
			if (mirror)
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t + kTeamRing), "r"(out[0] + 0x02020202u), "r"(out[1] + 0x02020202u), "r"(out[2] + 0x02020202u), "r"(out[3] + 0x02020202u) : "memory");
		}
#pragma unroll
		for (int j = 0; j < 6; ++j)
			n[j] = bytes_n2(w[j]);
		window_sums<4>(n, out);
		{
			const uint32_t t = smem + kTeamSmemTables + 2 * kTeamTable + pos;
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(out[0] + 0x04040404u), "r"(out[1] + 0x04040404u), "r"(out[2] + 0x04040404u), "r"(out[3] + 0x04040404u) : "memory");
			if (mirror)
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t + kTeamRing), "r"(out[0] + 0x04040404u), "r"(out[1] + 0x04040404u), "r"(out[2] + 0x04040404u), "r"(out[3] + 0x04040404u) : "memory");
		}
#pragma unroll
		for (int j = 0; j < 6; ++j)
			n[j] = bytes_n4(w[j]);
		window_sums<8>(n, out);
		{
			const uint32_t t = smem + kTeamSmemTables + 3 * kTeamTable + pos;
			asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(out[0] + 0x08080808u), "r"(out[1] + 0x08080808u), "r"(out[2] + 0x08080808u), "r"(out[3] + 0x08080808u) : "memory");
			if (mirror)
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(t + kTeamRing), "r"(out[0] + 0x08080808u), "r"(out[1] + 0x08080808u), "r"(out[2] + 0x08080808u), "r"(out[3] + 0x08080808u) : "memory");
		}
		__syncwarp();
		if (lane == 0)
			sts_release_u32(flags + (k % kTeamWindows) * 4, k + 1);
	}
}

// ---- chain warp ---------------------------------------------------------------------------------------------------------

// 32 bits of the stream at relative position rel (ring position = rel mod ring; reads run into the mirror, never wrap)
__device__ __forceinline__ uint32_t team_u32(uint32_t smem, uint32_t rel)
{
	const uint32_t p = rel & (kTeamRing - 1);
	uint32_t lo, hi;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(smem + kTeamSmemData + (p & ~3u)));
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(smem + kTeamSmemData + (p & ~3u) + 4));
	return __funnelshift_r(lo, hi, (p & 3u) * 8u);
}

// windows [rel / 512, (rel + span) / 512] are built; everything below rel's window may be overwritten
__device__ __forceinline__ void team_need(uint32_t smem, uint32_t rel, uint32_t span, uint32_t n_windows, uint32_t& have_until, uint32_t& released, uint32_t lane)
{
	const uint32_t first = rel / kTeamWindow;
	uint32_t last = (rel + span) / kTeamWindow;
	if (last >= n_windows)
		last = n_windows ? n_windows - 1 : 0;
	if (first > released)
	{
		released = first;
		if (lane == 0)
			sts_release_u32(smem + kTeamSmemFlags + kTeamWindows * 4, first);
	}
	while (have_until <= last && have_until < n_windows)
	{
		const uint32_t flag = smem + kTeamSmemFlags + (have_until % kTeamWindows) * 4;
		while (lds_acquire_u32(flag) != have_until + 1)
			__nanosleep(20);
		++have_until;
	}
}

__device__ void team_chain(const DevTables& T, uint32_t s, uint32_t lane, uint32_t smem, int framing_status, uint32_t version, uint32_t n_windows)
{
	const DevStream* d = T.streams + s;
	const uint8_t* src = d->src;
	const uint32_t size = d->src_size;
	const uint32_t vs = d->vertex_size;
	const uint32_t count = d->vertex_count;
	const uint32_t bv = d->block_groups * kGroup;
	const uint32_t nblocks = d->nblocks;
	uint32_t* boff = T.block_offset + d->block_base + s;

	int status = framing_status;
	const uint32_t rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
	const uint32_t rel_end = rel0 + size;
	uint32_t rel = rel0 + 1;
	uint32_t done = 0;
	const bool framed = status == 0;
	uint32_t have_until = 0, released = 0;
	const uint32_t tab = smem + kTeamSmemTables;

	if (framed && nblocks)
	{
		if (lane == 0)
			boff[0] = 1;
		for (uint32_t b = 0; b < nblocks && status == 0; ++b)
		{
			const uint32_t n = min(bv, count - b * bv);
			const uint32_t groups = (n + kGroup - 1) / kGroup;
			const uint32_t na = groups * kGroup;
			const uint32_t hdr = (groups + 3) / 4;
			const uint32_t ctrl_bytes = version ? vs / 4 : 0;
			bool bad = rel_end - rel < ctrl_bytes;

			// control bytes of the first 64 byte-channels are kept in registers; wider vertices read the rest from global memory
			const uint8_t* control = src + (rel - rel0);
			uint32_t cw0 = 0, cw1 = 0, cw2 = 0, cw3 = 0;
			if (!bad && version)
			{
				team_need(smem, rel, 20, n_windows, have_until, released, lane);
				cw0 = team_u32(smem, rel), cw1 = team_u32(smem, rel + 4), cw2 = team_u32(smem, rel + 8), cw3 = team_u32(smem, rel + 12);
			}
			if (!bad)
				rel += ctrl_bytes;

			for (uint32_t k = 0; k < vs && !bad; ++k)
			{
				uint32_t cbyte;
				if (k < 64)
				{
					const uint32_t wsel = k >> 4;
					const uint32_t word = wsel == 0 ? cw0 : (wsel == 1 ? cw1 : (wsel == 2 ? cw2 : cw3));
					cbyte = (word >> (((k >> 2) & 3u) * 8)) & 0xffu;
				}
				else
					cbyte = version ? __ldg(control + (k >> 2)) : 0u;
				const uint32_t ctrl = (cbyte >> ((k & 3) * 2)) & 3u;

				if (ctrl == 3)
				{
					if (rel_end - rel < na) // literal bytes (:1546-1554): the 16-aligned count must be readable
					{
						bad = true;
						break;
					}
					rel += n;
				}
				else if (ctrl != 2)
				{
					if (rel_end - rel < hdr) // (:1376)
					{
						bad = true;
						break;
					}
					team_need(smem, rel, hdr + 16 * kGroupReadLimit + 8, n_windows, have_until, released, lane);
					uint32_t sel_bits = team_u32(smem, rel);
					rel += hdr;

					// the chain: one table byte per group (tables indexed by ring position; a channel never wraps thanks to the
					// mirrored head of the ring).  The 24-byte rule (:1385,:1415) is tested once, on the last group's position.
					uint32_t p = rel & (kTeamRing - 1);
					const uint32_t p0 = p;
					uint32_t p_last = p;
					uint32_t idx = (sel_bits & 3u) + (version ? ctrl : (uint32_t)((sel_bits & 3u) != 0u));
					uint32_t slot_base = tab + idx * kTeamTable;
#pragma unroll 4
					for (uint32_t g = 0; g < groups; ++g)
					{
						uint32_t step;
						asm volatile("ld.shared.u8 %0, [%1];" : "=r"(step) : "r"(slot_base + p));
						p_last = p;
						sel_bits >>= 2;
						idx = (sel_bits & 3u) + (version ? ctrl : (uint32_t)((sel_bits & 3u) != 0u));
						slot_base = tab + idx * kTeamTable;
						asm volatile("" : "+r"(slot_base)); // (keeps the table base out of the sum with the running position)
						p += step;
					}
					const uint32_t rel_last = rel + (p_last - p0);
					rel += p - p0;
					if (rel_last > rel_end || rel_end - rel_last < kGroupReadLimit)
					{
						bad = true;
						break;
					}
				}
			}

			if (bad)
			{
				status = -2;
				break;
			}
			done = b + 1;
			if (lane == 0)
				boff[b + 1] = rel - rel0;
		}
	}

	if (framed && status == 0 && rel_end - rel != tail_padded(vs, version))
		status = -3; // (:1868-1869) the blocks were decodable, the stream is still rejected

	if (lane == 0)
	{
		if (done < nblocks)
			for (uint32_t b = framed ? done + 1 : 0; b <= nblocks; ++b)
				boff[b] = kInvalidOffset;
		T.status[d->caller_index] = status;
		// let helpers that wait for ring space run to their end
		sts_release_u32(smem + kTeamSmemFlags + kTeamWindows * 4, kTeamStop);
	}
}

// One CTA per stream (grid-stride over the streams): the LAST warp is the chain (the issue arbiter prefers the highest
// warp id: the chain is the critical path), the warps before it are the helpers.
__global__ void __launch_bounds__(kTeamThreads) walk_team_kernel(DevTables T)
{
	extern __shared__ __align__(16) uint8_t team_smem[];
	const uint32_t smem = smem_addr(team_smem);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

	for (uint32_t s = blockIdx.x; s < T.n_streams; s += gridDim.x)
	{
		// constant step tables (a zero group takes no bytes, an 8-bit group 16) and the flags
		for (uint32_t i = threadIdx.x * 16; i < kTeamTable; i += kTeamThreads * 16)
		{
			asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem + kTeamSmemTables + i), "r"(0u) : "memory");
			asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(smem + kTeamSmemTables + 4 * kTeamTable + i), "r"(0x10101010u) : "memory");
		}
		if (threadIdx.x <= kTeamWindows)
			reinterpret_cast<volatile uint32_t*>(team_smem + kTeamSmemFlags)[threadIdx.x] = 0;
		__syncthreads();

		const DevStream* d = T.streams + s;
		uint32_t version = 0;
		const int framing = walk_framing(d->src, d->src_size, d->vertex_size, d->nblocks, version);
		const uint32_t rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(d->src) & 15u);
		const uint32_t rel_limit = (framing == 0 && d->nblocks) ? ((rel0 + d->src_size + 15u) & ~15u) : 0u;
		const uint32_t n_windows = (rel_limit + kTeamWindow - 1) / kTeamWindow;
		if (warp == kTeamHelpers)
			team_chain(T, s, lane, smem, framing, version, n_windows);
		else
			team_helper(warp, lane, smem, d->src - rel0, rel_limit, n_windows);
		__syncthreads();
	}
}

} // namespace mob200
