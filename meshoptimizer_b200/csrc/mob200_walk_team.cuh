// mob200_walk_team.cuh -- phase 1 for batches of FEW LONG streams: the offsets-only team walk.
//
// A stream stores no index (reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance one pointer), so the
// offset chain of one stream is serial whatever the hardware.  What can be taken OFF that chain is everything that
// does not depend on the running offset.  One CTA walks one stream as a team:
//
//   helper warps   stream the encoded bytes through a shared-memory ring in 512-byte windows and, for EVERY byte
//                  position p of a window and every field width w in {1, 2, 4} bits, work out how many bytes a group of
//                  that width starting at p would take (2w + its all-ones fields: SWAR per-byte counts, sliding-window
//                  sums by doubling) into position-indexed step tables T_w, ahead of the chain, window after window;
//   chain warp     follows the stream: control bytes (turned into bit masks, so that only bit-packed byte-channels cost
//                  a loop iteration: runs of literal / zero channels are one popcount), per bit-packed channel the
//                  header, then ONE table byte and one add per group (add -> LDS -> add, ~40 cycles; which table a step
//                  reads is worked out from the header by the lanes in parallel, off the chain); it writes nothing but the
//                  end offset of every block (the block-offset table, DevTables::block_offset) and the stream's reference
//                  return code (:1827-1869).
//
// Measured (r2, B200): one monolithic stream 0.63 -> 1.2 GB/s, 1024 x 65536-vertex streams 7.95 -> 4.4 ms against the
// fused one-warp-per-stream walker.  What is left is the latency of a lone warp's instruction stream (~6 cycles per
// instruction, ~2500 per block), not the table look-ups: step tables for PAIRS of equal-width groups (half the chain
// steps, 2.5x the helpers' work) measured 5-15 % slower and were dropped.
//
// The decode then runs in block mode (mob200_walker.cuh walk_group<true>): every block is walked again by its own lane
// -- this time in parallel, group-table rows and all -- and decoded.
#pragma once

#include "mob200_device.cuh"
#include "mob200_walker.cuh"      // walk_framing
#include "mob200_walker_wide.cuh" // bytes_n1 / bytes_n2 / bytes_n4, window_sums

namespace mob200
{

constexpr uint32_t kTeamHelpers = 2;
constexpr uint32_t kTeamThreads = 32 * (1 + kTeamHelpers);
constexpr uint32_t kTeamWindow = 512;                          // bytes per window
constexpr uint32_t kTeamWindows = 4;                           // windows in the ring
constexpr uint32_t kTeamRing = kTeamWindow * kTeamWindows;     // 2 KB
constexpr uint32_t kTeamMirror = 512;                          // the first 512 entries again behind the end: a channel (<= 4 + 16 * 24 bytes) never wraps
constexpr uint32_t kTeamTable = kTeamRing + kTeamMirror;
// shared memory: data ring (+ mirror); five step tables indexed by ring position, one per width index {0,1,2,4,8 bits}
// (0 and 8 are constants: 0 and 16 bytes); the flags
constexpr uint32_t kTeamSmemData = 0;
constexpr uint32_t kTeamSmemT = kTeamSmemData + kTeamTable;
constexpr uint32_t kTeamSmemBars = kTeamSmemT + 5 * kTeamTable; // mbarriers full[kTeamWindows] (helper -> chain), empty[kTeamWindows] (chain -> helpers)
constexpr uint32_t kTeamSmemStop = kTeamSmemBars + 2 * kTeamWindows * 8; // set by the chain when it is done with the stream
constexpr uint32_t kTeamSmemBytes = kTeamSmemStop + 16;
static_assert((kTeamSmemBars & 7) == 0, "alignment");

__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}

// T_w of the 16 positions that start at `piece` (relative to org): bytes a w-bit group starting there takes
template <int kWidthIndex>
__device__ __forceinline__ void team_sizes(const uint32_t w[6], uint32_t out[4])
{
	uint32_t n[6];
#pragma unroll
	for (int j = 0; j < 6; ++j)
		n[j] = kWidthIndex == 1 ? bytes_n1(w[j]) : (kWidthIndex == 2 ? bytes_n2(w[j]) : bytes_n4(w[j]));
	window_sums<(kWidthIndex == 1 ? 2 : (kWidthIndex == 2 ? 4 : 8))>(n, out);
	const uint32_t fixed = (kWidthIndex == 1 ? 0x02020202u : (kWidthIndex == 2 ? 0x04040404u : 0x08080808u));
#pragma unroll
	for (int j = 0; j < 4; ++j)
		out[j] += fixed;
}

// 24 bytes from `piece` on (16-byte aligned), never beyond rel_limit
__device__ __forceinline__ void team_load(const uint8_t* org, uint32_t piece, uint32_t rel_limit, uint32_t w[6])
{
#pragma unroll
	for (int j = 0; j < 6; ++j)
		w[j] = 0;
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
}

// ---- helper warps: window k of the stream -> data ring + step tables --------------------------------------------------
__device__ void team_helper(uint32_t helper, uint32_t lane, uint8_t* base, const uint8_t* org, uint32_t rel_limit, uint32_t n_windows)
{
	const uint32_t smem = smem_addr(base);
	uint64_t* full = reinterpret_cast<uint64_t*>(base + kTeamSmemBars);
	uint64_t* empty = full + kTeamWindows;
	uint32_t* stop = reinterpret_cast<uint32_t*>(base + kTeamSmemStop);
	for (uint32_t k = helper; k < n_windows; k += kTeamHelpers)
	{
		// the slot's previous window (k - kTeamWindows) must have been released by the chain
		bool stopped = false;
		if (k >= kTeamWindows)
			while (!mbar_try_wait(empty + k % kTeamWindows, (k / kTeamWindows - 1u) & 1u))
				if (atomicAdd(stop, 0u)) // (the chain has finished, or given up on, the stream)
				{
					stopped = true;
					break;
				}
		if (__any_sync(0xffffffffu, stopped))
			break;

		const uint32_t piece = k * kTeamWindow + lane * 16;
		const uint32_t slot = (k % kTeamWindows) * kTeamWindow + lane * 16;
		const bool mirror = (k % kTeamWindows) == 0; // the first window of the ring is kept twice
		uint32_t w[6], t[4];
		team_load(org, piece, rel_limit, w);
		sts_v4(smem + kTeamSmemData + slot, w[0], w[1], w[2], w[3]);
		if (mirror)
			sts_v4(smem + kTeamSmemData + kTeamRing + slot, w[0], w[1], w[2], w[3]);
		team_sizes<1>(w, t);
		sts_v4(smem + kTeamSmemT + 1 * kTeamTable + slot, t[0], t[1], t[2], t[3]);
		if (mirror)
			sts_v4(smem + kTeamSmemT + 1 * kTeamTable + kTeamRing + slot, t[0], t[1], t[2], t[3]);
		team_sizes<2>(w, t);
		sts_v4(smem + kTeamSmemT + 2 * kTeamTable + slot, t[0], t[1], t[2], t[3]);
		if (mirror)
			sts_v4(smem + kTeamSmemT + 2 * kTeamTable + kTeamRing + slot, t[0], t[1], t[2], t[3]);
		team_sizes<3>(w, t);
		sts_v4(smem + kTeamSmemT + 3 * kTeamTable + slot, t[0], t[1], t[2], t[3]);
		if (mirror)
			sts_v4(smem + kTeamSmemT + 3 * kTeamTable + kTeamRing + slot, t[0], t[1], t[2], t[3]);
		mbar_arrive(full + k % kTeamWindows); // every lane arrives for its own stores
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
__device__ __forceinline__ void team_need(uint8_t* base, uint32_t rel, uint32_t span, uint32_t n_windows, uint32_t& have_until, uint32_t& released)
{
	uint64_t* full = reinterpret_cast<uint64_t*>(base + kTeamSmemBars);
	uint64_t* empty = full + kTeamWindows;
	const uint32_t first = min(rel / kTeamWindow, n_windows);
	uint32_t last = (rel + span) / kTeamWindow;
	if (last >= n_windows)
		last = n_windows ? n_windows - 1 : 0;
	for (;;)
	{
		// Release first, wait second: the helper that builds the window waited for below may need the slot of a window this
		// warp has left behind (a run of literal channels jumps over several windows at once).  A window is only released
		// after it has been waited for, so that the phases of its two barriers stay paired.
		for (; released < first && released < have_until; ++released)
			mbar_arrive(empty + released % kTeamWindows); // every lane arrives for its own reads
		if (!(have_until <= last && have_until < n_windows))
			break;
		mbar_wait(full + have_until % kTeamWindows, (have_until / kTeamWindows) & 1u);
		++have_until;
	}
}

// even bits of x (bit 2i -> bit i)
__device__ __forceinline__ uint32_t even_bits(uint32_t x)
{
	x &= 0x55555555u;
	x = (x | (x >> 1)) & 0x33333333u;
	x = (x | (x >> 2)) & 0x0f0f0f0fu;
	x = (x | (x >> 4)) & 0x00ff00ffu;
	return (x | (x >> 8)) & 0xffffu;
}

// One bit-packed byte-channel: header at rel, then `groups` groups.  Returns false if the channel is malformed.
__device__ __forceinline__ bool team_channel(uint8_t* base, uint32_t& rel, uint32_t rel_end, uint32_t groups, uint32_t hdr, uint32_t version, uint32_t ctrl,
    uint32_t n_windows, uint32_t& have_until, uint32_t& released, uint32_t lane)
{
	const uint32_t smem = smem_addr(base);
	if (rel_end - rel < hdr) // (:1376)
		return false;
	team_need(base, rel, hdr + 16 * kGroupReadLimit + 8, n_windows, have_until, released);
	uint32_t sel_bits = team_u32(smem, rel);
	rel += hdr;

	// The chain: one table byte per group.  Which table a step reads depends on the header only: lane g works out the table of
	// group g, the descriptors are broadcast up front, and the dependent chain itself is nothing but  LDS -> add -> add  per
	// step.  The reference's 24-byte rule (:1385,:1415) is tested once, on the position of the last group (positions only grow).
	// Tables are indexed by ring position; a channel never wraps thanks to the mirrored head of the ring.
	uint32_t mine = 0; // table of group `lane` (offset from `smem`); 0 = no such group
	{
		const uint32_t g = lane & 15u;
		const uint32_t sel = (sel_bits >> (2 * g)) & 3u;
		const uint32_t idx = version ? sel + ctrl : (sel ? sel + 1u : 0u);
		mine = g < groups ? kTeamSmemT + idx * kTeamTable : 0u;
	}
	const uint32_t p0 = rel & (kTeamRing - 1);
	uint32_t p = p0, p_last = p0;
#pragma unroll
	for (uint32_t g = 0; g < 16; ++g)
	{
		const uint32_t a = __shfl_sync(0xffffffffu, mine, g);
		if (a)
		{
			p_last = p;
			p += lds_u8(smem + a + p);
		}
	}
	const uint32_t rel_last = rel + (p_last - p0);
	rel += p - p0;
	return !(rel_last > rel_end || rel_end - rel_last < kGroupReadLimit);
}

__device__ void team_chain(const DevTables& T, uint32_t s, uint32_t lane, uint8_t* base, int framing_status, uint32_t version, uint32_t n_windows)
{
	const uint32_t smem = smem_addr(base);
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
			const uint8_t* control = src + (rel - rel0);

			// byte-channels in chunks of 32: their control values (2 bits each; v0: all bit-packed) become two bit masks
			for (uint32_t k0 = 0; k0 < vs && !bad; k0 += 32)
			{
				const uint32_t kn = min(32u, vs - k0);
				const uint32_t live = kn == 32 ? 0xffffffffu : ((1u << kn) - 1u);
				uint32_t c_lo = 0, c_hi = 0; // control bits of channels k0 .. k0 + 15 / k0 + 16 .. k0 + 31
				if (version)
				{
					if (k0 == 0)
					{
						team_need(base, rel, 12, n_windows, have_until, released);
						c_lo = team_u32(smem, rel), c_hi = team_u32(smem, rel + 4);
						rel += ctrl_bytes;
					}
					else
					{
						// (vertices wider than 32 bytes: the rest of the control bytes comes from global memory -- the ring has moved on)
						for (uint32_t j = 0; j < 4 && k0 / 4 + j < ctrl_bytes; ++j)
							c_lo |= (uint32_t)__ldg(control + k0 / 4 + j) << (8 * j);
						for (uint32_t j = 0; j < 4 && k0 / 4 + 4 + j < ctrl_bytes; ++j)
							c_hi |= (uint32_t)__ldg(control + k0 / 4 + 4 + j) << (8 * j);
					}
				}
				const uint32_t bit0 = even_bits(c_lo) | (even_bits(c_hi) << 16), bit1 = even_bits(c_lo >> 1) | (even_bits(c_hi >> 1) << 16);
				uint32_t packed = ~bit1 & live;            // control 0 / 1
				const uint32_t literal = bit1 & bit0 & live; // control 3 (n raw bytes); control 2 stores nothing
				uint32_t below_done = 0;                   // channels of the chunk already accounted for
				while (!bad)
				{
					const uint32_t k = packed ? (uint32_t)__ffs((int)packed) - 1u : 32u;
					const uint32_t upto = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);
					// the literal channels in front of channel k (:1546-1554: the 16-aligned count must be readable for each;
					// positions only grow, so the last one of the run decides)
					const uint32_t lits = (uint32_t)__popc(literal & upto & ~below_done);
					if (lits)
					{
						if (rel_end - rel < (lits - 1) * n || rel_end - rel - (lits - 1) * n < na)
						{
							bad = true;
							break;
						}
						rel += lits * n;
					}
					if (k >= 32)
						break;
					const uint32_t ctrl = (((k < 16 ? c_lo : c_hi) >> ((k & 15u) * 2u)) & 3u);
					if (!team_channel(base, rel, rel_end, groups, hdr, version, ctrl, n_windows, have_until, released, lane))
					{
						bad = true;
						break;
					}
					packed &= packed - 1u;
					below_done = upto | (1u << k);
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
		// helpers that wait for ring space see this and leave
		atomicExch(reinterpret_cast<uint32_t*>(base + kTeamSmemStop), 1u);
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
			sts_v4(smem + kTeamSmemT + i, 0u, 0u, 0u, 0u);
			sts_v4(smem + kTeamSmemT + 4 * kTeamTable + i, 0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u);
		}
		if (threadIdx.x == 0)
		{
			uint64_t* bars = reinterpret_cast<uint64_t*>(team_smem + kTeamSmemBars);
			for (uint32_t i = 0; i < 2 * kTeamWindows; ++i)
				mbar_init(bars + i, 32); // full: the 32 lanes of the helper that built the window; empty: the 32 lanes of the chain
			*reinterpret_cast<uint32_t*>(team_smem + kTeamSmemStop) = 0;
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();

		const DevStream* d = T.streams + s;
		uint32_t version = 0;
		const int framing = walk_framing(d->src, d->src_size, d->vertex_size, d->nblocks, version);
		const uint32_t rel0 = (uint32_t)(reinterpret_cast<uintptr_t>(d->src) & 15u);
		const uint32_t rel_limit = (framing == 0 && d->nblocks) ? ((rel0 + d->src_size + 15u) & ~15u) : 0u;
		const uint32_t n_windows = (rel_limit + kTeamWindow - 1) / kTeamWindow;
		if (warp == kTeamHelpers)
			team_chain(T, s, lane, team_smem, framing, version, n_windows);
		else
			team_helper(warp, lane, team_smem, d->src - rel0, rel_limit, n_windows);
		__syncthreads();
	}
}

} // namespace mob200
