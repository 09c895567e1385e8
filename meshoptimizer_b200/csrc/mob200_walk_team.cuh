// mob200_walk_team.cuh -- phase 1 for batches of FEW LONG streams: the offsets-only team walk.
//
// A stream stores no index (reference src/vertexcodec.cpp:1375-1425,1531-1568,1857-1866 advance one pointer), so the
// offset chain of one stream is serial whatever the hardware.  What can be taken OFF that chain is everything that
// does not depend on the running offset.  One CTA walks one stream as a team:
//
//   helper warps   stream the encoded bytes through a shared-memory ring in 512-byte windows and, for EVERY byte
//                  position p of a window and every field width w in {1, 2, 4} bits, work out how many bytes a group of
//                  that width starting at p would take (2w + its all-ones fields: SWAR per-byte counts, sliding-window
//                  sums by doubling) -- table T_w -- and how many bytes TWO consecutive groups of that width would
//                  take, P_w[p] = T_w[p] + T_w[p + T_w[p]], into position-indexed step tables, ahead of the chain,
//                  window after window;
//   chain warp     follows the stream: control bytes (turned into bit masks, so that only bit-packed byte-channels cost
//                  a loop iteration: runs of literal / zero channels are one popcount), per bit-packed channel the
//                  header, then ONE table byte and one add per group or per PAIR of groups of equal width
//                  (add -> LDS -> add, ~40 cycles); it writes nothing but the end offset of every block (the
//                  block-offset table, DevTables::block_offset) and the stream's reference return code (:1827-1869).
//
// The decode then runs in block mode (mob200_walker.cuh walk_group<true>): every block is walked again by its own lane
// -- this time in parallel, group-table rows and all -- and decoded.
#pragma once

#include "mob200_device.cuh"
#include "mob200_walker.cuh"      // walk_framing
#include "mob200_walker_wide.cuh" // bytes_n1 / bytes_n2 / bytes_n4, window_sums

namespace mob200
{

constexpr uint32_t kTeamHelpers = 3;
constexpr uint32_t kTeamThreads = 32 * (1 + kTeamHelpers);
constexpr uint32_t kTeamWindow = 512;                          // bytes per window
constexpr uint32_t kTeamWindows = 4;                           // windows in the ring
constexpr uint32_t kTeamRing = kTeamWindow * kTeamWindows;     // 2 KB
constexpr uint32_t kTeamMirror = 512;                          // the first 512 entries again behind the end: a channel (<= 4 + 16 * 24 bytes) never wraps
constexpr uint32_t kTeamTable = kTeamRing + kTeamMirror;
constexpr uint32_t kTeamStage = 512 + 32;                      // a helper's private copy of T_w for its window and the 32 positions behind it
// shared memory: data ring (+ mirror); step tables indexed by ring position -- T for the width indices {0,1,2,4,8 bits}
// (0 and 8 are constants: 0 and 16 bytes), P for {1,2,4,8 bits} (a pair of zero groups is the T_0 table again); the
// helpers' staging areas; the flags
constexpr uint32_t kTeamSmemData = 0;
constexpr uint32_t kTeamSmemT = kTeamSmemData + kTeamTable;
constexpr uint32_t kTeamSmemP = kTeamSmemT + 5 * kTeamTable; // P tables of width indices 1..4 at (idx - 1)
constexpr uint32_t kTeamSmemStage = kTeamSmemP + 4 * kTeamTable;
constexpr uint32_t kTeamSmemFlags = kTeamSmemStage + kTeamHelpers * 3 * kTeamStage; // ready[kTeamWindows], consumed
constexpr uint32_t kTeamSmemBytes = kTeamSmemFlags + (kTeamWindows + 1) * 4 + 12;
constexpr uint32_t kTeamStop = 0x7fffffffu; // "consumed" value that tells the helpers the chain is done
static_assert((kTeamSmemStage & 15) == 0 && (kTeamStage & 15) == 0 && (kTeamSmemFlags & 3) == 0, "alignment");

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
__device__ void team_helper(uint32_t helper, uint32_t lane, uint32_t smem, const uint8_t* org, uint32_t rel_limit, uint32_t n_windows)
{
	const uint32_t flags = smem + kTeamSmemFlags;
	const uint32_t stage = smem + kTeamSmemStage + helper * 3 * kTeamStage;
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

		const uint32_t piece = k * kTeamWindow + lane * 16;
		const uint32_t slot = (k % kTeamWindows) * kTeamWindow + lane * 16;
		const bool mirror = (k % kTeamWindows) == 0; // the first window of the ring is kept twice
		uint32_t w[6], t1[4], t2[4], t4[4];
		team_load(org, piece, rel_limit, w);
		sts_v4(smem + kTeamSmemData + slot, w[0], w[1], w[2], w[3]);
		if (mirror)
			sts_v4(smem + kTeamSmemData + kTeamRing + slot, w[0], w[1], w[2], w[3]);
		team_sizes<1>(w, t1);
		team_sizes<2>(w, t2);
		team_sizes<3>(w, t4);
		sts_v4(stage + 0 * kTeamStage + lane * 16, t1[0], t1[1], t1[2], t1[3]);
		sts_v4(stage + 1 * kTeamStage + lane * 16, t2[0], t2[1], t2[2], t2[3]);
		sts_v4(stage + 2 * kTeamStage + lane * 16, t4[0], t4[1], t4[2], t4[3]);
		// the 32 positions behind the window (a second group of a pair may start there): lanes 0 and 1 again
		if (lane < 2)
		{
			uint32_t wx[6], x1[4], x2[4], x4[4];
			team_load(org, (k + 1) * kTeamWindow + lane * 16, rel_limit, wx);
			team_sizes<1>(wx, x1);
			team_sizes<2>(wx, x2);
			team_sizes<3>(wx, x4);
			sts_v4(stage + 0 * kTeamStage + 512 + lane * 16, x1[0], x1[1], x1[2], x1[3]);
			sts_v4(stage + 1 * kTeamStage + 512 + lane * 16, x2[0], x2[1], x2[2], x2[3]);
			sts_v4(stage + 2 * kTeamStage + 512 + lane * 16, x4[0], x4[1], x4[2], x4[3]);
		}
		__syncwarp();

		// pairs: P_w[p] = T_w[p] + T_w[p + T_w[p]] (T <= 24: the second group starts inside the staged range)
#pragma unroll
		for (int wi = 0; wi < 3; ++wi)
		{
			const uint32_t* t = wi == 0 ? t1 : (wi == 1 ? t2 : t4);
			const uint32_t sbase = stage + wi * kTeamStage + lane * 16;
			uint32_t p[4];
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				uint32_t acc = 0;
#pragma unroll
				for (int b = 0; b < 4; ++b)
				{
					const uint32_t first = (t[j] >> (8 * b)) & 0xffu;
					acc |= (first + lds_u8(sbase + 4 * j + b + first)) << (8 * b);
				}
				p[j] = acc;
			}
			const uint32_t tt = smem + kTeamSmemT + (wi + 1) * kTeamTable + slot;
			const uint32_t pt = smem + kTeamSmemP + wi * kTeamTable + slot;
			sts_v4(tt, t[0], t[1], t[2], t[3]);
			sts_v4(pt, p[0], p[1], p[2], p[3]);
			if (mirror)
			{
				sts_v4(tt + kTeamRing, t[0], t[1], t[2], t[3]);
				sts_v4(pt + kTeamRing, p[0], p[1], p[2], p[3]);
			}
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
__device__ __forceinline__ bool team_channel(uint32_t smem, uint32_t& rel, uint32_t rel_end, uint32_t groups, uint32_t hdr, uint32_t version, uint32_t ctrl,
    uint32_t n_windows, uint32_t& have_until, uint32_t& released, uint32_t lane)
{
	if (rel_end - rel < hdr) // (:1376)
		return false;
	team_need(smem, rel, hdr + 16 * kGroupReadLimit + 8, n_windows, have_until, released, lane);
	uint32_t sel_bits = team_u32(smem, rel);
	rel += hdr;

	// The chain: one table byte per group, or per PAIR of groups (2j, 2j+1) with the same selector -- never the last group,
	// whose position is needed for the reference's 24-byte rule (:1385,:1415), tested once, there.  Which table a step
	// reads depends on the header only: lane j works out the (at most two) steps of pair slot j, the eight descriptors
	// are broadcast up front, and the dependent chain itself is nothing but  LDS -> add -> add  per step.
	// Tables are indexed by ring position; a channel never wraps thanks to the mirrored head of the ring.
	uint32_t mine = 0; // table base of step A | table base of step B << 16 (offsets from `smem`; 0 = no such step)
	{
		const uint32_t j = lane & 7u, g0 = 2 * j, g1 = g0 + 1;
		const uint32_t s0 = (sel_bits >> (4 * j)) & 3u, s1 = (sel_bits >> (4 * j + 2)) & 3u;
		const uint32_t shift = version ? ctrl : 0u;
		const uint32_t i0 = version ? s0 + shift : (s0 ? s0 + 1u : 0u), i1 = version ? s1 + shift : (s1 ? s1 + 1u : 0u);
		const bool pair = g1 + 1 < groups && s0 == s1 && i0 != 0; // (g1 is not the last group)
		const uint32_t a = g0 < groups ? (pair ? kTeamSmemP - kTeamTable : kTeamSmemT) + i0 * kTeamTable : 0u;
		const uint32_t b = (g1 < groups && !pair) ? kTeamSmemT + i1 * kTeamTable : 0u;
		mine = a | (b << 16);
	}
	const uint32_t p0 = rel & (kTeamRing - 1);
	uint32_t p = p0, p_last = p0;
#pragma unroll
	for (uint32_t j = 0; j < 8; ++j)
	{
		const uint32_t dsc = __shfl_sync(0xffffffffu, mine, j);
		const uint32_t a = dsc & 0xffffu, b = dsc >> 16;
		if (a)
		{
			p_last = p;
			p += lds_u8(smem + a + p);
		}
		if (b)
		{
			p_last = p;
			p += lds_u8(smem + b + p);
		}
	}
	const uint32_t rel_last = rel + (p_last - p0);
	rel += p - p0;
	return !(rel_last > rel_end || rel_end - rel_last < kGroupReadLimit);
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
						team_need(smem, rel, 12, n_windows, have_until, released, lane);
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
					if (!team_channel(smem, rel, rel_end, groups, hdr, version, ctrl, n_windows, have_until, released, lane))
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
		// constant step tables (a zero group takes no bytes, an 8-bit group 16, a pair of them 32) and the flags
		for (uint32_t i = threadIdx.x * 16; i < kTeamTable; i += kTeamThreads * 16)
		{
			sts_v4(smem + kTeamSmemT + i, 0u, 0u, 0u, 0u);
			sts_v4(smem + kTeamSmemT + 4 * kTeamTable + i, 0x10101010u, 0x10101010u, 0x10101010u, 0x10101010u);
			sts_v4(smem + kTeamSmemP + 3 * kTeamTable + i, 0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
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
