// mob200_decoder.cuh -- phases 2+3 of the decode path: the producer warp and the decoder warps of a CTA.
//
// A unit (mob200_kernels.h) decodes the blocks  t = unit, unit + units, ...  of the level-major decode order
// (block b of every stream before block b+1 of any: the order in which the walkers publish them).
//
//   producer warp   runs ahead of the decoders by up to kSlots blocks.  Per block: ticket -> (stream, block),
//                   acquire-wait until the walker has published it, allocate a piece of the CTA's staging
//                   ring, ONE TMA bulk copy for the encoded bytes (16-byte aligned window) and one for the
//                   block's group-table rows, completion on the slot's `full` mbarrier; then the cross-block
//                   carry: block 0 starts from the tail's first vertex, later blocks from a decoupled
//                   look-back over the stream's previous blocks ({flag,value} words published by the
//                   decoders) -- handed over on the slot's `carry` mbarrier.  Every global-memory latency of
//                   the block prologue is therefore paid by a warp that has nothing else to do.
//   decoder warps   thread <-> (4-byte lane q of the vertex, chunk c of 16 vertices); the 16 chunks of a lane
//                   sit in 16 consecutive threads of one warp.  A thread unpacks its own four 16-value groups
//                   (byte-channels 4q..4q+3, group c: the work of decodeBytesGroup, reference
//                   src/vertexcodec.cpp:582-641) straight into registers, transposes them to sixteen 32-bit
//                   vertex words, undoes zigzag / rotation (decodeDeltas1, :669-699), scans its 16 vertices,
//                   and the chunk totals are scanned across the 16 threads with warp shuffles -- no shared
//                   memory and no CTA barrier up to here.  The filter runs on the finished words, the words go
//                   to a padded vertex tile in shared memory, and after the only CTA barrier of the block the
//                   tile leaves with 16-byte coalesced stores.
#pragma once

#include "mob200_device.cuh"
#include "mob200_filters.cuh"

namespace mob200
{

constexpr uint32_t kRoundBlocks = 4;           // rounds variant: blocks the decoder warps of a unit take in one round (one warp per block)
#ifndef MOB200_PLAIN_RING
#define MOB200_PLAIN_RING 14336
#endif
constexpr uint32_t kStageRingBytes = 14336;    // staging ring: encoded bytes + group-table rows of the blocks in flight (rounds form; the least any form needs)
constexpr uint32_t kPlainRingBytes = MOB200_PLAIN_RING; // plain form: 16 KB measured the same, 18 KB 2 % slower
constexpr uint32_t kRowsInRingMaxVs = 32;      // rows (32 bytes per byte-channel) travel through the ring up to this vertex size
constexpr uint32_t kRowsInGlobal = 0xffffffffu;
constexpr uint32_t kStageSlack = 16;           // bytes of the ring behind a block's encoded window that belong to the block: the unpack's 32-bit windows may
                                               // reach a few bytes past the last escape byte (never selected) and must not touch the next block's copy
constexpr uint32_t kTilePad = 8;               // bytes of padding per 16-vertex chunk of the output tile

struct BlockParams // written by the producer, read by the decoders after the slot's `full` barrier (80 bytes)
{
	uint32_t valid;
	uint32_t vs;
	uint32_t n;
	uint32_t groups;      // 16-vertex chunks of the block
	uint32_t gshift;      // log2 of the chunk count rounded up to a power of two: work item = (q << gshift) | c
	uint32_t items;       // (vs / 4) << gshift
	uint32_t stage_off;   // staging-ring offset of the block's first encoded byte
	uint32_t rows_off;    // staging-ring offset of the block's group-table rows, or kRowsInGlobal
	uint32_t first;       // block 0 of its stream
	uint32_t filter;      // enum mob200_Filter
	uint32_t filter_kind; // 0 none, 1 every 32-bit word on its own (Exp, Oct/Color on 4-byte elements), 2 8-byte elements
	uint32_t m_chunk;     // ceil(2^32 / (16 * vs)): output byte offset -> chunk
	uint8_t* out;
	const uint16_t* rows_global;
	unsigned long long* lookback; // this block's entries (vs/4 of them); predecessors lie vs/4 entries lower each
	uint32_t round_members;       // rounds variant: > 0: this block opens a decode round of so many consecutive blocks; 0: it continues one
	uint32_t chain;               // rounds variant, on a round's first block: 1 = the members are consecutive blocks of ONE stream (run-major
	                              // order): the producer resolves the carry of the first member only, the decoder warps chain the others
};

struct SlotData
{
	BlockParams P;          // 80 bytes
	uint32_t chain_total[4]; // chained round: this block's aggregate per 4-byte lane (vertex sizes <= 16), written by its decoder warp
	uint32_t carry[64];     // per 4-byte lane: value of the vertex before the block
	uint8_t channels[64];   // per 4-byte lane: channel byte (v1) or 0
};

// Shared-memory map of one unit (dynamic shared memory), for the two forms of the decode roles:
//   kRounds = false  one block at a time goes round the four decoder warps (vertex sizes that fill them: > 16 bytes);
//   kRounds = true   blocks of small vertices (one or two work quanta) are decoded up to four at a time, each in its
//                    own part of the tile; eight blocks in flight between producer and decoders.
template <bool kRounds>
struct Lay
{
	static constexpr uint32_t kSlots = kRounds ? 8 : 4; // blocks in flight between producer and decoders
	static constexpr uint32_t kTileBytes = kBlockBytes + (kRounds ? kRoundBlocks : 1) * 16 * kTilePad; // one 8 KB block, or up to four smaller ones side by side
	static constexpr uint32_t kSmemStage = 0;
	static constexpr uint32_t kRing = kRounds ? kStageRingBytes : kPlainRingBytes;
	static constexpr uint32_t kSmemTile = kSmemStage + kRing;
	static constexpr uint32_t kSmemPatch = kSmemTile + kTileBytes;            // escape-byte selector table: 16 x 4 bytes
	static constexpr uint32_t kSmemSlots = kSmemPatch + 64;
	static constexpr uint32_t kSmemBars = kSmemSlots + kSlots * sizeof(SlotData); // full[kSlots], carry[kSlots], empty[kSlots], tile_free
	static constexpr uint32_t kSmemProducer = kSmemBars + (3 * kSlots + 1) * 8;   // producer-private: ring_start[kSlots], ring_len[kSlots]
	static constexpr uint32_t kSmemPrefix = kSmemProducer + 2 * kSlots * 4;       // plain form: the unit's running value per 4-byte lane, two copies (block parity)
	static constexpr uint32_t kSmemWalker = (kSmemPrefix + 2 * 64 * 4 + 511) & ~511u; // walker warp (either form): rings, tables, barriers
	static constexpr uint32_t kSmemTotal = (kSmemWalker + 32 * 512 + 6 * 4 * 32 + 16 + 1023) & ~1023u; // one unit
	static constexpr uint32_t kSmemCta = kSmemTotal * kUnitsPerCta;
	static_assert(kSmemCta <= 227 * 1024, "shared memory of one CTA");
	static_assert((kTileBytes & 15) == 0 && (kSmemTile & 15) == 0 && (kSmemPatch & 15) == 0 && (kSmemSlots & 15) == 0 && (kSmemBars & 7) == 0, "alignment");
};
constexpr uint32_t kSmemWalkerBytes = 32 * 512 + 6 * 4 * 32 + 16; // = kWalkSmemBytes (mob200_walker.cuh) >= kWideSmemBytes

static_assert(sizeof(BlockParams) == 80 && sizeof(SlotData) == 416, "SlotData layout");
static_assert(kStageRingBytes >= kMaxEncodedBlock + 32 + kStageSlack + 32 * kRowsInRingMaxVs && kPlainRingBytes >= kStageRingBytes && kPlainRingBytes % 16 == 0, "staging ring must hold the largest block");

uint32_t decode_smem_bytes()
{
	return Lay<true>::kSmemCta > Lay<false>::kSmemCta ? Lay<true>::kSmemCta : Lay<false>::kSmemCta;
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

__device__ __forceinline__ uint32_t lane_mask(uint32_t channel)
{
	const uint32_t mode = channel & 3u;
	return mode == 0 ? 0x80808080u : (mode == 1 ? 0x80008000u : 0xffffffffu);
}

// output tile: row r (vertex) of vs bytes; every 16-row chunk is displaced by kTilePad bytes so that the
// 4-byte column writes of the 16 chunks of a lane (16 threads of one warp) fall into different banks
__device__ __forceinline__ uint32_t tile_offset(uint32_t r, uint32_t vs)
{
	return r * vs + (r >> 4) * kTilePad;
}

#ifdef MOB200_DEBUG_ENDS
__device__ unsigned long long g_dbg_spins; // diagnostics: failed mbarrier polls of the decoder warps
#endif

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
#ifdef MOB200_DEBUG_ENDS
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity))
		++spins;
	if (spins && (threadIdx.x & 31u) == 0)
		atomicAdd(&g_dbg_spins, (unsigned long long)spins);
#else
	while (!mbar_try_wait(bar, parity))
	{
	}
#endif
}

// same for waits that are expected to be long (the producer waiting for a slot: a block takes the decoders a
// few microseconds): sleep between attempts instead of spinning on issue slots the decoders need
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity, uint32_t ns)
{
	while (!mbar_try_wait(bar, parity))
		__nanosleep(ns);
}

// byte-lane helpers (channel mode 0) and 16-bit / xor helpers (modes 1, 2)
__device__ __forceinline__ uint32_t unzig8x4(uint32_t x)
{
	return ((x >> 1) & 0x7f7f7f7fu) ^ ((x & 0x01010101u) * 0xffu);
}

__device__ __forceinline__ uint32_t add8x4(uint32_t a, uint32_t b)
{
	return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}

// X = 0: two 16-bit lanes (VIADD.16x2), X = ~0: xor
__device__ __forceinline__ uint32_t combine16(uint32_t a, uint32_t b, uint32_t X)
{
	return (__vadd2(a, b) & ~X) | ((a ^ b) & X);
}

__device__ __forceinline__ uint32_t sum_bytes(const uint4& v)
{
	return __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
}

// ------------------------------------------------------------------------------------------------
// producer warp
// ------------------------------------------------------------------------------------------------

#ifndef MOB200_CARRY_LAG
#define MOB200_CARRY_LAG 1
#endif
constexpr uint32_t kCarryLag = MOB200_CARRY_LAG; // block mode: blocks the producer may stage ahead of the carry it is resolving (< slots in flight)
#ifndef MOB200_ROUND_LAG
#define MOB200_ROUND_LAG 0
#endif
constexpr uint32_t kRoundLag = MOB200_ROUND_LAG; // rounds form, block mode: rounds staged ahead of the oldest unresolved carry (0 or 1)
#ifndef MOB200_BLOCK_BATCH
#define MOB200_BLOCK_BATCH 32
#endif
constexpr uint32_t kProducerBatch = 16;// blocks whose metadata chains (ticket -> stream -> progress -> offsets) are in flight together, one per lane
constexpr uint32_t kProducerBatchBlock = MOB200_BLOCK_BATCH; // ... in block mode: the walkers are far ahead, a larger batch halves the pauses between batches
constexpr uint32_t kPrefetchBlocks = 16; // blocks of a batch whose predecessor's look-back entries are prefetched (two lanes each)
static_assert(kRoundBlocks == (1u << kRunShiftMax) && kProducerBatch % kRoundBlocks == 0, "a ticket chunk (one run of a stream) must not straddle two metadata batches");

// debug counters (cycles, summed over CTAs): see mob200_plan_debug_counters
enum
{
	kDbgDecoderTotal = 0,
	kDbgDecoderWaitFull,
	kDbgDecoderWaitCarry,
	kDbgDecoderWaitTile,
	kDbgProducerTotal,
	kDbgProducerMeta,
	kDbgProducerWaitSlot,
	kDbgProducerLookback,
};

__device__ __forceinline__ unsigned long long* debug_counters(const DevTables& T)
{
	return reinterpret_cast<unsigned long long*>(T.counters + 16);
}

// The decode order is dealt to the units in chunks of 1 << ticket_shift tickets, chunk c to unit c % units.
__device__ __forceinline__ uint32_t unit_block_count(const DevTables& T, uint32_t unit)
{
	const uint32_t sh = T.ticket_shift;
	const uint32_t chunks = (T.total_blocks + (1u << sh) - 1u) >> sh;
	if (unit >= chunks)
		return 0u;
	const uint32_t mine = (chunks - unit + T.units - 1) / T.units;
	const uint32_t last = unit + (mine - 1) * T.units; // the last chunk of the order may be short
	return (mine << sh) - (last == chunks - 1 ? (chunks << sh) - T.total_blocks : 0u);
}

__device__ __forceinline__ uint32_t unit_ticket(const DevTables& T, uint32_t unit, uint32_t i)
{
	const uint32_t sh = T.ticket_shift;
	return ((((i >> sh) * T.units) + unit) << sh) + (i & ((1u << sh) - 1u));
}

template <bool kRounds, bool kBlock>
__device__ void producer_main(const DevTables& T, uint8_t* smem, const uint32_t unit)
{
	using L = Lay<kRounds>;
	constexpr uint32_t kSlots = L::kSlots;
	const uint32_t lane = threadIdx.x & 31u;
	uint8_t* ring = smem + L::kSmemStage;
	SlotData* slots = reinterpret_cast<SlotData*>(smem + L::kSmemSlots);
	uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kSmemBars);
	uint64_t* carry_bar = full + kSlots;
	uint64_t* empty = carry_bar + kSlots;
	uint32_t* ring_start = reinterpret_cast<uint32_t*>(smem + L::kSmemProducer);
	uint32_t* ring_len = ring_start + kSlots;

	uint32_t head = 0;  // next free byte of the staging ring
	uint32_t freed = 0; // blocks [freed, i) of this CTA's sequence are in flight (their ring pieces are live)
	const uint32_t my_count = unit_block_count(T, unit);
	uint32_t prev_s = 0xffffffffu, prev_b = 0, prev_valid = 0; // the unit's previous block (unit-local chains)

	long long dbg_meta = 0, dbg_slot = 0, dbg_look = 0;
	const long long dbg_t0 = dbg_clock();

	constexpr uint32_t kBatch = kBlock ? kProducerBatchBlock : kProducerBatch;
	static_assert(kBatch <= 32 && kBatch % kRoundBlocks == 0, "one lane per block of a batch");
	for (uint32_t i0 = 0; i0 < my_count; i0 += kBatch)
	{
		// ---- metadata of up to kProducerBatch blocks, one per lane: the dependent global loads of all of them overlap ----
		const long long c0 = dbg_clock();
		MOB200_TRACE_EVENT(T, unit, lane, 1, i0);
		// level 1: ticket -> (stream, block); level 2: stream descriptor + walker progress (which also carries the codec
		// version); level 3: block byte range, channel bytes, and -- four entries per lane, two lanes per block -- the
		// look-back entries of the predecessor block.  Everything of one level is requested before any of it is used.
		const uint32_t mi = i0 + lane;
		const bool has = lane < kBatch && mi < my_count;
		uint32_t m_valid = 0, m_vs = 4, m_n = 0, m_filter = 0, m_version = 0, m_b = 0, m_enc = 0, m_shift = 0, m_ready = 0;
		uint32_t m_s = 0xffffffffu; // stream of the block (chained rounds; unit-local chains of the plain form)
		uint32_t m_chain = 0;       // plain form, block mode: the unit's previous block is this block's predecessor and decodable
		uint32_t m_magic = 0;       // ceil(2^32 / (16 * vertex size)): a division, done here for all blocks of the batch at once
		uint32_t m_quanta = 0, m_len = 0; // rounds variant: decoder work quanta (32 items each) and staged bytes of the block
		unsigned long long m_lo = 0, m_tail = 0, m_out = 0, m_rows = 0, m_look = 0;
		const uint32_t* boff = nullptr;
		const uint8_t* src = nullptr;
		uint32_t src_size = 0;
		if (has)
		{
			const uint32_t t = unit_ticket(T, unit, mi);
			const uint2 info = __ldg(T.ticket_info + t);
			const uint32_t s = info.x, b = info.y;
			const DevStream* d = T.streams + s;
			m_s = s;
			src = d->src;
			src_size = d->src_size;
			const uint32_t vs = d->vertex_size;
			m_vs = vs;
			m_magic = magic_for(16 * vs);
			m_b = b;
			m_filter = d->filter;
			const uint32_t bv = d->block_groups * kGroup;
			m_n = min(bv, d->vertex_count - b * bv);
			m_out = reinterpret_cast<unsigned long long>(d->dst + (uint64_t)b * bv * vs);
			m_rows = reinterpret_cast<unsigned long long>(T.group_table + (d->chan_base + (uint64_t)b * vs) * 16);
			m_look = reinterpret_cast<unsigned long long>(T.lookback + (d->chan_base >> 2) + (uint64_t)b * (vs >> 2));
			boff = T.block_offset + d->block_base + s + b;

			// wait until a walker has published this block (back-off: a starved producer must not take issue
			// slots from the walker warps)
			const uint32_t* ready = T.block_ready + d->block_base + b;
			uint32_t rv = ld_acquire_u32(ready);
			for (uint32_t ns = 32; (rv >> 2) != T.epoch;)
			{
				__nanosleep(ns);
				ns = ns < 1024 ? ns * 2 : ns;
				rv = ld_acquire_u32(ready);
			}
			m_version = (rv >> 1) & 1u;
			m_ready = rv & 1u;
		}
		__syncwarp();

		unsigned long long pre0 = 0, pre1 = 0, pre2 = 0, pre3 = 0;
		{
			const uint32_t tj = lane >> 1, q0 = (lane & 1u) * 4u;
			const uint32_t pvs = __shfl_sync(0xffffffffu, m_vs, tj), pb = __shfl_sync(0xffffffffu, m_b, tj);
			const unsigned long long* plook = reinterpret_cast<const unsigned long long*>(__shfl_sync(0xffffffffu, m_look, tj));
			const uint32_t pnq = pvs >> 2;
			if (plook && pb > 0 && pnq <= 8)
			{
				// (a block that turns out not to be decodable has a decodable predecessor or none: the entries exist)
				const unsigned long long* e = plook + q0 - pnq;
				if (q0 < pnq)
					pre0 = ld_volatile_u64(e);
				if (q0 + 1 < pnq)
					pre1 = ld_volatile_u64(e + 1);
				if (q0 + 2 < pnq)
					pre2 = ld_volatile_u64(e + 2);
				if (q0 + 3 < pnq)
					pre3 = ld_volatile_u64(e + 3);
			}
		}
		uint32_t m_ch_lo = 0, m_ch_hi = 0;
		if (has)
		{
			const uint32_t off = __ldcg(boff), end = __ldcg(boff + 1);
			const uint32_t vs = m_vs;
			const uint8_t* tail = src + src_size - tail_bytes(vs, m_version);
			if (m_version && vs <= 32 && src_size >= tail_bytes(vs, m_version))
			{
				// channel bytes (up to 8); harmless for a stream whose walk failed: the tail lies inside the input
				const uint8_t* ch = tail + vs;
				const uint32_t nqj = vs >> 2;
				m_ch_lo = ldg_u32_at(ch, min(nqj, 4u)) & (nqj >= 4 ? 0xffffffffu : ((1u << (8 * nqj)) - 1u));
				if (nqj > 4)
					m_ch_hi = ldg_u32_at(ch + 4, nqj - 4) & (nqj >= 8 ? 0xffffffffu : ((1u << (8 * (nqj - 4))) - 1u));
			}
			if (m_ready && off != kInvalidOffset && end != kInvalidOffset)
			{
				m_valid = 1;
				m_tail = reinterpret_cast<unsigned long long>(tail);
				// 16-byte aligned window around [off, end)
				const uintptr_t a0 = reinterpret_cast<uintptr_t>(src) + off;
				const uintptr_t a1 = reinterpret_cast<uintptr_t>(src) + end;
				const uintptr_t lo = a0 & ~uintptr_t(15);
				const uintptr_t hi = (a1 + 15) & ~uintptr_t(15);
				m_lo = lo;
				m_enc = (uint32_t)(hi - lo);
				m_shift = (uint32_t)(a0 - lo);
				m_len = m_enc + kStageSlack + (vs <= kRowsInRingMaxVs ? 32u * vs : 0u);
				if (kRounds)
				{
					const uint32_t groups_j = (m_n + kGroup - 1) / kGroup;
					const uint32_t gshift_j = groups_j > 8 ? 4u : (groups_j > 4 ? 3u : (groups_j > 2 ? 2u : (groups_j > 1 ? 1u : 0u)));
					m_quanta = (((vs >> 2) << gshift_j) + 31u) >> 5;
				}
			}
		}
		__syncwarp();
		if (!kRounds && kBlock)
		{
			// (lane - 1 holds the unit's previous block; for lane 0 it is the last block of the previous batch)
			const uint32_t ps = __shfl_up_sync(0xffffffffu, m_s, 1), pb = __shfl_up_sync(0xffffffffu, m_b, 1), pv = __shfl_up_sync(0xffffffffu, m_valid, 1);
			const uint32_t qs = lane ? ps : prev_s, qb = lane ? pb : prev_b, qv = lane ? pv : prev_valid;
			m_chain = (has && m_valid && qv && qs == m_s && qb + 1 == m_b && T.ticket_shift) ? 1u : 0u;
			const uint32_t lastl = min(kBatch, my_count - i0) - 1;
			prev_s = __shfl_sync(0xffffffffu, m_s, lastl), prev_b = __shfl_sync(0xffffffffu, m_b, lastl), prev_valid = __shfl_sync(0xffffffffu, m_valid, lastl);
		}
		dbg_meta += dbg_clock() - c0;
		MOB200_TRACE_EVENT(T, unit, lane, 2, i0);

		// ---- hand the blocks to the decoders, in order -------------------------------------------------------------
		const uint32_t in_batch = min(kBatch, my_count - i0);
		// Rounds variant: the members of a round are all staged (slot, ring piece, TMA) before the carry of any of them is
		// resolved -- the decoders start a round when every member has landed, and a carry may depend on a block that
		// another unit decodes in a round of the same age, so a copy that waited for a carry could close a cycle.
		// Plain form, block mode: the carry of block j is resolved after the blocks up to j + kCarryLag have been staged (a
		// carry may have to wait for blocks that other units are still unpacking -- with one stream the predecessor is the
		// ticket right before this one -- and the decoders should find their next blocks staged meanwhile).
		// Rounds form, block mode: the same one round ahead (kRoundLag) -- round N is staged while the carries of the round
		// before it (A) are still open, if its ring pieces can be had without waiting for A's.
		uint32_t g = 0, mode = 0;   // rounds form: 0 decide, 1 staging round N, 2 resolving the carries of round A
		uint32_t sj = 0;            // rounds form: next block of the batch to stage
		uint32_t n_j0 = 0, n_members = 1, a_j0 = 0, a_members = 0;
		bool n_chained = false, n_wait = false, a_chained = false, a_have = false;
		uint32_t js = 0, jc = 0;    // plain form: next block to stage / to resolve the carry of
		for (;;)
		{
			if (!kRounds && jc >= in_batch)
				break;
			bool plain_stage = !kRounds && js < in_batch && js == jc;
			if (!kRounds && kBlock && js < in_batch && js > jc && js <= jc + kCarryLag)
			{
				// up to kCarryLag blocks ahead of the carry -- but only if the ring piece of block js can be had without waiting
				// for a block whose carry this warp has not resolved yet (its decoders wait for that carry): the pieces of
				// the unresolved blocks jc .. js-1 are the ones in front of `head`; with everything older released, the
				// allocator below takes `head` if the piece fits behind it, else offset 0 if it fits in front of the oldest
				const uint32_t lens = __shfl_sync(0xffffffffu, m_len, js);
				uint32_t so = 0xffffffffu;
				for (uint32_t u = jc; u < js && so == 0xffffffffu; ++u)
					if (ring_len[(i0 + u) & (kSlots - 1)])
						so = ring_start[(i0 + u) & (kSlots - 1)];
				plain_stage = so == 0xffffffffu || (head > so ? (head + lens <= L::kRing || lens < so) : head + lens < so);
			}
			if (kRounds && mode == 0)
			{
				g = 0;
				if (sj < in_batch && !n_wait)
				{
					// a block of at most two work quanta is joined by the following blocks of this batch as long as the round
					// stays within the four decoder warps and within half of the staging ring (all members are staged together)
					const uint32_t j0 = sj;
					uint32_t qs = __shfl_sync(0xffffffffu, m_quanta, j0);
					uint32_t members = 1;
					if (qs <= 2)
					{
						uint32_t bs = __shfl_sync(0xffffffffu, m_len, j0);
						bool open = true;
#pragma unroll 1
						for (uint32_t k = 1; k < kRoundBlocks && open && j0 + k < in_batch; ++k)
						{
							const uint32_t qn = __shfl_sync(0xffffffffu, m_quanta, (j0 + k) & 31u), bn = __shfl_sync(0xffffffffu, m_len, (j0 + k) & 31u);
							if (qs + qn <= kDecodeThreads / 32 && bs + bn <= L::kRing / 2)
							{
								qs += qn, bs += bn;
								members = k + 1;
								open = qs < kDecodeThreads / 32;
							}
							else
								open = false;
						}
					}
					// run-major order: members that are consecutive decodable blocks of one stream (<= 16-byte vertices) chain their
					// carries among the decoder warps; only the first member's carry comes from the look-back
					bool chained = false;
					if (kBlock && T.ticket_shift && members > 1)
					{
						const uint32_t s0 = __shfl_sync(0xffffffffu, m_s, j0), b0 = __shfl_sync(0xffffffffu, m_b, j0);
						chained = __shfl_sync(0xffffffffu, m_valid, j0) != 0 && __shfl_sync(0xffffffffu, m_vs, j0) <= 16;
						for (uint32_t k = 1; k < members; ++k)
							chained = chained && __shfl_sync(0xffffffffu, m_valid, (j0 + k) & 31u) != 0 && __shfl_sync(0xffffffffu, m_s, (j0 + k) & 31u) == s0 &&
							          __shfl_sync(0xffffffffu, m_b, (j0 + k) & 31u) == b0 + k;
					}
					n_j0 = j0, n_members = members, n_chained = chained;

					bool ahead = !a_have;
					if (kBlock && kRoundLag && a_have)
					{
						// Round A's decoders wait for carries this warp has not resolved: staging N must not wait for A's ring
						// pieces.  Everything older than A finishes on its own: wait for it, then A's pieces are the only live
						// ones and the allocator's choices for N's pieces can be played through exactly.
						const uint32_t a_i0 = i0 + a_j0;
						while (freed < a_i0)
						{
							mbar_wait_long(empty + (freed & (kSlots - 1)), (freed / kSlots) & 1u, 400);
							++freed;
						}
						uint32_t so = 0xffffffffu;
						for (uint32_t u = 0; u < a_members && so == 0xffffffffu; ++u)
							if (ring_len[(a_i0 + u) & (kSlots - 1)])
								so = ring_start[(a_i0 + u) & (kSlots - 1)];
						ahead = true;
						uint32_t h = head;
						for (uint32_t k = 0; k < members && ahead && so != 0xffffffffu; ++k)
						{
							const uint32_t len = __shfl_sync(0xffffffffu, m_len, (j0 + k) & 31u);
							if (len == 0)
								continue;
							if (h > so)
							{
								if (h + len <= L::kRing)
									h += len;
								else if (len < so)
									h = len;
								else
									ahead = false;
							}
							else if (h + len < so)
								h += len;
							else
								ahead = false;
						}
					}
					mode = ahead ? 1u : 2u;
				}
				else if (a_have)
					mode = 2;
				else
					break;
			}
			const uint32_t members = mode == 1 ? n_members : a_members;
			const bool chained = mode == 1 ? n_chained : a_chained;
			const uint32_t j = kRounds ? (mode == 1 ? n_j0 : a_j0) + g : (plain_stage ? js : jc);
			const bool do_stage = kRounds ? mode == 1 : plain_stage, do_carry = kRounds ? mode == 2 : !plain_stage;
			const uint32_t i = i0 + j;
			const uint32_t slot = i & (kSlots - 1);
			SlotData& S = slots[slot];
			const bool valid = __shfl_sync(0xffffffffu, m_valid, j) != 0;
			const uint32_t vs = __shfl_sync(0xffffffffu, m_vs, j);
			const uint32_t nq = vs >> 2;
			const uint32_t b = __shfl_sync(0xffffffffu, m_b, j);

			if (do_stage)
			{
				// Blocks j .. j + cnt - 1 are staged in ONE pass, each by the lane that holds its metadata (rounds form: all
				// members of the round; the producer's code runs once per round, not once per block -- its instructions are
				// cold in the instruction cache every time, which is what a small-vertex round used to spend its time on).
				const uint32_t cnt = kRounds ? members : 1u;
				const bool owner = lane >= j && lane < j + cnt;
				const uint32_t i_l = i0 + lane, slot_l = i_l & (kSlots - 1);
				SlotData& Sl = slots[slot_l];
				const long long c1 = dbg_clock();
				MOB200_TRACE_EVENT(T, unit, lane, 3, i);
				// a slot's previous block (i - kSlots) must have been unpacked by every decoder warp
				if (owner)
					mbar_wait_long(empty + slot_l, ((i_l / kSlots) & 1u) ^ 1u, 400);
				__syncwarp();
				MOB200_TRACE_EVENT(T, unit, lane, 7, i);
				const uint32_t i_last = i + cnt - 1;
				if (i_last >= kSlots && freed < i_last - (kSlots - 1))
					freed = i_last - (kSlots - 1);

				// ring bytes of the pass (an undecodable block takes none) and of the blocks in front of this lane's
				uint32_t total = 0, before = 0;
				for (uint32_t k = 0; k < cnt; ++k)
				{
					const uint32_t l = __shfl_sync(0xffffffffu, m_len, (j + k) & 31u);
					before += j + k < lane ? l : 0u;
					total += l;
				}
				uint32_t start = 0;
				if (total)
				{
					// allocate `total` contiguous bytes of the ring (pieces are freed in order)
					for (;;)
					{
						uint32_t f = freed; // oldest live piece
						while (f < i && ring_len[f & (kSlots - 1)] == 0)
							++f;
						if (f == i)
						{
							start = 0;
							break;
						}
						const uint32_t so = ring_start[f & (kSlots - 1)];
						if (head > so)
						{
							if (head + total <= L::kRing)
							{
								start = head;
								break;
							}
							if (total < so)
							{
								start = 0;
								break;
							}
						}
						else if (head + total < so)
						{
							start = head;
							break;
						}
						mbar_wait_long(empty + (freed & (kSlots - 1)), (freed / kSlots) & 1u, 400);
						++freed;
					}
					head = start + total;
				}
				dbg_slot += dbg_clock() - c1;
				__syncwarp();
				if (owner)
				{
					const uint32_t vs_l = m_vs, nq_l = m_vs >> 2;
					BlockParams& P = Sl.P;
					if (kRounds)
					{
						P.round_members = lane == j ? cnt : 0u;
						P.chain = (lane == j && chained) ? 1u : 0u;
					}
					else
						P.chain = m_chain;
					if (m_valid)
					{
						const uint32_t rows_bytes = vs_l <= kRowsInRingMaxVs ? 32u * vs_l : 0u;
						const uint32_t at = start + before;
						// channel bytes: needed by the decoders from the start of the block
						if (nq_l <= 8)
							*reinterpret_cast<uint2*>(Sl.channels) = make_uint2(m_ch_lo, m_ch_hi);
						else
						{
							const uint8_t* ch = reinterpret_cast<const uint8_t*>(m_tail) + vs_l;
							for (uint32_t q = 0; q < nq_l; ++q)
								Sl.channels[q] = m_version ? __ldg(ch + q) : (uint8_t)0;
						}
						ring_start[slot_l] = at;
						ring_len[slot_l] = m_len;
						const uint32_t groups = (m_n + kGroup - 1) / kGroup;
						const uint32_t gshift = groups > 8 ? 4u : (groups > 4 ? 3u : (groups > 2 ? 2u : (groups > 1 ? 1u : 0u)));
						P.valid = 1;
						P.vs = vs_l;
						P.n = m_n;
						P.groups = groups;
						P.gshift = gshift;
						P.items = nq_l << gshift;
						P.stage_off = at + m_shift;
						P.rows_off = rows_bytes ? at + m_enc + kStageSlack : kRowsInGlobal;
						P.first = m_b == 0;
						P.filter = m_filter;
						P.filter_kind = m_filter == MOB200_FILTER_NONE ? 0u : ((m_filter == MOB200_FILTER_EXP || vs_l == 4) ? 1u : 2u);
						P.m_chunk = m_magic;
						P.out = reinterpret_cast<uint8_t*>(m_out);
						P.rows_global = reinterpret_cast<const uint16_t*>(m_rows);
						P.lookback = reinterpret_cast<unsigned long long*>(m_look);

						fence_proxy_async(); // the decoders' generic-proxy reads of the reused ring bytes are ordered before the copies
						mbar_expect_tx(full + slot_l, m_enc + rows_bytes);
						tma_load_bulk(ring + at, reinterpret_cast<const void*>(m_lo), m_enc, full + slot_l);
						if (rows_bytes)
							tma_load_bulk(ring + at + m_enc + kStageSlack, P.rows_global, rows_bytes, full + slot_l);
					}
					else
					{
						ring_len[slot_l] = 0;
						P.valid = 0;
						mbar_arrive(full + slot_l);
						if (kBlock)
						{
							// block mode: later blocks of the stream may still be decodable (their output is garbage, as the
							// reference allows for a rejected stream) and must not wait for this block's look-back entries
							unsigned long long* look = reinterpret_cast<unsigned long long*>(m_look);
							for (uint32_t q = 0; q < nq_l; ++q)
								st_volatile_u64(look + q, ((unsigned long long)((T.epoch << 2) | 2u)) << 32);
						}
					}
				}
				__syncwarp();
				g += cnt - 1; // (rounds form: the pass covered the whole round)

				MOB200_TRACE_EVENT(T, unit, lane, 4, i);
			} // do_stage

			// ---- carry into the block, per 4-byte lane ---------------------------------------------------------------
			if (do_carry)
			{
				const long long c2 = dbg_clock();
				MOB200_TRACE_EVENT(T, unit, lane, 5, i);
				bool carry_done = !kRounds && kBlock && __shfl_sync(0xffffffffu, m_chain, j) != 0; // the decoders hand the value on
				if (!carry_done && valid && b > 0 && nq <= 8 && j < kPrefetchBlocks)
				{
					// the prefetched look-back entries: good if every one of them is an inclusive prefix of this run
					const bool mine = (lane >> 1) == j;
					const uint32_t q0 = (lane & 1u) * 4u;
					const uint32_t want_flag = ((T.epoch & 0x3fffffffu) << 2) | 2u;
					const bool good = (q0 >= nq || (uint32_t)(pre0 >> 32) == want_flag) && (q0 + 1 >= nq || (uint32_t)(pre1 >> 32) == want_flag) &&
					                  (q0 + 2 >= nq || (uint32_t)(pre2 >> 32) == want_flag) && (q0 + 3 >= nq || (uint32_t)(pre3 >> 32) == want_flag);
					carry_done = __all_sync(0xffffffffu, !mine || good);
					if (carry_done && mine)
					{
						if (q0 < nq)
							S.carry[q0] = (uint32_t)pre0;
						if (q0 + 1 < nq)
							S.carry[q0 + 1] = (uint32_t)pre1;
						if (q0 + 2 < nq)
							S.carry[q0 + 2] = (uint32_t)pre2;
						if (q0 + 3 < nq)
							S.carry[q0 + 3] = (uint32_t)pre3;
					}
				}
				if (!kRounds && kBlock && valid && !carry_done && b > 0 && nq <= 32)
				{
					// (plain form; a round's predecessor is usually still in flight, the extra round trip measured 8 % slower there)
					// the usual case when the predecessor lies a generation of blocks back: it has published its inclusive
					// prefix -- one load per 4-byte lane, none of the look-back loop's instructions fetched
					const unsigned long long* look1 = reinterpret_cast<const unsigned long long*>(__shfl_sync(0xffffffffu, m_look, j));
					const unsigned long long e = lane < nq ? ld_volatile_u64(look1 + lane - nq) : 0ull;
					carry_done = __all_sync(0xffffffffu, lane >= nq || (uint32_t)(e >> 32) == (((T.epoch & 0x3fffffffu) << 2) | 2u));
					if (carry_done && lane < nq)
						S.carry[lane] = (uint32_t)e;
				}
				if (kBlock && valid && !carry_done && b > 0 && nq <= 32)
				{
					// Decoupled look-back, 2 * 32 / nqp predecessors per step (nqp = nq rounded up to a power of two): lane =
					// (predecessor pj, 4-byte lane q).  Per q the lanes consume the leading run of published predecessors up
					// to the first inclusive prefix (state 2), combine their values with a butterfly over pj, and go on
					// behind that run until a prefix has been met.  (Block 0 of a stream only ever publishes a prefix, so
					// the walk never steps below it.)
					const unsigned long long* look = reinterpret_cast<const unsigned long long*>(__shfl_sync(0xffffffffu, m_look, j));
					const uint32_t nqp = nq <= 1 ? 1u : (nq <= 2 ? 2u : (nq <= 4 ? 4u : (nq <= 8 ? 8u : (nq <= 16 ? 16u : 32u))));
					const uint32_t q = lane & (nqp - 1u), pj = lane / nqp;
					// bit pj * nqp for every predecessor slot of a step
					const uint32_t pattern = nqp == 1 ? 0xffffffffu : (nqp == 2 ? 0x55555555u : (nqp == 4 ? 0x11111111u : (nqp == 8 ? 0x01010101u : (nqp == 16 ? 0x00010001u : 1u))));
					const bool qon = q < nq;
					const uint32_t Hq = lane_mask(qon ? S.channels[q] : 0u);
					const uint32_t want_epoch = T.epoch & 0x3fffffffu;
					// every lane looks at TWO consecutive predecessors per step (one when nqp == 1: 32 lanes = 32 predecessors
					// already): predecessor slot 2 * pj + h sits at bit pj * nqp + h * nqp / 2 of the masks, i.e. in order of distance
					const bool two = nqp >= 2;
					const uint32_t half = nqp >> 1, per = two ? 2u : 1u;
					const uint32_t slots = two ? (pattern | (pattern << half)) : pattern;
					bool done = !qon;
					uint32_t acc = 0, dist = 1;
					while (!__all_sync(0xffffffffu, done))
					{
						const uint32_t d0 = dist + per * pj, d1 = d0 + 1u;
						const bool here0 = !done && d0 <= b, here1 = two && !done && d1 <= b;
						unsigned long long e0 = 0, e1 = 0;
						if (here0)
							e0 = ld_volatile_u64(look + q - (size_t)d0 * nq);
						if (here1)
							e1 = ld_volatile_u64(look + q - (size_t)d1 * nq);
						const uint32_t f0 = (uint32_t)(e0 >> 32), f1 = (uint32_t)(e1 >> 32);
						// (beyond block 0: "prefix 0", never reached because block 0 itself is a prefix)
						const uint32_t s0 = here0 ? ((f0 >> 2) == want_epoch ? (f0 & 3u) : 0u) : 2u;
						const uint32_t s1 = here1 ? ((f1 >> 2) == want_epoch ? (f1 & 3u) : 0u) : 2u;
						uint32_t rmask = (__ballot_sync(0xffffffffu, s0 != 0) >> q) & pattern;
						uint32_t pmask = (__ballot_sync(0xffffffffu, s0 == 2) >> q) & pattern;
						if (two)
						{
							rmask |= ((__ballot_sync(0xffffffffu, s1 != 0) >> q) & pattern) << half;
							pmask |= ((__ballot_sync(0xffffffffu, s1 == 2) >> q) & pattern) << half;
						}
						const uint32_t nr = ~rmask & slots;                                // predecessors that have not published yet
						const uint32_t below = nr ? ((nr & (0u - nr)) - 1u) : 0xffffffffu; // ... and everything nearer than the first of them
						const uint32_t fp = (pmask & below) & (0u - (pmask & below));      // nearest inclusive prefix inside that run
						uint32_t consume = below & slots;
						if (fp)
							consume &= (fp << 1) - 1u;
						uint32_t v = (!done && ((consume >> (pj * nqp)) & 1u)) ? (uint32_t)e0 : 0u;
						if (two && !done && ((consume >> (pj * nqp + half)) & 1u))
							v = lane_combine(v, (uint32_t)e1, Hq);
						for (uint32_t st = nqp; st < 32; st <<= 1)
							v = lane_combine(v, __shfl_xor_sync(0xffffffffu, v, st), Hq);
						if (!done)
						{
							acc = lane_combine(acc, v, Hq);
							if (fp)
								done = true;
							else
								dist += __popc(consume);
#ifndef MOB200_LOOKBACK_SLEEP
#define MOB200_LOOKBACK_SLEEP 64
#endif
							if (consume == 0)
								__nanosleep(MOB200_LOOKBACK_SLEEP);
						}
					}
					if (qon && pj == 0)
						S.carry[q] = acc;
				}
				else if (valid && !carry_done)
				{
					const unsigned long long* look = reinterpret_cast<const unsigned long long*>(__shfl_sync(0xffffffffu, m_look, j));
					const uint8_t* tail = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, m_tail, j));
					for (uint32_t q = lane; q < nq; q += 32)
					{
						const uint32_t channel = S.channels[q];
						uint32_t prefix;
						if (b == 0)
						{
							// first vertex stored in the tail (:1846-1849)
							const uint8_t* fv = tail + q * 4;
							prefix = (uint32_t)__ldg(fv) | ((uint32_t)__ldg(fv + 1) << 8) | ((uint32_t)__ldg(fv + 2) << 16) | ((uint32_t)__ldg(fv + 3) << 24);
						}
						else
						{
							// decoupled look-back: add block aggregates (state 1) until an inclusive prefix (state 2)
							const uint32_t Hq = lane_mask(channel);
							prefix = 0;
							const unsigned long long* prev = look + q - nq;
							for (;;)
							{
								const unsigned long long e = ld_volatile_u64(prev);
								const uint32_t flag = (uint32_t)(e >> 32);
								if ((flag >> 2) != (T.epoch & 0x3fffffffu) || (flag & 3u) == 0)
									continue; // not published yet in this run
								prefix = lane_combine(prefix, (uint32_t)e, Hq);
								if ((flag & 3u) == 2)
									break;
								prev -= nq;
							}
						}
						S.carry[q] = prefix;
					}
				}
				// every lane releases its own carry words (the barrier counts the 32 producer lanes)
				mbar_arrive(carry_bar + slot);
				if (kRounds && chained)
				{
					// chained round: the decoder warps hand the carry on from member to member; the other members' barriers
					// only have to complete their phase
					for (uint32_t k = 1; k < members; ++k)
						mbar_arrive(carry_bar + ((i + k) & (kSlots - 1)));
					g = members - 1;
				}
				__syncwarp();
				dbg_look += dbg_clock() - c2;
				MOB200_TRACE_EVENT(T, unit, lane, 6, i);
			} // do_carry

			if (!kRounds)
			{
				if (plain_stage)
					++js;
				else
					++jc;
			}
			else if (++g == members)
			{
				if (mode == 1)
				{
					sj += n_members;
					if (a_have)
						n_wait = true, mode = 2, g = 0; // N is staged: now the carries of A
					else
						a_j0 = n_j0, a_members = n_members, a_chained = n_chained, a_have = true, mode = 0;
				}
				else
				{
					a_have = false;
					if (n_wait)
						a_j0 = n_j0, a_members = n_members, a_chained = n_chained, a_have = true, n_wait = false;
					mode = 0;
				}
			}
		}
	}

#ifdef MOB200_DEBUG_COUNTERS
	if (lane == 0)
	{
		unsigned long long* dbg = debug_counters(T);
		atomicAdd(dbg + kDbgProducerTotal, (unsigned long long)(dbg_clock() - dbg_t0));
		atomicAdd(dbg + kDbgProducerMeta, (unsigned long long)dbg_meta);
		atomicAdd(dbg + kDbgProducerWaitSlot, (unsigned long long)dbg_slot);
		atomicAdd(dbg + kDbgProducerLookback, (unsigned long long)dbg_look);
	}
#endif
}

// ------------------------------------------------------------------------------------------------
// decoder warps
// ------------------------------------------------------------------------------------------------

// one 16-value group -> 16 bytes in registers.  entry = 0: all zero; else (offset << 2) | log2(bits).
//
// Fields equal to the all-ones value are replaced, in order, by the escape bytes that follow the packed fields
// (reference src/vertexcodec.cpp:582-641).  The replacement is branch-free and works on the expanded bytes, four
// values at a time: bit 7 of (value + 0x80 - sentinel) flags a sentinel (values never exceed the sentinel), one
// multiply gathers the four flags into an index, the escape bytes consumed by earlier words are a popcount, and a
// 16-entry table turns the index into the PRMT selector that merges the word with the next escape bytes.
constexpr uint32_t kFlagGather = 0x02040810u; // (flags at bits 7/15/23/31) * this -> bits 32..35 of the product

__device__ __forceinline__ uint32_t sentinel_index(uint32_t v, uint32_t bias)
{
	return __umulhi((v + bias) & 0x80808080u, kFlagGather) & 15u;
}

// selector for a word whose bytes with a set index bit take the next escape bytes (operand b of PRMT), in order
__device__ __forceinline__ uint32_t patch_selector(uint32_t idx)
{
	uint32_t sel = 0, rank = 0;
	for (uint32_t k = 0; k < 4; ++k)
	{
		const bool hit = (idx >> k) & 1u;
		sel |= (hit ? 4u + rank : k) << (4 * k);
		rank += hit;
	}
	return sel;
}

#ifdef MOB200_SHARED_PATCH
// (variant: ONE copy of the escape-byte merge for the four unpack copies of a work item)
__device__ __noinline__ uint4 patch_group(const uint8_t* ring, uint4 r, uint32_t esc, uint32_t idx, const uint32_t* patch_lut)
{
	const uint32_t i0 = idx & 15u, i1 = (idx >> 4) & 15u, i2 = (idx >> 8) & 15u, i3 = idx >> 12;
	const uint32_t e1 = esc + __popc(i0), e2 = e1 + __popc(i1), e3 = e2 + __popc(i2);
	r.x = __byte_perm(r.x, lds_u32_at(ring, esc), patch_lut[i0]);
	r.y = __byte_perm(r.y, lds_u32_at(ring, e1), patch_lut[i1]);
	r.z = __byte_perm(r.z, lds_u32_at(ring, e2), patch_lut[i2]);
	r.w = __byte_perm(r.w, lds_u32_at(ring, e3), patch_lut[i3]);
	return r;
}
#endif

// (inlined: the four groups of a work item interleave; as a call the fused kernel is 4% slower, the decoders alone 9%)
__device__ __forceinline__ uint4 unpack_group(
    const uint8_t* ring, uint32_t base, uint32_t entry, const uint32_t* patch_lut)
{
	uint4 r = make_uint4(0, 0, 0, 0);
	if (entry == 0)
		return r;
	const uint32_t o = base + (entry >> 2);
	const uint32_t code = entry & 3u;
	uint32_t bias, esc; // 0x80 - sentinel in every byte; offset of the first escape byte

	if (code == 3)
	{
		const uint32_t* w = reinterpret_cast<const uint32_t*>(ring) + (o >> 2);
		const uint32_t s8 = (o & 3u) * 8u;
		const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
		r.x = __funnelshift_r(a0, a1, s8);
		r.y = __funnelshift_r(a1, a2, s8);
		r.z = __funnelshift_r(a2, a3, s8);
		r.w = __funnelshift_r(a3, a4, s8);
		return r;
	}
	if (code == 2)
	{
		const uint32_t* w = reinterpret_cast<const uint32_t*>(ring) + (o >> 2);
		const uint32_t s8 = (o & 3u) * 8u;
		const uint32_t a0 = w[0], a1 = w[1], a2 = w[2];
		const uint32_t x0 = __funnelshift_r(a0, a1, s8), x1 = __funnelshift_r(a1, a2, s8);
		const uint32_t h0 = (x0 >> 4) & 0x0f0f0f0fu, l0 = x0 & 0x0f0f0f0fu;
		const uint32_t h1 = (x1 >> 4) & 0x0f0f0f0fu, l1 = x1 & 0x0f0f0f0fu;
		r.x = __byte_perm(h0, l0, 0x5140);
		r.y = __byte_perm(h0, l0, 0x7362);
		r.z = __byte_perm(h1, l1, 0x5140);
		r.w = __byte_perm(h1, l1, 0x7362);
		bias = 0x71717171u;
		esc = o + 8;
	}
	else if (code == 1)
	{
		const uint32_t x = lds_u32_at(ring, o);
		const uint32_t b0 = x & 0xff, b1 = (x >> 8) & 0xff, b2 = (x >> 16) & 0xff, b3 = x >> 24;
		r.x = ((b0 * 0x01004010u) & 0x03030300u) | (b0 >> 6);
		r.y = ((b1 * 0x01004010u) & 0x03030300u) | (b1 >> 6);
		r.z = ((b2 * 0x01004010u) & 0x03030300u) | (b2 >> 6);
		r.w = ((b3 * 0x01004010u) & 0x03030300u) | (b3 >> 6);
		bias = 0x7d7d7d7du;
		esc = o + 4;
	}
	else
	{
		const uint32_t x = lds_u32_at(ring, o) & 0xffffu; // bit i = value i
		r.x = ((x & 15u) * 0x00204081u) & 0x01010101u;
		r.y = (((x >> 4) & 15u) * 0x00204081u) & 0x01010101u;
		r.z = (((x >> 8) & 15u) * 0x00204081u) & 0x01010101u;
		r.w = ((x >> 12) * 0x00204081u) & 0x01010101u;
		bias = 0x7f7f7f7fu;
		esc = o + 2;
	}

	const uint32_t i0 = sentinel_index(r.x, bias), i1 = sentinel_index(r.y, bias), i2 = sentinel_index(r.z, bias), i3 = sentinel_index(r.w, bias);
	if ((i0 | i1 | i2 | i3) == 0)
		return r;
#ifdef MOB200_SHARED_PATCH
	return patch_group(ring, r, esc, i0 | (i1 << 4) | (i2 << 8) | (i3 << 12), patch_lut);
#else
	// (the 32-bit windows may reach a few bytes past the last escape byte: still inside the staging ring, never selected)
	const uint32_t e1 = esc + __popc(i0), e2 = e1 + __popc(i1), e3 = e2 + __popc(i2);
	r.x = __byte_perm(r.x, lds_u32_at(ring, esc), patch_lut[i0]);
	r.y = __byte_perm(r.y, lds_u32_at(ring, e1), patch_lut[i1]);
	r.z = __byte_perm(r.z, lds_u32_at(ring, e2), patch_lut[i2]);
	r.w = __byte_perm(r.w, lds_u32_at(ring, e3), patch_lut[i3]);
	return r;
#endif
}

struct BlockRegs
{
	uint32_t vs, n, groups, gshift, items, stage_off, rows_off, filter, filter_kind, m_chunk;
	bool first_block;
	uint32_t chain; // plain form: the block follows the unit's previous block in its stream: its carry is the unit's running value
	const uint16_t* rows_global;
	uint8_t* out;
	unsigned long long* lookback;
};

__device__ __forceinline__ BlockRegs load_block(const SlotData& S)
{
	const uint4 p0 = *reinterpret_cast<const uint4*>(&S.P.valid);
	const uint4 p1 = *reinterpret_cast<const uint4*>(&S.P.gshift);
	const uint4 p2 = *reinterpret_cast<const uint4*>(&S.P.first);
	BlockRegs B;
	B.vs = p0.y, B.n = p0.z, B.groups = p0.w;
	B.gshift = p1.x, B.items = p1.y, B.stage_off = p1.z, B.rows_off = p1.w;
	B.first_block = p2.x != 0, B.filter = p2.y, B.filter_kind = p2.z, B.m_chunk = p2.w;
	B.rows_global = S.P.rows_global;
	B.out = S.P.out;
	B.lookback = S.P.lookback;
	B.chain = S.P.chain;
	return B;
}

// bytes of the output tile one block occupies: whole 16-vertex chunks (the last chunk is written in full even when
// the block ends inside it) plus the per-chunk padding; a multiple of 16 because vs is a multiple of 4
__device__ __forceinline__ uint32_t tile_span(uint32_t groups, uint32_t vs)
{
	return groups * (16u * vs + kTilePad) + ((groups & 1u) ? kTilePad : 0u);
}

struct DecoderCtx
{
	uint8_t* ring;
	const uint32_t* patch_lut;
	uint64_t* tile_free;
	unsigned long long tag;
	uint32_t lane;
	uint32_t* unit_prefix; // plain form, block mode: the running value the unit hands from a block to the next one of the same stream
	const DevTables* T; // (event trace)
	uint32_t unit, warp, index;
};

// One work quantum: 32 items of one block (item = 4 byte-channels x 16 vertices): unpack, transposes, deltas and scans
// in registers, finished words into the block's part of the output tile.
// kChain (rounds form): the quantum belongs to member `chain_g` of a chained round whose first member sits in slot S0
// (carry_slot / phase are that member's): the members' aggregates meet in shared memory, one barrier of the decoder
// warps later every member knows the value in front of it without a look-back through global memory.
// kUnit (plain form): a block whose predecessor in the stream was this unit's previous block takes its carry from the
// unit's running value in shared memory (two copies, by block parity: a block reads one and writes the other) -- no
// look-back, no global round trip; the producer only completes the carry barrier's phase.
template <bool kChain, bool kUnit = false>
__device__ __forceinline__ void decode_quantum(const DecoderCtx& X, const SlotData& S, const BlockRegs& B, uint8_t* tile, uint64_t* carry_slot, uint32_t phase,
    uint32_t base, bool first_of_block, uint32_t tile_uses, long long& dbg_carry, long long& dbg_tile, SlotData* slots = nullptr, uint32_t slot0 = 0, uint32_t slot_mask = 0,
    uint32_t chain_g = 0, uint32_t bar_id = 0)
{
	uint8_t* ring = X.ring;
	const uint32_t* patch_lut = X.patch_lut;
	uint64_t* tile_free = X.tile_free;
	const unsigned long long tag = X.tag;
	const uint32_t lane = X.lane;
	const uint32_t vs = B.vs, groups = B.groups, gshift = B.gshift, items = B.items, stage_off = B.stage_off, rows_off = B.rows_off;
	const bool first_block = B.first_block;
	const uint16_t* rows_global = B.rows_global;
	unsigned long long* lookback = B.lookback;
	const uint32_t gstride = 1u << gshift;

	const uint32_t item = base + lane;
	// item -> (4-byte lane q, chunk c): consecutive items share a lane, a warp takes two NEIGHBOURING lanes of a 32-byte
	// vertex.  (Neighbours are of one kind and take the same unpack paths.  The unit trace shows the price -- unpack 2.2 /
	// 1.8 / 2.4 / 2.7 us per warp on C2b, the block's barrier waits for the slowest -- but pairing lane k with lane
	// k + nq/2 to even the warps out measured 7 % SLOWER: two kinds of lanes in a warp run both kinds' paths.)
	const uint32_t q = item >> gshift;
	const uint32_t c = item & (gstride - 1u);
	const bool active = item < items && c < groups;

	uint32_t w[16];
	uint32_t total = 0;      // lane-wise sum of the 16 deltas
	uint32_t channel = 0;
	if (active)
	{
		channel = S.channels[q];

		// this thread's four groups: byte-channels 4q..4q+3, group c
		uint32_t e0, e1, e2, e3;
		if (rows_off != kRowsInGlobal)
		{
			const uint16_t* rows = reinterpret_cast<const uint16_t*>(ring + rows_off) + (4 * q) * 16 + c;
			e0 = rows[0], e1 = rows[16], e2 = rows[32], e3 = rows[48];
		}
		else
		{
			const uint16_t* rows = rows_global + (4 * q) * 16 + c;
			e0 = __ldcg(rows), e1 = __ldcg(rows + 16), e2 = __ldcg(rows + 32), e3 = __ldcg(rows + 48);
		}
#if defined(MOB200_UNPACK_ONE_COPY)
		// (variant: one copy of the unpack code, four passes)
		uint4 pa = make_uint4(0, 0, 0, 0), pb = pa, pc = pa, pd = pa;
#pragma unroll 1
		for (int h = 0; h < 4; ++h)
		{
			const uint4 x = unpack_group(ring, stage_off, h == 0 ? e0 : (h == 1 ? e1 : (h == 2 ? e2 : e3)), patch_lut);
			if (h == 0)
				pa = x;
			else if (h == 1)
				pb = x;
			else if (h == 2)
				pc = x;
			else
				pd = x;
		}
#elif !defined(MOB200_UNPACK_FOUR_COPIES)
		// two passes over two copies of the unpack code: the pair of a pass still interleaves, and the hot code of the
		// three roles, which share the SM's instruction cache, is 4.9 KB smaller than with four copies (measured r2:
		// 1.583 -> 1.559 ms serial, 1.686 -> 1.656 ms block mode; ONE copy in four passes loses the interleaving: 2.09 ms)
		uint4 pa = make_uint4(0, 0, 0, 0), pb = pa, pc = pa, pd = pa;
#pragma unroll 1
		for (int h = 0; h < 2; ++h)
		{
			const uint4 x = unpack_group(ring, stage_off, h ? e2 : e0, patch_lut);
			const uint4 y = unpack_group(ring, stage_off, h ? e3 : e1, patch_lut);
			if (h == 0)
				pa = x, pb = y;
			else
				pc = x, pd = y;
		}
#else
		uint4 pa = unpack_group(ring, stage_off, e0, patch_lut);
		uint4 pb = unpack_group(ring, stage_off, e1, patch_lut);
		uint4 pc = unpack_group(ring, stage_off, e2, patch_lut);
		uint4 pd = unpack_group(ring, stage_off, e3, patch_lut);
#endif

		const bool bytes = (channel & 3u) == 0;
		if (bytes)
		{
			// byte deltas: un-zigzag in plane form, totals with dp4a (one add per four values)
			pa.x = unzig8x4(pa.x), pa.y = unzig8x4(pa.y), pa.z = unzig8x4(pa.z), pa.w = unzig8x4(pa.w);
			pb.x = unzig8x4(pb.x), pb.y = unzig8x4(pb.y), pb.z = unzig8x4(pb.z), pb.w = unzig8x4(pb.w);
			pc.x = unzig8x4(pc.x), pc.y = unzig8x4(pc.y), pc.z = unzig8x4(pc.z), pc.w = unzig8x4(pc.w);
			pd.x = unzig8x4(pd.x), pd.y = unzig8x4(pd.y), pd.z = unzig8x4(pd.z), pd.w = unzig8x4(pd.w);
			const uint32_t ta = sum_bytes(pa), tb = sum_bytes(pb), tc = sum_bytes(pc), td = sum_bytes(pd);
			total = __byte_perm(__byte_perm(ta, tb, 0x0040), __byte_perm(tc, td, 0x0040), 0x5410);
		}

		// 4 planes x 16 bytes -> 16 vertex words
		const uint32_t A[4] = {pa.x, pa.y, pa.z, pa.w};
		const uint32_t B[4] = {pb.x, pb.y, pb.z, pb.w};
		const uint32_t C[4] = {pc.x, pc.y, pc.z, pc.w};
		const uint32_t D[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			const uint32_t t0 = __byte_perm(A[j], B[j], 0x5140);
			const uint32_t t1 = __byte_perm(A[j], B[j], 0x7362);
			const uint32_t u0 = __byte_perm(C[j], D[j], 0x5140);
			const uint32_t u1 = __byte_perm(C[j], D[j], 0x7362);
			w[4 * j + 0] = __byte_perm(t0, u0, 0x5410);
			w[4 * j + 1] = __byte_perm(t0, u0, 0x7632);
			w[4 * j + 2] = __byte_perm(t1, u1, 0x5410);
			w[4 * j + 3] = __byte_perm(t1, u1, 0x7632);
		}

		if (!bytes)
		{
			if ((channel & 3u) == 1)
			{
				// 16-bit deltas: un-zigzag per half word, totals with two-lane adds (VIADD.16x2)
#pragma unroll
				for (int j = 0; j < 16; ++j)
					w[j] = ((w[j] >> 1) & 0x7fff7fffu) ^ ((w[j] & 0x00010001u) * 0xffffu);
				total = w[0];
#pragma unroll
				for (int j = 1; j < 16; ++j)
					total = __vadd2(total, w[j]);
			}
			else
			{
				// xor deltas: rotate right by the channel's amount
				const uint32_t rot = (32u - (channel >> 4)) & 31u;
#pragma unroll
				for (int j = 0; j < 16; ++j)
					w[j] = __funnelshift_l(w[j], w[j], rot);
				total = w[0];
#pragma unroll
				for (int j = 1; j < 16; ++j)
					total ^= w[j];
			}
		}
	}
	else
	{
#pragma unroll
		for (int j = 0; j < 16; ++j)
			w[j] = 0;
	}
	const uint32_t H = lane_mask(channel);

	// scan of the chunk totals across the gstride threads of this lane q (inactive chunks add 0)
	uint32_t incl = total;
#pragma unroll
	for (uint32_t dlt = 1; dlt < 16; dlt <<= 1)
	{
		const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dlt, 16); // (gstride divides 16: c >= dlt keeps the segments apart)
		if (c >= dlt)
			incl = lane_combine(o, incl, H);
	}
	uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1, gstride);
	if (c == 0)
		excl = 0;

	// (block 0 publishes its inclusive prefix only: a look-back must never step below it)
	const bool last = active && c == groups - 1;
	if (last && !first_block)
		st_volatile_u64(lookback + q, tag | (1ull << 32) | incl); // state 1: aggregate of this block

	if (kChain)
	{
		if (last)
			slots[(slot0 + chain_g) & slot_mask].chain_total[q] = incl;
		decoder_sync(bar_id); // all four decoder warps (a warp without a quantum in this round joins in decoder_main)
	}

	if (first_of_block)
	{
		const long long c0 = dbg_clock();
		MOB200_TRACE_EVENT(*X.T, X.unit, lane, 32 + X.warp * 8 + 2, X.index);
		mbar_wait(carry_slot, phase);
		MOB200_TRACE_EVENT(*X.T, X.unit, lane, 32 + X.warp * 8 + 3, X.index);
		const long long c1 = dbg_clock();
		mbar_wait(tile_free, (tile_uses & 1u) ^ 1u); // every warp has finished storing the previous tile
		dbg_carry += c1 - c0;
		dbg_tile += dbg_clock() - c1;
	}

	if (active)
	{
		uint32_t carry = kChain ? slots[slot0 & slot_mask].carry[q] : S.carry[q];
		if (kChain)
			for (uint32_t h = 0; h < chain_g; ++h)
				carry = lane_combine(carry, slots[(slot0 + h) & slot_mask].chain_total[q], H);
		if (kUnit && B.chain)
			carry = X.unit_prefix[(tile_uses & 1u) * 64u + q];
		if (last)
		{
			const uint32_t prefix = lane_combine(carry, incl, H);
			st_volatile_u64(lookback + q, tag | (2ull << 32) | prefix); // state 2: inclusive prefix
			if (kUnit)
				X.unit_prefix[((tile_uses & 1u) ^ 1u) * 64u + q] = prefix;
		}
		uint32_t v = lane_combine(carry, excl, H);
		uint8_t* col = tile + tile_offset(c * 16, vs) + q * 4;
		// one instruction stream for the three channel modes: a warp whose two lanes differ in mode does not run it twice
#pragma unroll
		for (int j = 0; j < 16; ++j)
		{
			v = lane_combine(v, w[j], H);
			*reinterpret_cast<uint32_t*>(col) = v;
			col += vs;
		}
	}
}

// kWide: a 16-byte piece's elements go through the filter in one call (rounds form, where filtered small vertices are the
// norm).  The plain-form kernels keep the single-element calls: their code must not move -- with the wide functions
// linked in, the serial-walk kernel measured 4.7 % slower on unfiltered 32-byte vertices (layout of the hot code).
template <bool kWide>
__device__ __forceinline__ void store_block(const BlockRegs& B, uint8_t* tile, uint32_t tid, uint32_t bar_id)
{
	const uint32_t vs = B.vs, n = B.n;
	uint8_t* out = B.out;
	const uint4 p2 = make_uint4(0u, B.filter, B.filter_kind, B.m_chunk);

	// ---- tile -> global memory, decode filter on the way out ---------------------------------------------------------
	{
		const uint32_t nbytes = n * vs;
		const uint32_t m_chunk = p2.w;
		const int filter = (int)p2.y;
		uint32_t fk = p2.z;
		const uintptr_t oa = reinterpret_cast<uintptr_t>(out);
		if (fk != 0 && (oa & 15) != 0)
		{
			// unaligned destination: filter inside the tile first (rare)
			if (fk == 1)
				for (uint32_t o = tid * 4; o < nbytes; o += kDecodeThreads * 4)
				{
					uint32_t* e = reinterpret_cast<uint32_t*>(tile + o + __umulhi(o, m_chunk) * kTilePad);
					*e = apply_filter32(*e, filter);
				}
			else
				for (uint32_t r = tid; r < n; r += kDecodeThreads)
				{
					uint2* e = reinterpret_cast<uint2*>(tile + tile_offset(r, vs));
					*e = apply_filter64(*e, filter);
				}
			decoder_sync(bar_id);
			fk = 0;
		}
		if ((oa & 15) == 0)
		{
			const uint32_t pieces = nbytes >> 4;
			if (fk == 0)
			{
#pragma unroll 4
				for (uint32_t j = tid; j < pieces; j += kDecodeThreads)
				{
					const uint32_t o = j << 4;
					const uint2* p = reinterpret_cast<const uint2*>(tile + o + __umulhi(o, m_chunk) * kTilePad);
					// (two 8-byte loads at a 16-byte lane stride are 2-way bank conflicts; a conflict-free lane order
					// with four selects measured the same)
					const uint2 lo = p[0], hi = p[1];
					*reinterpret_cast<uint4*>(out + o) = make_uint4(lo.x, lo.y, hi.x, hi.y);
				}
			}
			else
			{
				for (uint32_t j = tid; j < pieces; j += kDecodeThreads)
				{
					const uint32_t o = j << 4;
					const uint2* p = reinterpret_cast<const uint2*>(tile + o + __umulhi(o, m_chunk) * kTilePad);
					uint2 lo = p[0], hi = p[1];
					if (kWide)
					{
						const uint4 piece = make_uint4(lo.x, lo.y, hi.x, hi.y);
						*reinterpret_cast<uint4*>(out + o) = fk == 1 ? apply_filter32x4(piece, filter) : apply_filter64x2(piece, filter);
						continue;
					}
					if (fk == 1)
					{
						lo.x = apply_filter32(lo.x, filter), lo.y = apply_filter32(lo.y, filter);
						hi.x = apply_filter32(hi.x, filter), hi.y = apply_filter32(hi.y, filter);
					}
					else
					{
						lo = apply_filter64(lo, filter);
						hi = apply_filter64(hi, filter);
					}
					*reinterpret_cast<uint4*>(out + o) = make_uint4(lo.x, lo.y, hi.x, hi.y);
				}
			}
			const uint32_t rem_words = (nbytes & 15u) >> 2;
			if (tid < rem_words)
			{
				const uint32_t o = (pieces << 4) + tid * 4;
				const uint32_t* e = reinterpret_cast<const uint32_t*>(tile + o + __umulhi(o, m_chunk) * kTilePad);
				if (fk == 2)
				{
					// one 8-byte element is left over (vs = 8, odd vertex count): thread 0 / 1 keep their half
					const uint2 r = apply_filter64(*reinterpret_cast<const uint2*>(e - tid), filter);
					*reinterpret_cast<uint32_t*>(out + o) = tid ? r.y : r.x;
				}
				else
					*reinterpret_cast<uint32_t*>(out + o) = fk == 1 ? apply_filter32(*e, filter) : *e;
			}
		}
		else if ((oa & 3) == 0)
		{
			for (uint32_t o = tid * 4; o < nbytes; o += kDecodeThreads * 4)
				*reinterpret_cast<uint32_t*>(out + o) = *reinterpret_cast<const uint32_t*>(tile + o + __umulhi(o, m_chunk) * kTilePad);
		}
		else
		{
			for (uint32_t o = tid; o < nbytes; o += kDecodeThreads)
				out[o] = tile[o + __umulhi(o, m_chunk) * kTilePad];
		}
	}
}

// A decoder warp is done with a slot.  Plain form: the elected lane arrives for the warp (after __syncwarp: ordered
// under the PTX memory model; one shared-memory atomic per warp and block).  Rounds form: every thread arrives for its
// own reads -- the members' store parameters are read from the slots by all threads right before, and compute-
// sanitizer's racecheck credits an arrival only to the arriving thread; per round the cost does not show.
template <bool kRounds>
__device__ __forceinline__ void release_slot(uint64_t* empty_bar, uint32_t lane)
{
	if (kRounds)
		mbar_arrive(empty_bar);
	else
	{
		__syncwarp();
		if (lane == 0)
			mbar_arrive(empty_bar);
	}
}

template <bool kRounds, bool kBlock>
__device__ void decoder_main(const DevTables& T, uint8_t* smem, const uint32_t unit, const uint32_t tid, const uint32_t bar_id)
{
	using L = Lay<kRounds>;
	constexpr uint32_t kSlots = L::kSlots;
	uint8_t* tile = smem + L::kSmemTile;
	SlotData* slots = reinterpret_cast<SlotData*>(smem + L::kSmemSlots);
	uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kSmemBars);
	uint64_t* carry_bar = full + kSlots;
	uint64_t* empty = carry_bar + kSlots;
	uint64_t* tile_free = empty + kSlots;

	const uint32_t lane = tid & 31u;
	const uint32_t warp_base = tid & ~31u;
	DecoderCtx X;
	X.ring = smem + L::kSmemStage;
	X.patch_lut = reinterpret_cast<const uint32_t*>(smem + L::kSmemPatch);
	X.tile_free = tile_free;
	X.tag = (unsigned long long)(T.epoch << 2) << 32;
	X.lane = lane;
	X.unit_prefix = reinterpret_cast<uint32_t*>(smem + L::kSmemPrefix);
	X.T = &T, X.unit = unit, X.warp = tid >> 5, X.index = 0;
	uint32_t tile_uses = 0;
	long long dbg_full = 0, dbg_carry = 0, dbg_tile = 0;
	const long long dbg_t0 = dbg_clock();
	const uint32_t my_count = unit_block_count(T, unit);

	// Rounds variant: one ROUND = consecutive blocks of this unit's sequence whose work fits the four decoder warps: a
	// work quantum is 32 items (one warp), a block has ceil(items / 32) of them.  Blocks of small vertices (4 ... 16
	// bytes: one or two quanta) share a round, each in its own part of the output tile, so that no decoder warp idles
	// while its unit works on a block too small for four warps.  The producer decides the rounds
	// (BlockParams::round_members of a round's first block).
	for (uint32_t i = 0; i < my_count;)
	{
		const uint32_t slot = i & (kSlots - 1);
		const uint32_t phase = (i / kSlots) & 1u;
		const SlotData& S = slots[slot];
		X.index = i;
		MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 0, i);
		{
			const long long c0 = dbg_clock();
			mbar_wait(full + slot, phase);
			dbg_full += dbg_clock() - c0;
		}
		const uint32_t members = kRounds ? S.P.round_members : 1u;

		if (members <= 1)
		{
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 1, i);
			// ---- one block: its quanta go round the four warps ---------------------------------------------------------
			++i;
			// (lane 0 reads the flag for the warp: it is the lane that hands the slot back below, and compute-sanitizer's
			// racecheck credits an arrival only to the arriving thread)
			if (!__shfl_sync(0xffffffffu, lane == 0 ? S.P.valid : 0u, 0))
			{
				release_slot<kRounds>(empty + slot, lane);
				continue;
			}
			const BlockRegs B = load_block(S);
			for (uint32_t base = warp_base; base < B.items; base += kDecodeThreads)
				decode_quantum<false, !kRounds && kBlock>(X, S, B, tile, carry_bar + slot, phase, base, base == warp_base, tile_uses, dbg_carry, dbg_tile);
			++tile_uses;

			// this warp no longer needs the slot (staging bytes, rows, params, carry)
			release_slot<kRounds>(empty + slot, lane);

			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 4, i - 1);
			decoder_sync(bar_id); // the tile is complete
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 5, i - 1);
			store_block<kRounds>(B, tile, tid, bar_id);
			// every thread arrives for its own reads of the tile (an elected lane's arrival after __syncwarp is just as
			// ordered under the PTX memory model and measured the same, but compute-sanitizer's racecheck credits an
			// arrival only to the arriving thread)
			mbar_arrive(tile_free);
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 6, i - 1);
			continue;
		}

		if (kRounds)
		{
			// ---- a round of several small blocks: one quantum per warp, every block in its own part of the tile ----------
			uint32_t quanta = 0, tile_used = 0, round_n = 0;
			uint32_t mq[kRoundBlocks], moff[kRoundBlocks];
#pragma unroll
			for (uint32_t g = 0; g < kRoundBlocks; ++g)
			{
				mq[g] = 0, moff[g] = 0;
				if (g < members)
				{
					const uint32_t sg = (i + g) & (kSlots - 1);
					if (g > 0)
					{
						const long long c0 = dbg_clock();
						mbar_wait(full + sg, ((i + g) / kSlots) & 1u);
						dbg_full += dbg_clock() - c0;
					}
					const SlotData& Sg = slots[sg];
					if (Sg.P.valid)
					{
						mq[g] = (Sg.P.items + 31u) >> 5;
						moff[g] = tile_used;
						tile_used += tile_span(Sg.P.groups, Sg.P.vs);
						quanta += mq[g];
						round_n += Sg.P.n;
					}
				}
			}
			const uint32_t warp = tid >> 5;
			const bool chained = S.P.chain != 0;
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 1, i);
			if (warp >= quanta)
			{
				if (chained)
					decoder_sync(bar_id); // the barrier inside the chained decode_quantum
			}
			else
			{
				// the member this warp's quantum belongs to
				uint32_t g = 0, qb = 0;
#pragma unroll
				for (uint32_t h = 0; h + 1 < kRoundBlocks; ++h)
					if (g == h && h + 1 < members && warp >= qb + mq[h])
						qb += mq[h], g = h + 1;
				uint8_t* btile = tile + (g == 0 ? moff[0] : (g == 1 ? moff[1] : (g == 2 ? moff[2] : moff[3])));
				const uint32_t sg = (i + g) & (kSlots - 1);
				const SlotData& Sg = slots[sg];
				const BlockRegs B = load_block(Sg);
				if (chained)
					decode_quantum<true>(X, Sg, B, btile, carry_bar + slot, phase, (warp - qb) * 32u, true, tile_uses, dbg_carry, dbg_tile, slots, i, kSlots - 1, g, bar_id);
				else
					decode_quantum<false>(X, Sg, B, btile, carry_bar + sg, ((i + g) / kSlots) & 1u, (warp - qb) * 32u, true, tile_uses, dbg_carry, dbg_tile);
			}
			++tile_uses;

			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 4, i);
			decoder_sync(bar_id); // the tiles of the round are complete
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 5, i);

			// a chained round is one piece of output: consecutive blocks of one stream, every member but the last 256 vertices
			// (16 chunks), so the members' outputs AND their parts of the tile follow each other -- one store pass
			const uint32_t stores = chained ? 1u : members;
#pragma unroll 1
			for (uint32_t g = 0; g < stores; ++g)
			{
				const SlotData& Sg = slots[(i + g) & (kSlots - 1)];
				if (!Sg.P.valid)
					continue;
				BlockRegs B = load_block(Sg);
				if (chained)
					B.n = round_n;
				// (member g starts at warp g: a block of 4-byte vertices has 64 sixteen-byte pieces, two warps' worth)
				store_block<true>(B, tile + (g == 0 ? moff[0] : (g == 1 ? moff[1] : (g == 2 ? moff[2] : moff[3]))), (tid + g * 32u) & (kDecodeThreads - 1), bar_id);
			}
			// the slots are released after the stores: the store parameters of the members are read from them (with eight
			// slots the producer still stages the whole next round meanwhile)
#pragma unroll
			for (uint32_t g = 0; g < kRoundBlocks; ++g)
				if (g < members)
					release_slot<true>(empty + ((i + g) & (kSlots - 1)), lane);
			mbar_arrive(tile_free);
			MOB200_TRACE_EVENT(T, unit, lane, 32 + X.warp * 8 + 6, i);
			i += members;
		}
	}

#ifdef MOB200_DEBUG_COUNTERS
	if (tid == 0)
	{
		unsigned long long* dbg = debug_counters(T);
		atomicAdd(dbg + kDbgDecoderTotal, (unsigned long long)(dbg_clock() - dbg_t0));
		atomicAdd(dbg + kDbgDecoderWaitFull, (unsigned long long)dbg_full);
		atomicAdd(dbg + kDbgDecoderWaitCarry, (unsigned long long)dbg_carry);
		atomicAdd(dbg + kDbgDecoderWaitTile, (unsigned long long)dbg_tile);
	}
#endif
}

} // namespace mob200
