// mob200_common.h -- wire-format constants and the structures shared by the host shim and the kernels.
//
// Wire format facts (not code) come from the reference: src/vertexcodec.cpp:123-147 (constants and the
// block-size rule), :1638-1693 (stream framing), :1531-1568 (v1 control modes); SURVEY.md Appendix A
// restates them precisely.
#pragma once

#include <stdint.h>

namespace mob200
{

constexpr uint32_t kMagic = 0xa0;          // high nibble of byte 0; low nibble = version (0 or 1)
constexpr uint32_t kBlockBytes = 8192;     // decoded bytes per block never exceed this
constexpr uint32_t kBlockMaxVerts = 256;   // vertices per block never exceed this
constexpr uint32_t kGroup = 16;            // values per group
constexpr uint32_t kGroupReadLimit = 24;   // bytes that must remain readable before each group
constexpr uint32_t kTailMinV0 = 32;
constexpr uint32_t kTailMinV1 = 24;

constexpr uint32_t kPlainRunShift = 4; // plain form, block mode: runs of 16 consecutive blocks of a stream per unit (= half a producer batch)
constexpr uint32_t kRunShiftMax = 2; // run-major decode order: 1 << shift consecutive blocks of a stream per ticket chunk (= the blocks of a decode round:
                                     // four of <= 8-byte vertices, two of 12- / 16-byte vertices)
constexpr uint32_t kInvalidOffset = 0xffffffffu; // walk result for a block that must not be decoded

// largest encoded size of one block over all legal vertex sizes (vs = 256: 256*(1+2*24)+64 = 12608)
constexpr uint32_t kMaxEncodedBlock = 12608;

#if defined(__CUDACC__)
#define MOB200_HD __host__ __device__ __forceinline__
#else
#define MOB200_HD inline
#endif

// vertices per block for a vertex size (reference src/vertexcodec.cpp:140-147)
MOB200_HD uint32_t block_vertices(uint32_t vertex_size)
{
	uint32_t n = (kBlockBytes / vertex_size) & ~(kGroup - 1);
	return n < kBlockMaxVerts ? n : kBlockMaxVerts;
}

MOB200_HD uint32_t tail_bytes(uint32_t vertex_size, uint32_t version)
{
	return vertex_size + (version ? vertex_size / 4 : 0);
}

MOB200_HD uint32_t tail_padded(uint32_t vertex_size, uint32_t version)
{
	uint32_t t = tail_bytes(vertex_size, version);
	uint32_t m = version ? kTailMinV1 : kTailMinV0;
	return t < m ? m : t;
}

// One encoded stream as the kernels see it.  48 bytes.  The host sorts streams by block count
// (descending) before upload, so that (a) the 32 lanes of a walker warp get streams of similar
// length and (b) "level b" of the decode order is a prefix of the array.
struct DevStream
{
	const uint8_t* src;    // encoded bytes (device)
	uint8_t* dst;          // decoded vertices (device)
	uint64_t chan_base;    // first row of this stream in the group table (one row = one byte-channel
	                       // of one block); chan_base/4 is its first entry in the look-back table
	uint32_t src_size;
	uint32_t vertex_count;
	uint32_t block_base;   // global id of this stream's block 0
	uint32_t nblocks;
	uint16_t vertex_size;
	uint8_t filter;        // enum mob200_Filter
	uint8_t block_groups;  // vertices per block / 16 (block_vertices(vertex_size): spares the kernels an integer division)
	uint32_t caller_index; // position of the stream in the caller's array (status is reported there)
};

// Per-plan device tables.
struct DevTables
{
	const DevStream* streams;
	uint32_t* block_offset;       // [total_blocks + n_streams]: stream s, block b at block_base + s + b (entry
	                              // nblocks = end of the last block); kInvalidOffset = do not decode.  Written
	                              // by the walker warps, read by the decoders.
	uint16_t* group_table;        // per (block, byte-channel): 16 entries, one per 16-value group:
	                              // 0 = all zero, else (offset_in_block << 2) | log2(bits).  Walker -> decoder.
	uint32_t* block_ready;        // [total_blocks], block b of stream s at block_base + b: (epoch << 2) | codec version << 1 | decodable,
	                              // release-stored by the walker that walked (normal mode) or verified (block mode) the block,
	                              // acquire-polled by the producers
	unsigned long long* lookback; // per block, vertex_size/4 entries: (epoch << 2 | state) << 32 | value
	const uint2* ticket_info;     // [total_blocks]: decode order -> (stream, block)
	int32_t* status;              // [n_streams] in CALLER order: reference return code per stream
	uint32_t* counters;           // [0] decode ticket, [1] walker stream ticket, [2] finished roles
	uint32_t n_streams;
	uint32_t n_with_blocks;       // streams [0, n_with_blocks) of the sorted array have at least one block
	uint32_t total_blocks;
	uint32_t units;               // decode units that take part in this run (unit u decodes tickets u, u + units, ...)
	uint32_t epoch;               // changes every run, so progress / look-back entries never need clearing
	uint32_t walker_lead;         // 0xffffffff = walk-only diagnostic mode (decoders off); otherwise unused
	uint32_t wide_walk;           // 1: one warp per stream (few streams), 0: one lane per stream (many streams)
	uint32_t rounds;              // 1: most blocks are small-vertex blocks (<= 16 bytes per vertex): decode them in rounds of up to four
	uint32_t block_mode;          // 1: block_offset is an INPUT (a block-offset sidecar: caller-provided, or kept from an earlier run):
	                              // every block is walked on its own by one walker lane, which also checks that the block ends where
	                              // the next one is said to start -- a stale sidecar is detected, never trusted (kStatusSidecar)
	uint32_t ticket_shift;        // log2 of the ticket chunk: unit u takes tickets in chunks of 1 << ticket_shift (chunk c of the order goes to
	                              // unit c % units).  0 with the level-major order; 2 with the run-major order of block mode + rounds,
	                              // where a chunk is four consecutive blocks of one stream: they share a decode round, so only the first
	                              // of them looks back across units
	uint32_t keep_status;         // block mode after the team walk: the status words already hold the reference codes; only a
	                              // stream that is still 0 may be marked kStatusSidecar
};

constexpr int32_t kStatusSidecar = -102; // = MOB200_ERR_SIDECAR (include/meshopt_b200.h)

} // namespace mob200
