// mob200_encode.cpp -- encoder-side helper of the decode path: the SEGMENTER.
//
// The GPU decoder gets its parallelism from independent chains (section 3 of DESIGN.md): many short streams, or
// long streams that come with their block-offset sidecar.  Nothing in the reference produces either, so this file
// does, on the host, for an asset pipeline that targets the GPU decoder:
//
//   mob200_encode_vertex_buffer   one stream in the reference wire format (v0 / v1, levels 0-3), optionally writing
//                                 the stream's sidecar -- the offsets fall out of encoding for free;
//   mob200_encode_segments        a vertex array -> K independently decodable streams (each a valid input of the
//                                 unmodified meshopt_decodeVertexBuffer), packed on 16-byte boundaries, with the
//                                 descriptor array and all sidecars, encoded on several host threads.
//
// Wire format and the encoder's choices follow the reference (src/vertexcodec.cpp:183-579 group / channel / control
// selection, :1615-1693 framing), so that a stream made here has the size the reference encoder would give it -- the
// tests compare the bytes with the reference encoder's and decode every stream with the reference decoder.  The code
// is organised differently: a block is first turned into all of its delta PLANES (one 256-byte row per byte of the
// vertex), every 16-value group of a plane is summarised once by three counters (values above the 1-, 2- and 4-bit
// sentinels) from which the cost of every width follows, and planes are emitted from those summaries.
//
// Host code only; no CUDA.  (The decode side never calls into this file.)
#include "../../include/meshopt_b200.h"

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include <vector_types.h> // (uint2 in the table structs of mob200_common.h; only the format constants are used here)

#include "mob200_common.h"

namespace
{

using namespace mob200;

constexpr size_t kGroupBytes = kGroup;

inline uint32_t rotl32(uint32_t v, int r)
{
	return (v << r) | (v >> ((32 - r) & 31));
}

// ---- group summaries ----------------------------------------------------------------------------------------------

// what the encoder needs to know about 16 delta bytes: how many of them do not fit below the sentinel of each width
struct GroupInfo
{
	uint8_t over1, over2, over4; // values >= 1, >= 3, >= 15
	bool zero() const { return over1 == 0; }
	// encoded size for a field width in bits (0: only an all-zero group can use it)
	size_t cost(int bits) const
	{
		switch (bits)
		{
		case 0: return zero() ? 0 : SIZE_MAX;
		case 1: return 2 + over1;
		case 2: return 4 + over2;
		case 4: return 8 + over4;
		default: return kGroupBytes;
		}
	}
};

inline GroupInfo summarise(const uint8_t* g)
{
	GroupInfo s = {0, 0, 0};
	for (size_t i = 0; i < kGroupBytes; ++i)
	{
		s.over1 += g[i] >= 1;
		s.over2 += g[i] >= 3;
		s.over4 += g[i] >= 15;
	}
	return s;
}

// ---- emitting one plane ---------------------------------------------------------------------------------------------

// 16 values as `bits`-wide fields (MSB-first inside a byte; 1-bit groups LSB-first, reference :243-246) followed by
// one escape byte per value that did not fit
uint8_t* emit_group(uint8_t* out, const uint8_t* g, int bits)
{
	if (bits == 0)
		return out;
	if (bits == 8)
	{
		memcpy(out, g, kGroupBytes);
		return out + kGroupBytes;
	}
	const unsigned sentinel = (1u << bits) - 1;
	const size_t per_byte = 8 / bits;
	for (size_t i = 0; i < kGroupBytes; i += per_byte)
	{
		unsigned packed = 0;
		for (size_t k = 0; k < per_byte; ++k)
		{
			const unsigned field = g[i + k] >= sentinel ? sentinel : g[i + k];
			if (bits == 1)
				packed |= field << k; // LSB-first
			else
				packed = (packed << bits) | field;
		}
		*out++ = (uint8_t)packed;
	}
	for (size_t i = 0; i < kGroupBytes; ++i)
		if (g[i] >= sentinel)
			*out++ = g[i];
	return out;
}

// One bit-packed plane: header of 2-bit width selectors, then the groups.  widths[4] is the table the selectors
// index ({0,2,4,8} for v0; {0,1,2,4} / {1,2,4,8} for the two v1 control modes).  Returns NULL when out of space.
uint8_t* emit_plane(uint8_t* out, uint8_t* out_end, const uint8_t* plane, const GroupInfo* info, size_t groups, const int widths[4])
{
	const size_t header_bytes = (groups + 3) / 4;
	if ((size_t)(out_end - out) < header_bytes)
		return nullptr;
	uint8_t* header = out;
	memset(header, 0, header_bytes);
	out += header_bytes;

	int previous = -1;
	for (size_t g = 0; g < groups; ++g)
	{
		if ((size_t)(out_end - out) < kGroupReadLimit)
			return nullptr;
		// cheapest width; among equals the widest wins unless a narrower one continues the previous group's width
		// (never in place of a raw 8-bit group) -- the reference's rule, :281-293
		int pick = 3;
		size_t best = info[g].cost(widths[3]);
		for (int k = 0; k < 3; ++k)
		{
			const size_t c = info[g].cost(widths[k]);
			if (c < best || (c == best && widths[k] == previous && widths[pick] != 8))
			{
				pick = k;
				best = c;
			}
		}
		header[g / 4] |= (uint8_t)(pick << ((g % 4) * 2));
		out = emit_group(out, plane + g * kGroupBytes, widths[pick]);
		previous = widths[pick];
	}
	return out;
}

// ---- deltas ---------------------------------------------------------------------------------------------------------

// Delta planes of the four bytes of one 32-bit lane (bytes k .. k+3 of every vertex) of a block of n vertices.
// mode 0: per-byte zigzag deltas; 1: per-16-bit zigzag deltas; 2: xor with the previous word, rotated left by rot.
// `prev` = the four bytes of the vertex before the block.  planes: 4 rows of `stride` bytes.
void delta_planes(uint8_t* planes, size_t stride, const uint8_t* vertices, size_t n, size_t vertex_size, size_t k, const uint8_t* prev, int mode, int rot)
{
	uint32_t p = (uint32_t)prev[0] | ((uint32_t)prev[1] << 8) | ((uint32_t)prev[2] << 16) | ((uint32_t)prev[3] << 24);
	const uint8_t* v = vertices + k;
	for (size_t i = 0; i < n; ++i, v += vertex_size)
	{
		const uint32_t w = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
		uint32_t d;
		if (mode == 0)
		{
			d = 0;
			for (int j = 0; j < 4; ++j)
			{
				const uint8_t delta = (uint8_t)((w >> (8 * j)) - (p >> (8 * j)));
				const uint8_t z = (uint8_t)((delta << 1) ^ (0 - (delta >> 7)));
				d |= (uint32_t)z << (8 * j);
			}
		}
		else if (mode == 1)
		{
			d = 0;
			for (int j = 0; j < 2; ++j)
			{
				const uint16_t delta = (uint16_t)((w >> (16 * j)) - (p >> (16 * j)));
				const uint16_t z = (uint16_t)((delta << 1) ^ (0 - (delta >> 15)));
				d |= (uint32_t)z << (16 * j);
			}
		}
		else
			d = rotl32(w ^ p, rot);
		planes[i] = (uint8_t)d;
		planes[stride + i] = (uint8_t)(d >> 8);
		planes[2 * stride + i] = (uint8_t)(d >> 16);
		planes[3 * stride + i] = (uint8_t)(d >> 24);
		p = w;
	}
}

// ---- per-stream channel choice (v1, levels 2 and 3; reference :359-466) ------------------------------------------------

// cost class of a byte that must be representable in a group: 0, 2, 4 or 8 bits
inline int bits_needed(uint8_t v)
{
	return v == 0 ? 0 : (v <= 3 ? 2 : (v <= 15 ? 4 : 8));
}

// rotation (0..7) that packs the bits which change inside 16-vertex groups into as few bytes as possible
int choose_rotation(const uint8_t* vertices, size_t count, size_t vertex_size, size_t k)
{
	size_t score[8] = {};
	const uint8_t* v = vertices + k;
	uint32_t last = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
	for (size_t i = 0; i < count; i += kGroup)
	{
		uint32_t changed = 0;
		for (size_t j = 0; j < kGroup && i + j < count; ++j, v += vertex_size)
		{
			const uint32_t w = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
			changed |= w ^ last;
			last = w;
		}
		for (int r = 0; r < 8; ++r)
		{
			const uint32_t x = rotl32(changed, r);
			score[r] += bits_needed((uint8_t)x) + bits_needed((uint8_t)(x >> 8)) + bits_needed((uint8_t)(x >> 16)) + bits_needed((uint8_t)(x >> 24));
		}
	}
	int best = 0;
	for (int r = 1; r < 8; ++r)
		if (score[r] < score[best])
			best = r;
	return best;
}

// channel mode of one 32-bit lane: the candidate whose planes cost least on a sample of the blocks (every third one)
int choose_channel(const uint8_t* vertices, size_t count, size_t vertex_size, size_t k, size_t block, int candidates, int rot)
{
	uint8_t planes[4 * kBlockMaxVerts];
	size_t total[3] = {};
	for (size_t first = 0; first < count; first += block * 3)
	{
		const size_t n = std::min(block, count - first);
		const size_t padded = (n + kGroup - 1) & ~(size_t)(kGroup - 1);
		uint8_t prev[4];
		memcpy(prev, vertices + (first == 0 ? 0 : first - 1) * vertex_size + k, 4);
		for (int mode = 0; mode < candidates; ++mode)
		{
			memset(planes, 0, sizeof(planes));
			delta_planes(planes, kBlockMaxVerts, vertices + first * vertex_size, n, vertex_size, k, prev, mode, rot);
			for (int j = 0; j < 4; ++j)
				for (size_t g = 0; g < padded; g += kGroup)
				{
					if (g >= n)
						break; // (whole groups of padding are not part of the sample)
					const GroupInfo s = summarise(planes + j * kBlockMaxVerts + g);
					total[mode] += std::min(std::min(s.cost(1), s.cost(2)), std::min(s.cost(4), s.cost(8)));
				}
		}
	}
	int best = 0;
	for (int mode = 1; mode < candidates; ++mode)
		if (total[mode] < total[best])
			best = mode;
	return best == 2 ? (2 | (rot << 4)) : best;
}

// ---- one block ------------------------------------------------------------------------------------------------------

const int kWidthsV0[4] = {0, 2, 4, 8};
const int kWidthsV1[5] = {0, 1, 2, 4, 8};

// v1 control mode of a plane: 2 = all zero, 3 = raw bytes, 0 / 1 = bit-packed with widths {0,1,2,4} / {1,2,4,8}
int choose_control(const GroupInfo* info, size_t groups, size_t n, int level)
{
	bool all_zero = true;
	for (size_t g = 0; g < groups; ++g)
		all_zero = all_zero && info[g].zero();
	if (all_zero)
		return 2;
	if (level == 0)
		return 1;
	const size_t header = (groups + 3) / 4;
	size_t with0 = header, with8 = header;
	for (size_t g = 0; g < groups; ++g)
	{
		const size_t c124 = std::min(std::min(info[g].cost(1), info[g].cost(2)), info[g].cost(4));
		with0 += std::min(c124, info[g].cost(0));
		with8 += std::min(c124, info[g].cost(8));
	}
	if (with0 < n || with8 < n)
		return with0 < with8 ? 0 : 1;
	return 3;
}

uint8_t* emit_block(uint8_t* out, uint8_t* out_end, const uint8_t* vertices, size_t n, size_t vertex_size, const uint8_t* prev, const uint8_t* channels, int version, int level)
{
	const size_t padded = (n + kGroup - 1) & ~(size_t)(kGroup - 1);
	const size_t groups = padded / kGroup;

	uint8_t* control = out;
	if (version != 0)
	{
		if ((size_t)(out_end - out) < vertex_size / 4)
			return nullptr;
		memset(control, 0, vertex_size / 4);
		out += vertex_size / 4;
	}

	uint8_t planes[4 * kBlockMaxVerts];
	GroupInfo info[kBlockMaxVerts / kGroup];
	for (size_t k = 0; k < vertex_size; k += 4)
	{
		const int channel = version == 0 ? 0 : channels[k / 4];
		memset(planes, 0, sizeof(planes)); // (positions beyond n inside the last group are encoded as zeros, reference :517-520)
		delta_planes(planes, kBlockMaxVerts, vertices, n, vertex_size, k, prev + k, channel & 3, channel >> 4);
		for (int j = 0; j < 4; ++j)
		{
			const uint8_t* plane = planes + j * kBlockMaxVerts;
			for (size_t g = 0; g < groups; ++g)
				info[g] = summarise(plane + g * kGroup);
			if (version == 0)
			{
				out = emit_plane(out, out_end, plane, info, groups, kWidthsV0);
			}
			else
			{
				const int ctrl = choose_control(info, groups, n, level);
				control[k / 4] |= (uint8_t)(ctrl << (2 * j));
				if (ctrl == 3)
				{
					if ((size_t)(out_end - out) < n)
						return nullptr;
					memcpy(out, plane, n);
					out += n;
				}
				else if (ctrl != 2)
					out = emit_plane(out, out_end, plane, info, groups, kWidthsV1 + ctrl);
			}
			if (!out)
				return nullptr;
		}
	}
	return out;
}

size_t encode_stream(uint8_t* buffer, size_t buffer_size, const uint8_t* vertices, size_t count, size_t vertex_size, int level, int version, unsigned int* sidecar)
{
	uint8_t* out = buffer;
	uint8_t* out_end = buffer + buffer_size;
	if (buffer_size < 1)
		return 0;
	*out++ = (uint8_t)(kMagic | version);

	uint8_t first[256] = {};
	if (count)
		memcpy(first, vertices, vertex_size);
	const size_t block = block_vertices((uint32_t)vertex_size);

	// per-lane channel modes (v1, level >= 2): byte deltas, 16-bit deltas, or (level 3) xor with a rotation
	uint8_t channels[64] = {};
	if (version != 0 && level > 1 && count > 1)
		for (size_t k = 0; k < vertex_size; k += 4)
		{
			const int rot = level >= 3 ? choose_rotation(vertices, count, vertex_size, k) : 0;
			channels[k / 4] = (uint8_t)choose_channel(vertices, count, vertex_size, k, block, level >= 3 ? 3 : 2, rot);
		}

	uint8_t prev[256];
	memcpy(prev, first, sizeof(prev));
	size_t b = 0;
	for (size_t at = 0; at < count; at += block, ++b)
	{
		const size_t n = std::min(block, count - at);
		if (sidecar)
			sidecar[b] = (unsigned int)(out - buffer);
		out = emit_block(out, out_end, vertices + at * vertex_size, n, vertex_size, prev, channels, version, level);
		if (!out)
			return 0;
		memcpy(prev, vertices + (at + n - 1) * vertex_size, vertex_size);
	}
	if (sidecar && count)
		sidecar[b] = (unsigned int)(out - buffer);

	// tail: first vertex (+ channel bytes for v1), zero-padded in front to the minimum tail size
	const size_t tail = tail_bytes((uint32_t)vertex_size, (uint32_t)version);
	const size_t padded_tail = tail_padded((uint32_t)vertex_size, (uint32_t)version);
	if ((size_t)(out_end - out) < padded_tail)
		return 0;
	memset(out, 0, padded_tail - tail);
	out += padded_tail - tail;
	memcpy(out, first, vertex_size);
	out += vertex_size;
	if (version != 0)
	{
		memcpy(out, channels, vertex_size / 4);
		out += vertex_size / 4;
	}
	return (size_t)(out - buffer);
}

bool encode_args_ok(size_t vertex_size, int level, int version)
{
	return vertex_size > 0 && vertex_size <= 256 && vertex_size % 4 == 0 && level >= 0 && level <= 9 && (version == 0 || version == 1);
}

size_t align16(size_t v)
{
	return (v + 15) & ~(size_t)15;
}

} // namespace

extern "C" size_t mob200_encode_vertex_bound(size_t vertex_count, size_t vertex_size)
{
	if (vertex_size == 0 || vertex_size > 256 || vertex_size % 4 != 0)
		return 0;
	// worst case per block: control bytes, and per byte of the vertex a selector header plus the raw plane
	const size_t block = block_vertices((uint32_t)vertex_size);
	const size_t blocks = (vertex_count + block - 1) / block;
	const size_t per_plane = (block / kGroup + 3) / 4 + block;
	const size_t tail = std::max<size_t>(vertex_size + vertex_size / 4, std::max(kTailMinV0, kTailMinV1));
	return 1 + blocks * vertex_size * (vertex_size / 4 + per_plane) + tail;
}

extern "C" size_t mob200_encode_vertex_buffer(unsigned char* buffer, size_t buffer_size, const void* vertices, size_t vertex_count, size_t vertex_size, int level, int version, unsigned int* sidecar)
{
	if (!buffer || (!vertices && vertex_count) || !encode_args_ok(vertex_size, level, version) || vertex_count >= 0xffffffffull)
		return 0;
	return encode_stream(buffer, buffer_size, static_cast<const uint8_t*>(vertices), vertex_count, vertex_size, level, version, sidecar);
}

extern "C" size_t mob200_segment_count(size_t vertex_count, size_t segment_vertices)
{
	if (segment_vertices == 0)
		return vertex_count ? 1 : 0;
	return (vertex_count + segment_vertices - 1) / segment_vertices;
}

extern "C" size_t mob200_encode_segments_bound(size_t vertex_count, size_t vertex_size, size_t segment_vertices)
{
	const size_t n = mob200_segment_count(vertex_count, segment_vertices);
	if (n == 0)
		return 0;
	const size_t seg = segment_vertices ? segment_vertices : vertex_count;
	const size_t last = vertex_count - (n - 1) * seg;
	return (n - 1) * align16(mob200_encode_vertex_bound(seg, vertex_size)) + align16(mob200_encode_vertex_bound(last, vertex_size));
}

extern "C" size_t mob200_segments_sidecar_entries(size_t vertex_count, size_t vertex_size, size_t segment_vertices)
{
	const size_t n = mob200_segment_count(vertex_count, segment_vertices);
	if (n == 0)
		return 0;
	const size_t seg = segment_vertices ? segment_vertices : vertex_count;
	const size_t last = vertex_count - (n - 1) * seg;
	return (n - 1) * mob200_sidecar_entries(seg, vertex_size) + mob200_sidecar_entries(last, vertex_size);
}

extern "C" int mob200_encode_segments(const void* vertices, size_t vertex_count, size_t vertex_size, size_t segment_vertices, int level, int version, int threads,
    unsigned char* out, size_t out_capacity, size_t* out_size, mob200_Segment* segments, size_t segment_capacity, unsigned int* sidecars, size_t sidecar_capacity)
{
	if ((!vertices && vertex_count) || !encode_args_ok(vertex_size, level, version) || !out || !segments)
		return MOB200_ERR_ARGUMENT;
	const size_t n = mob200_segment_count(vertex_count, segment_vertices);
	if (n > segment_capacity || mob200_encode_segments_bound(vertex_count, vertex_size, segment_vertices) > out_capacity)
		return MOB200_ERR_ARGUMENT;
	if (sidecars && mob200_segments_sidecar_entries(vertex_count, vertex_size, segment_vertices) > sidecar_capacity)
		return MOB200_ERR_ARGUMENT;
	if (out_size)
		*out_size = 0;
	if (n == 0)
		return 0;

	// every segment is encoded into its worst-case slot (in parallel), then the streams are moved down to 16-byte
	// boundaries in order (the slots never overlap their final places from below)
	const size_t seg = segment_vertices ? segment_vertices : vertex_count;
	if (seg >= 0xffffffffull)
		return MOB200_ERR_ARGUMENT;
	const size_t slot = align16(mob200_encode_vertex_bound(seg, vertex_size));
	const size_t side_per_seg = mob200_sidecar_entries(seg, vertex_size);
	const uint8_t* base = static_cast<const uint8_t*>(vertices);
	for (size_t i = 0; i < n; ++i)
	{
		mob200_Segment& s = segments[i];
		s.first_vertex = i * seg;
		s.vertex_count = std::min(seg, vertex_count - s.first_vertex);
		s.offset = i * slot;
		s.size = 0;
		s.sidecar_offset = i * side_per_seg;
		s.sidecar_entries = sidecars ? mob200_sidecar_entries(s.vertex_count, vertex_size) : 0;
	}

	std::atomic<size_t> next(0);
	std::atomic<int> failed(0);
	auto work = [&]() {
		for (;;)
		{
			const size_t i = next.fetch_add(1);
			if (i >= n)
				return;
			mob200_Segment& s = segments[i];
			const size_t cap = i + 1 < n ? slot : out_capacity - s.offset;
			s.size = encode_stream(out + s.offset, cap, base + s.first_vertex * vertex_size, s.vertex_count, vertex_size, level, version, sidecars ? sidecars + s.sidecar_offset : nullptr);
			if (s.size == 0)
				failed.store(1);
		}
	};
	size_t nthreads = threads > 0 ? (size_t)threads : std::max(1u, std::thread::hardware_concurrency());
	nthreads = std::min(nthreads, n);
	std::vector<std::thread> pool;
	for (size_t t = 1; t < nthreads; ++t)
		pool.emplace_back(work);
	work();
	for (std::thread& t : pool)
		t.join();
	if (failed.load())
		return MOB200_ERR_ARGUMENT;

	size_t cursor = 0;
	for (size_t i = 0; i < n; ++i)
	{
		mob200_Segment& s = segments[i];
		if (s.offset != cursor)
			memmove(out + cursor, out + s.offset, s.size);
		s.offset = cursor;
		const size_t end = cursor + s.size;
		cursor = align16(end);
		if (cursor <= out_capacity)
			memset(out + end, 0, cursor - end);
		else
			cursor = end;
	}
	if (out_size)
		*out_size = cursor;
	return (int)std::min<size_t>(n, 0x7fffffff);
}
