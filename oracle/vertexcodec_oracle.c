/*
 * oracle/vertexcodec_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the meshoptimizer vertex-buffer *decoder*
 * (wire formats v0 and v1).  It exists only so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py can check the CUDA path byte for byte.  Nothing under
 * meshoptimizer_b200/ may include, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file against
 *   - every codec known-answer vector the reference tests hold
 *     (reference demo/tests.cpp:45-78,403-499,626-649 and js/meshopt_decoder.test.js:10-136),
 *   - the return-code tests (demo/tests.cpp:521-572,743-751),
 *   - the reference decoder itself (oracle/_ref/libmeshopt_ref.so, built from the reference
 *     sources by oracle/Makefile) on encoder-produced streams, when that library is present.
 *
 * What is restated (reference = /root/reference/src/vertexcodec.cpp):
 *   format constants                      :123-138
 *   block size rule                       :140-147   -> oracle_block_size
 *   16-value group unpack                 :582-641   -> unpack_group
 *   channel (byte plane) decode           :643-667,1369-1425 -> decode_plane
 *   delta reconstruction                  :669-699   -> undelta_quad
 *   block decode incl. v1 control modes   :701-780,1513-1595 -> decode_block
 *   stream framing / error codes          :1782-1872 -> oracle_decodeVertexBuffer / ...Version
 *
 * Error behaviour follows the x86 SIMD path (that is what tools/codecbench runs): a literal channel
 * needs the 16-aligned vertex count to be available (:1549), a packed group needs 24 readable bytes
 * (:1385,1415).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum
{
	MAGIC = 0xa0,       /* high nibble of byte 0 */
	MAX_VERSION = 1,
	BLOCK_BYTES = 8192, /* decoded bytes per block, upper bound */
	BLOCK_MAX = 256,    /* vertices per block, upper bound */
	GROUP = 16,         /* values per group */
	GROUP_READ = 24,    /* bytes that must be readable in front of every packed group */
	TAIL_MIN_V0 = 32,
	TAIL_MIN_V1 = 24
};

ORACLE_API size_t oracle_block_size(size_t vertex_size)
{
	size_t n = (BLOCK_BYTES / vertex_size) & ~(size_t)(GROUP - 1);
	return n < BLOCK_MAX ? n : BLOCK_MAX;
}

/* width in bits of a group, from the stream version, the v1 per-channel control value and the 2-bit
 * group selector: v0 uses {0,2,4,8}; v1 uses a window of {0,1,2,4,8} that starts at `ctrl` (0 or 1) */
static int group_width(int version, int ctrl, int sel)
{
	static const int widths[5] = {0, 1, 2, 4, 8};
	if (version == 0)
		return sel == 0 ? 0 : (1 << sel);
	return widths[ctrl + sel];
}

/* Unpack one group of 16 delta bytes of `width` bits each from src; returns bytes consumed. */
static size_t unpack_group(const uint8_t* src, uint8_t out[GROUP], int width)
{
	if (width == 0)
	{
		memset(out, 0, GROUP);
		return 0;
	}
	if (width == 8)
	{
		memcpy(out, src, GROUP);
		return GROUP;
	}

	const unsigned sentinel = (1u << width) - 1;
	const size_t fixed = (size_t)(GROUP * width / 8);
	const uint8_t* escape = src + fixed;

	for (int i = 0; i < GROUP; ++i)
	{
		unsigned field;
		if (width == 1)
		{
			/* 1-bit fields are stored least-significant bit first */
			field = (src[i >> 3] >> (i & 7)) & 1u;
		}
		else
		{
			/* 2- and 4-bit fields are stored most-significant field first */
			int per_byte = 8 / width;
			int slot = i % per_byte;
			field = (src[i / per_byte] >> (8 - width - width * slot)) & sentinel;
		}
		out[i] = (uint8_t)(field == sentinel ? *escape++ : field);
	}
	return (size_t)(escape - src);
}

/* Decode one bit-packed byte plane (one byte-channel of every vertex in the block): `count` is a
 * multiple of 16.  Returns the number of bytes consumed, or (size_t)-1 when the input is too short. */
static size_t decode_plane(const uint8_t* src, size_t avail, uint8_t* plane, size_t count, int version, int ctrl)
{
	size_t groups = count / GROUP;
	size_t header = (groups + 3) / 4;
	if (avail < header)
		return (size_t)-1;

	size_t pos = header;
	for (size_t g = 0; g < groups; ++g)
	{
		if (avail - pos < GROUP_READ)
			return (size_t)-1;
		int sel = (src[g / 4] >> ((g % 4) * 2)) & 3;
		pos += unpack_group(src + pos, plane + g * GROUP, group_width(version, ctrl, sel));
	}
	return pos;
}

static uint32_t rotl32(uint32_t v, unsigned r)
{
	r &= 31;
	return r ? (v << r) | (v >> (32 - r)) : v;
}

/* Turn four delta planes (byte k..k+3 of every vertex) into vertex bytes.  `mode` is the low two
 * bits of the v1 channel byte (0 for v0): 0 = four byte lanes, 1 = two little-endian 16-bit lanes,
 * 2 = one 32-bit lane with xor and a rotation.  prev[] carries the previous vertex and is updated
 * to vertex count-1 (the caller passes the real, unpadded vertex count). */
static void undelta_quad(const uint8_t* planes, size_t plane_stride, uint8_t* dst, size_t dst_stride, size_t count, uint8_t prev[4], int mode, unsigned rot)
{
	uint32_t p = (uint32_t)prev[0] | ((uint32_t)prev[1] << 8) | ((uint32_t)prev[2] << 16) | ((uint32_t)prev[3] << 24);

	for (size_t i = 0; i < count; ++i)
	{
		uint32_t d = (uint32_t)planes[i] | ((uint32_t)planes[plane_stride + i] << 8) |
		             ((uint32_t)planes[2 * plane_stride + i] << 16) | ((uint32_t)planes[3 * plane_stride + i] << 24);
		uint32_t v;

		if (mode == 0)
		{
			v = 0;
			for (int b = 0; b < 4; ++b)
			{
				uint8_t z = (uint8_t)(d >> (8 * b));
				uint8_t u = (uint8_t)((z >> 1) ^ (uint8_t)(0 - (z & 1)));
				uint8_t s = (uint8_t)(u + (uint8_t)(p >> (8 * b)));
				v |= (uint32_t)s << (8 * b);
			}
		}
		else if (mode == 1)
		{
			v = 0;
			for (int h = 0; h < 2; ++h)
			{
				uint16_t z = (uint16_t)(d >> (16 * h));
				uint16_t u = (uint16_t)((z >> 1) ^ (uint16_t)(0 - (z & 1)));
				uint16_t s = (uint16_t)(u + (uint16_t)(p >> (16 * h)));
				v |= (uint32_t)s << (16 * h);
			}
		}
		else
		{
			v = rotl32(d, rot) ^ p;
		}

		uint8_t* o = dst + i * dst_stride;
		o[0] = (uint8_t)v;
		o[1] = (uint8_t)(v >> 8);
		o[2] = (uint8_t)(v >> 16);
		o[3] = (uint8_t)(v >> 24);
		p = v;
	}

	prev[0] = (uint8_t)p;
	prev[1] = (uint8_t)(p >> 8);
	prev[2] = (uint8_t)(p >> 16);
	prev[3] = (uint8_t)(p >> 24);
}

/* Decode one block of `count` vertices.  Returns bytes consumed or (size_t)-1 on malformed input. */
static size_t decode_block(const uint8_t* src, size_t avail, uint8_t* dst, size_t count, size_t vertex_size, uint8_t* prev, const uint8_t* channels, int version)
{
	uint8_t planes[4][BLOCK_MAX];
	size_t padded = (count + GROUP - 1) & ~(size_t)(GROUP - 1);
	size_t pos = 0;

	const uint8_t* control = src;
	if (version != 0)
	{
		if (avail < vertex_size / 4)
			return (size_t)-1;
		pos = vertex_size / 4;
	}

	for (size_t k = 0; k < vertex_size; k += 4)
	{
		unsigned cbyte = version == 0 ? 0u : control[k / 4];

		for (int j = 0; j < 4; ++j)
		{
			int ctrl = (cbyte >> (2 * j)) & 3;

			if (ctrl == 3)
			{
				/* raw bytes, one per real vertex; the SIMD reference insists on the padded count being readable */
				if (avail - pos < padded)
					return (size_t)-1;
				memcpy(planes[j], src + pos, count);
				memset(planes[j] + count, 0, padded - count);
				pos += count;
			}
			else if (ctrl == 2)
			{
				memset(planes[j], 0, padded);
			}
			else
			{
				size_t used = decode_plane(src + pos, avail - pos, planes[j], padded, version, ctrl);
				if (used == (size_t)-1)
					return (size_t)-1;
				pos += used;
			}
		}

		unsigned channel = version == 0 ? 0u : channels[k / 4];
		int mode = channel & 3;
		if (mode == 3)
			return (size_t)-1;

		undelta_quad(planes[0], BLOCK_MAX, dst + k, vertex_size, count, prev + k, mode, (32 - (channel >> 4)) & 31);
	}

	return pos;
}

ORACLE_API int oracle_decodeVertexVersion(const unsigned char* buffer, size_t buffer_size)
{
	if (buffer_size < 1)
		return -1;
	if ((buffer[0] & 0xf0) != MAGIC)
		return -1;
	int version = buffer[0] & 0x0f;
	return version > MAX_VERSION ? -1 : version;
}

/* Same contract as meshopt_decodeVertexBuffer (reference src/meshoptimizer.h:396):
 * 0 ok, -1 bad magic/version, -2 truncated or malformed, -3 trailing bytes do not match the tail. */
ORACLE_API int oracle_decodeVertexBuffer(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size)
{
	if (vertex_size == 0 || vertex_size > 256 || vertex_size % 4 != 0)
		return -4; /* the reference asserts; the oracle reports */

	if (buffer_size < 1)
		return -2;
	if ((buffer[0] & 0xf0) != MAGIC)
		return -1;
	int version = buffer[0] & 0x0f;
	if (version > MAX_VERSION)
		return -1;

	size_t tail = vertex_size + (version == 0 ? 0 : vertex_size / 4);
	size_t tail_min = version == 0 ? TAIL_MIN_V0 : TAIL_MIN_V1;
	size_t tail_padded = tail < tail_min ? tail_min : tail;

	if (buffer_size - 1 < tail_padded)
		return -2;

	const uint8_t* tail_ptr = buffer + buffer_size - tail;
	uint8_t prev[256];
	memcpy(prev, tail_ptr, vertex_size);
	const uint8_t* channels = version == 0 ? NULL : tail_ptr + vertex_size;

	size_t block = oracle_block_size(vertex_size);
	size_t pos = 1;
	uint8_t* out = (uint8_t*)destination;

	for (size_t first = 0; first < vertex_count; first += block)
	{
		size_t n = vertex_count - first < block ? vertex_count - first : block;
		size_t used = decode_block(buffer + pos, buffer_size - pos, out + first * vertex_size, n, vertex_size, prev, channels, version);
		if (used == (size_t)-1)
			return -2;
		pos += used;
	}

	return buffer_size - pos == tail_padded ? 0 : -3;
}

/* Block-offset table of a stream ("sidecar", include/meshopt_b200.h section 2b): offsets[b] = byte offset of block
 * b for b < nblocks, offsets[nblocks] = end of the last block.  Walks exactly as oracle_decodeVertexBuffer does
 * (the blocks are decoded into scratch).  Returns the decoder's return code; `offsets` needs nblocks + 1 entries. */
ORACLE_API int oracle_vertexBlockOffsets(unsigned int* offsets, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size)
{
	if (vertex_size == 0 || vertex_size > 256 || vertex_size % 4 != 0)
		return -4;
	if (buffer_size < 1)
		return -2;
	if ((buffer[0] & 0xf0) != MAGIC)
		return -1;
	int version = buffer[0] & 0x0f;
	if (version > MAX_VERSION)
		return -1;

	size_t tail = vertex_size + (version == 0 ? 0 : vertex_size / 4);
	size_t tail_min = version == 0 ? TAIL_MIN_V0 : TAIL_MIN_V1;
	size_t tail_padded = tail < tail_min ? tail_min : tail;
	if (buffer_size - 1 < tail_padded)
		return -2;

	const uint8_t* tail_ptr = buffer + buffer_size - tail;
	uint8_t prev[256];
	memcpy(prev, tail_ptr, vertex_size);
	const uint8_t* channels = version == 0 ? NULL : tail_ptr + vertex_size;

	static __thread uint8_t scratch[BLOCK_BYTES];
	size_t block = oracle_block_size(vertex_size);
	size_t pos = 1, b = 0;
	for (size_t first = 0; first < vertex_count; first += block, ++b)
	{
		size_t n = vertex_count - first < block ? vertex_count - first : block;
		offsets[b] = (unsigned int)pos;
		size_t used = decode_block(buffer + pos, buffer_size - pos, scratch, n, vertex_size, prev, channels, version);
		if (used == (size_t)-1)
			return -2;
		pos += used;
	}
	if (vertex_count)
		offsets[b] = (unsigned int)pos;
	return buffer_size - pos == tail_padded ? 0 : -3;
}
