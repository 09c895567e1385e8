/*
 * oracle/meshletcodec_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the meshoptimizer meshlet *decoder* (reference src/meshletcodec.cpp).  It exists
 * only so that tests/ can check the CUDA meshlet kernel (meshoptimizer_b200/csrc/mob200_meshlet.cu).
 * Nothing under meshoptimizer_b200/ may include, link or call this file.
 *
 * Parity status: PINNED against the reference library (oracle/_ref/libmeshopt_ref.so, unmodified
 * src/meshletcodec.cpp) by tests/test_meshlet_cpu.py: encoder-produced meshlets of every shape the tests
 * generate (all four output formats), plus truncated and corrupted inputs for the return codes.  The
 * reference tests hold no literal known-answer vectors for this codec (demo/tests.cpp only round-trips
 * through the encoder), so the reference itself, run here, is the anchor.
 *
 * What is restated:
 *   buffer layout (data | gap | ctrl | codes)      :981-1013  -> oracle_decodeMeshlet
 *   vertex references: 4 per control byte          :324-357   -> vertex section below
 *   triangles: 4-bit codes, 3-entry triangle FIFO  :262-322   -> triangle section below
 * Error behaviour follows the x86 SIMD path, which is what the reference runs on the box: the overrun check
 * sits in front of every group of four vertices (:760-765), every PAIR of packed triangles (:627-633) and every
 * FOUR byte triangles (:683-693), and vertex slots are bytes (they wrap at 256).
 */
#include <stddef.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

static uint32_t le_bytes(const unsigned char* p, unsigned n)
{
	uint32_t v = 0;
	for (unsigned i = 0; i < n; ++i)
		v |= (uint32_t)p[i] << (8 * i);
	return v;
}

ORACLE_API int oracle_decodeMeshlet(void* vertices, size_t vertex_count, size_t vertex_size, void* triangles, size_t triangle_count, size_t triangle_size,
    const unsigned char* buffer, size_t buffer_size)
{
	/* the three sections are anchored at the END of the buffer; the data section starts at its beginning */
	size_t code_bytes = (triangle_count + 1) / 2;
	size_t ctrl_bytes = (vertex_count + 3) / 4;
	size_t gap = code_bytes + ctrl_bytes < 16 ? 16 - (code_bytes + ctrl_bytes) : 0;
	if (buffer_size < code_bytes + ctrl_bytes + gap)
		return -2;

	const unsigned char* codes = buffer + buffer_size - code_bytes;
	const unsigned char* ctrl = codes - ctrl_bytes;
	const unsigned char* limit = ctrl - gap; /* the data section must end exactly here */
	const unsigned char* data = buffer;

	/* ---- vertex references: running value, zigzag deltas of 0..4 bytes, +1 per step ---- */
	uint32_t running = 0xffffffffu;
	for (size_t first = 0; first < vertex_count; first += 4)
	{
		if (data > limit)
			return -2;
		unsigned c4 = ctrl[first / 4];
		for (unsigned k = 0; k < 4; ++k)
		{
			/* length of value k: low bit in the low nibble, high bit in the high nibble; 0xff = four 4-byte values */
			unsigned len = c4 == 0xff ? 4 : (((c4 >> k) & 1) | (((c4 >> (k + 4)) & 1) << 1));
			uint32_t z = le_bytes(data, len);
			data += len;
			uint32_t delta = (z & 1) ? ~(z >> 1) : (z >> 1);
			running += delta + 1;
			if (first + k < vertex_count)
			{
				if (vertex_size == 4)
					((uint32_t*)vertices)[first + k] = running;
				else
					((uint16_t*)vertices)[first + k] = (uint16_t)running;
			}
		}
	}

	/* ---- triangles: each corner is "next unseen slot", an explicit byte, or taken from an edge of one of the
	 *      three previous triangles ---- */
	size_t check_mask = triangle_size == 3 ? 3 : 1; /* byte triangles are decoded four at a time, packed ones in pairs */
	unsigned next_slot = 0;
	unsigned char recent[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; /* corners (a, b, c) of the last three triangles */
	for (size_t t = 0; t < triangle_count; ++t)
	{
		if ((t & check_mask) == 0 && data > limit)
			return -2;
		unsigned code = (codes[t / 2] >> ((t & 1) * 4)) & 15;
		unsigned char a, b, c;
		if (code < 12)
		{
			/* reuse: triangle `age` back, one of its two most recent edges, reversed */
			const unsigned char* r = recent[code / 4];
			if (code & 2)
				a = r[2], b = r[1]; /* edge (b, c) -> new (c, b) */
			else
				a = r[0], b = r[2]; /* edge (c, a) -> new (a, c) */
			if (code & 1)
				c = *data++;
			else
				c = (unsigned char)next_slot++;
		}
		else
		{
			/* restart: the first `code - 12` corners are explicit bytes, the others take the next unseen slots */
			unsigned explicit_corners = code - 12;
			a = explicit_corners > 0 ? *data++ : (unsigned char)next_slot++;
			b = explicit_corners > 1 ? *data++ : (unsigned char)next_slot++;
			c = explicit_corners > 2 ? *data++ : (unsigned char)next_slot++;
		}
		if (triangle_size == 4)
			((uint32_t*)triangles)[t] = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16);
		else
		{
			unsigned char* out = (unsigned char*)triangles + t * 3;
			out[0] = a, out[1] = b, out[2] = c;
		}
		for (int k = 0; k < 3; ++k)
			recent[2][k] = recent[1][k], recent[1][k] = recent[0][k];
		recent[0][0] = a, recent[0][1] = b, recent[0][2] = c;
	}

	return data == limit ? 0 : -3;
}
