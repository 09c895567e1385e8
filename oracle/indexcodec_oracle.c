/*
 * oracle/indexcodec_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the meshoptimizer index-stream *decoders* (triangle lists and index sequences,
 * wire versions 0 and 1).  It exists only so that tests/ can check the CUDA index kernels
 * (meshoptimizer_b200/csrc/mob200_index.cu) index for index.  Nothing under meshoptimizer_b200/ may
 * include, link or call this file.
 *
 * Parity status: PINNED.  tests/test_index_cpu.py checks this file against
 *   - the index known-answer vectors of the reference tests (demo/tests.cpp: kIndexBuffer,
 *     kIndexDataV0, kIndexDataV1, kIndexSequence, kIndexSequenceV1 and the decode tests that use them),
 *   - the return-code tests (truncated / malformed / version mismatch),
 *   - the reference decoders themselves (oracle/_ref/libmeshopt_ref.so) on encoder-produced streams.
 *
 * What is restated (reference = /root/reference/src/indexcodec.cpp):
 *   varint / zigzag delta                  :95-136   -> take_varint, take_delta
 *   triangle decoder state machine         :384-576  -> oracle_decodeIndexBuffer
 *   sequence decoder                       :647-703  -> oracle_decodeIndexSequence
 *   header probe                           :364-382  -> oracle_decodeIndexVersion
 *
 * Formulation: one struct holds the whole decoder state (two ring buffers addressed by age, the "next
 * unseen vertex" counter, the last explicitly coded index, the data cursor); each triangle code is first
 * classified (edge reuse / table / explicit), then its three corners are resolved through one helper.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum
{
	TRI_MAGIC = 0xe0,
	SEQ_MAGIC = 0xd0,
	INDEX_MAX_VERSION = 1,
	RING = 16
};

typedef struct
{
	uint32_t vertex[RING];  /* recently seen vertices */
	uint32_t edge[RING][2]; /* recently emitted edges */
	unsigned vpos, epos;    /* next write position of each ring */
	uint32_t next;          /* first vertex id that has not been seen yet */
	uint32_t last;          /* last index that was coded explicitly (delta base) */
	const unsigned char* cursor;
} TriState;

static uint32_t take_varint(const unsigned char** cursor)
{
	const unsigned char* p = *cursor;
	uint32_t value = 0;
	unsigned shift = 0;
	int groups = 0;
	for (;;)
	{
		unsigned char byte = *p++;
		value |= (uint32_t)(byte & 0x7f) << shift;
		shift += 7;
		groups++;
		if (byte < 0x80 || groups == 5) /* at most five groups, whatever the data says */
			break;
	}
	*cursor = p;
	return value;
}

static uint32_t take_delta(const unsigned char** cursor, uint32_t base)
{
	uint32_t z = take_varint(cursor);
	uint32_t delta = (z & 1) ? ~(z >> 1) : (z >> 1);
	return base + delta;
}

static uint32_t vertex_by_age(const TriState* s, unsigned age)
{
	return s->vertex[(s->vpos - age) & (RING - 1)];
}

static void remember_vertex(TriState* s, uint32_t v, int advance)
{
	s->vertex[s->vpos] = v; /* written even when the ring does not advance */
	if (advance)
		s->vpos = (s->vpos + 1) & (RING - 1);
}

static void remember_edge(TriState* s, uint32_t from, uint32_t to)
{
	s->edge[s->epos][0] = from;
	s->edge[s->epos][1] = to;
	s->epos = (s->epos + 1) & (RING - 1);
}

static void put_triangle(void* destination, size_t index_size, size_t tri, uint32_t a, uint32_t b, uint32_t c)
{
	if (index_size == 2)
	{
		uint16_t* out = (uint16_t*)destination + tri * 3;
		out[0] = (uint16_t)a, out[1] = (uint16_t)b, out[2] = (uint16_t)c;
	}
	else
	{
		uint32_t* out = (uint32_t*)destination + tri * 3;
		out[0] = a, out[1] = b, out[2] = c;
	}
}

ORACLE_API int oracle_decodeIndexVersion(const unsigned char* buffer, size_t buffer_size)
{
	if (buffer_size == 0)
		return -1;
	unsigned magic = buffer[0] & 0xf0, version = buffer[0] & 0x0f;
	if (magic != TRI_MAGIC && magic != SEQ_MAGIC)
		return -1;
	return version <= INDEX_MAX_VERSION ? (int)version : -1;
}

ORACLE_API int oracle_decodeIndexBuffer(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size)
{
	size_t triangles = index_count / 3;

	/* header byte + one code byte per triangle + 16-byte table at the very end */
	if (buffer_size < 1 + triangles + 16)
		return -2;
	if ((buffer[0] & 0xf0) != TRI_MAGIC)
		return -1;
	unsigned version = buffer[0] & 0x0f;
	if (version > INDEX_MAX_VERSION)
		return -1;

	/* codes 13 and 14 of the low nibble mean "last -/+ 1" from version 1 on */
	unsigned first_explicit = version >= 1 ? 13 : 15;

	const unsigned char* codes = buffer + 1;
	const unsigned char* table = buffer + buffer_size - 16; /* also the limit of the data section */

	TriState s;
	memset(s.vertex, 0xff, sizeof(s.vertex));
	memset(s.edge, 0xff, sizeof(s.edge));
	s.vpos = s.epos = 0;
	s.next = s.last = 0;
	s.cursor = codes + triangles;

	for (size_t t = 0; t < triangles; ++t)
	{
		unsigned code = codes[t];
		uint32_t a, b, c;

		if (code < 0xf0)
		{
			/* two corners come from a remembered edge */
			unsigned edge_age = code >> 4, third = code & 15;
			const uint32_t* e = s.edge[(s.epos - 1 - edge_age) & (RING - 1)];
			a = e[0];
			b = e[1];

			if (third < first_explicit)
			{
				/* third corner: a brand-new vertex (0) or a remembered one (age third) */
				uint32_t remembered = vertex_by_age(&s, 1 + third);
				if (third == 0)
				{
					c = s.next++;
					remember_vertex(&s, c, 1);
				}
				else
				{
					c = remembered;
					remember_vertex(&s, c, 0);
				}
			}
			else
			{
				if (s.cursor > table)
					return -2;
				if (third == 15)
					c = take_delta(&s.cursor, s.last);
				else
					c = s.last + (third == 13 ? (uint32_t)-1 : 1u);
				s.last = c;
				remember_vertex(&s, c, 1);
			}
			remember_edge(&s, c, b);
			remember_edge(&s, a, c);
		}
		else
		{
			/* no edge is reused: every corner is new (0), remembered (1..14) or explicit (15) */
			unsigned fa, fb, fc;
			int explicit_form = code >= 0xfe;
			if (!explicit_form)
			{
				unsigned aux = table[code & 15];
				fa = 0;
				fb = aux >> 4;
				fc = aux & 15;
			}
			else
			{
				if (s.cursor > table)
					return -2;
				unsigned aux = *s.cursor++;
				fa = code == 0xfe ? 0 : 15;
				fb = aux >> 4;
				fc = aux & 15;
				if (aux == 0)
					s.next = 0; /* restart of the vertex numbering */
			}

			/* the remembered corners refer to the ring as it is BEFORE this triangle is added */
			uint32_t rb = vertex_by_age(&s, fb), rc = vertex_by_age(&s, fc);
			int adv_b, adv_c;
			if (!explicit_form)
			{
				a = s.next++;
				b = fb == 0 ? s.next : rb;
				s.next += fb == 0;
				c = fc == 0 ? s.next : rc;
				s.next += fc == 0;
				adv_b = fb == 0;
				adv_c = fc == 0;
			}
			else
			{
				a = fa == 0 ? s.next++ : 0;
				b = fb == 0 ? s.next++ : rb;
				c = fc == 0 ? s.next++ : rc;
				if (fa == 15)
					s.last = a = take_delta(&s.cursor, s.last);
				if (fb == 15)
					s.last = b = take_delta(&s.cursor, s.last);
				if (fc == 15)
					s.last = c = take_delta(&s.cursor, s.last);
				adv_b = fb == 0 || fb == 15;
				adv_c = fc == 0 || fc == 15;
			}
			remember_vertex(&s, a, 1);
			remember_vertex(&s, b, adv_b);
			remember_vertex(&s, c, adv_c);
			remember_edge(&s, b, a);
			remember_edge(&s, c, b);
			remember_edge(&s, a, c);
		}
		put_triangle(destination, index_size, t, a, b, c);
	}

	return s.cursor == table ? 0 : -3;
}

ORACLE_API int oracle_decodeIndexSequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size)
{
	/* header byte + at least one byte per index + 4-byte tail */
	if (buffer_size < 1 + index_count + 4)
		return -2;
	if ((buffer[0] & 0xf0) != SEQ_MAGIC)
		return -1;
	if ((unsigned)(buffer[0] & 0x0f) > INDEX_MAX_VERSION)
		return -1;

	const unsigned char* cursor = buffer + 1;
	const unsigned char* limit = buffer + buffer_size - 4;
	uint32_t base[2] = {0, 0};

	for (size_t i = 0; i < index_count; ++i)
	{
		if (cursor >= limit)
			return -2;
		uint32_t z = take_varint(&cursor);
		unsigned which = z & 1; /* which of the two running baselines this index continues */
		z >>= 1;
		uint32_t delta = (z & 1) ? ~(z >> 1) : (z >> 1);
		uint32_t index = base[which] + delta;
		base[which] = index;
		if (index_size == 2)
			((uint16_t*)destination)[i] = (uint16_t)index;
		else
			((uint32_t*)destination)[i] = index;
	}
	return cursor == limit ? 0 : -3;
}
