/*
 * oracle/harness.cpp -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small multi-threaded driver that is compiled twice by oracle/Makefile:
 *   -DHARNESS_REF : linked with the UNMODIFIED reference sources from /root/reference/src into
 *                   oracle/_ref/libmeshopt_ref.so (calls meshopt_*), used as
 *                     - the input generator (north_star: "streams encoded by the reference encoder"),
 *                     - the parity oracle of record,
 *                     - the timed CPU baseline (cpu_baseline.kind = "reference");
 *   (no define)   : linked with oracle/vertexcodec_oracle.c + vertexfilter_oracle.c into
 *                   oracle/_build/libmeshopt_oracle.so (calls oracle_*), the "port" baseline.
 *
 * Timing follows tools/codecbench.cpp:82-105 (one decode call per stream, GB/s of decoded bytes,
 * best of N passes), extended with a std::thread pool pulling whole streams from an atomic counter
 * as BASELINE.md section 3 specifies.  Synthetic data generators restate tools/codecbench.cpp:47-56,
 * 387-417 (grid) and js/benchmark.js:16-30 (16-byte stream); see SURVEY.md Appendix D.
 */
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define HARNESS_API extern "C" __attribute__((visibility("default")))

extern "C"
{
#ifdef HARNESS_REF
	int meshopt_decodeVertexBuffer(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size);
	void meshopt_decodeFilterOct(void* buffer, size_t count, size_t stride);
	void meshopt_decodeFilterQuat(void* buffer, size_t count, size_t stride);
	void meshopt_decodeFilterExp(void* buffer, size_t count, size_t stride);
	void meshopt_decodeFilterColor(void* buffer, size_t count, size_t stride);
	size_t meshopt_encodeVertexBufferLevel(unsigned char* buffer, size_t buffer_size, const void* vertices, size_t vertex_count, size_t vertex_size, int level, int version);
	size_t meshopt_encodeVertexBufferBound(size_t vertex_count, size_t vertex_size);
	void meshopt_optimizeVertexCache(unsigned int* destination, const unsigned int* indices, size_t index_count, size_t vertex_count);
	size_t meshopt_optimizeVertexFetch(void* destination, unsigned int* indices, size_t index_count, const void* vertices, size_t vertex_count, size_t vertex_size);
	int meshopt_decodeIndexBuffer(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);
	int meshopt_decodeIndexSequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);
	int meshopt_decodeMeshlet(void* vertices, size_t vertex_count, size_t vertex_size, void* triangles, size_t triangle_count, size_t triangle_size, const unsigned char* buffer, size_t buffer_size);
#define DECODE_MESHLET meshopt_decodeMeshlet
#define DECODE meshopt_decodeVertexBuffer
#define DECODE_TRI meshopt_decodeIndexBuffer
#define DECODE_SEQ meshopt_decodeIndexSequence
#define F_OCT meshopt_decodeFilterOct
#define F_QUAT meshopt_decodeFilterQuat
#define F_EXP meshopt_decodeFilterExp
#define F_COLOR meshopt_decodeFilterColor
#else
	int oracle_decodeVertexBuffer(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size);
	void oracle_decodeFilterOct(void* buffer, size_t count, size_t stride);
	void oracle_decodeFilterQuat(void* buffer, size_t count, size_t stride);
	void oracle_decodeFilterExp(void* buffer, size_t count, size_t stride);
	void oracle_decodeFilterColor(void* buffer, size_t count, size_t stride);
	int oracle_decodeIndexBuffer(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);
	int oracle_decodeIndexSequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);
	int oracle_decodeMeshlet(void* vertices, size_t vertex_count, size_t vertex_size, void* triangles, size_t triangle_count, size_t triangle_size, const unsigned char* buffer, size_t buffer_size);
#define DECODE_MESHLET oracle_decodeMeshlet
#define DECODE oracle_decodeVertexBuffer
#define DECODE_TRI oracle_decodeIndexBuffer
#define DECODE_SEQ oracle_decodeIndexSequence
#define F_OCT oracle_decodeFilterOct
#define F_QUAT oracle_decodeFilterQuat
#define F_EXP oracle_decodeFilterExp
#define F_COLOR oracle_decodeFilterColor
#endif
}

/* one independent stream, mirroring what gltf/parsegltf.cpp:561-627 hands the decoder per bufferView */
struct HarnessStream
{
	const unsigned char* src;
	size_t src_size;
	void* dst;
	size_t vertex_count;
	size_t vertex_size;
	int filter; /* 0 none, 1 oct, 2 quat, 3 exp, 4 color; 16 / 17: an INDEX stream (triangle list / sequence),
	               vertex_count = index count, vertex_size = index size */
	int status; /* out: return code of the decode call */
};

static void decode_one(HarnessStream& s)
{
	if (s.filter >= 16)
	{
		s.status = s.filter == 16 ? DECODE_TRI(s.dst, s.vertex_count, s.vertex_size, s.src, s.src_size) : DECODE_SEQ(s.dst, s.vertex_count, s.vertex_size, s.src, s.src_size);
		return;
	}
	s.status = DECODE(s.dst, s.vertex_count, s.vertex_size, s.src, s.src_size);
	if (s.status != 0)
		return;
	switch (s.filter)
	{
	case 1: F_OCT(s.dst, s.vertex_count, s.vertex_size); break;
	case 2: F_QUAT(s.dst, s.vertex_count, s.vertex_size); break;
	case 3: F_EXP(s.dst, s.vertex_count, s.vertex_size); break;
	case 4: F_COLOR(s.dst, s.vertex_count, s.vertex_size); break;
	default: break;
	}
}

template <typename Fn>
static void parallel_for(size_t n, int threads, Fn fn)
{
	if (threads <= 1 || n <= 1)
	{
		for (size_t i = 0; i < n; ++i)
			fn(i);
		return;
	}
	std::atomic<size_t> next(0);
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; ++t)
		pool.emplace_back([&]() {
			for (;;)
			{
				size_t i = next.fetch_add(1, std::memory_order_relaxed);
				if (i >= n)
					break;
				fn(i);
			}
		});
	for (auto& th : pool)
		th.join();
}

/* Decode every stream `passes` times with `threads` host threads; returns the best pass in seconds
 * (and the per-pass times in seconds_out[passes] when non-NULL). */
HARNESS_API double harness_decode_mt(HarnessStream* streams, size_t n, int threads, int passes, double* seconds_out)
{
	double best = 1e30;
	for (int p = 0; p < passes; ++p)
	{
		auto t0 = std::chrono::steady_clock::now();
		parallel_for(n, threads, [&](size_t i) { decode_one(streams[i]); });
		auto t1 = std::chrono::steady_clock::now();
		double s = std::chrono::duration<double>(t1 - t0).count();
		if (seconds_out)
			seconds_out[p] = s;
		if (s < best)
			best = s;
	}
	return best;
}

/* meshlets: same pool, one decode call per meshlet */
struct HarnessMeshlet
{
	const unsigned char* src;
	size_t src_size;
	void* vertices;
	size_t vertex_count, vertex_size;
	void* triangles;
	size_t triangle_count, triangle_size;
	int status;
};

HARNESS_API double harness_decode_meshlets_mt(HarnessMeshlet* m, size_t n, int threads, int passes)
{
	double best = 1e30;
	for (int p = 0; p < passes; ++p)
	{
		auto t0 = std::chrono::steady_clock::now();
		parallel_for(n, threads, [&](size_t i) {
			m[i].status = DECODE_MESHLET(m[i].vertices, m[i].vertex_count, m[i].vertex_size, m[i].triangles, m[i].triangle_count, m[i].triangle_size, m[i].src, m[i].src_size);
		});
		auto t1 = std::chrono::steady_clock::now();
		double s = std::chrono::duration<double>(t1 - t0).count();
		if (s < best)
			best = s;
	}
	return best;
}

HARNESS_API int harness_hw_threads(void)
{
	unsigned n = std::thread::hardware_concurrency();
	return n ? (int)n : 1;
}

/* ---------------------------------------------------------------------------------------------
 * synthetic vertex data (no reference code needed)
 * ------------------------------------------------------------------------------------------- */
static inline uint32_t murmur3_fmix(uint32_t h)
{
	h ^= h >> 16;
	h *= 0x85ebca6bu;
	h ^= h >> 13;
	h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}

/* keep the top `keep` mantissa bits of a float (round to nearest, zero/denormal -> 0, inf/nan kept):
 * same effect as the reference's meshopt_quantizeFloat (src/quantization.cpp:37-58); only used to
 * make C2's float words compressible */
static inline uint32_t keep_mantissa_bits(float v, int keep)
{
	uint32_t u;
	memcpy(&u, &v, 4);
	uint32_t expo = u & 0x7f800000u;
	if (expo == 0)
		return 0;
	if (expo == 0x7f800000u || keep >= 23)
		return u;
	unsigned drop = 23u - (unsigned)keep;
	return (u + (1u << (drop - 1))) & ~((1u << drop) - 1u);
}

/* C1a vertex data: (side+1)^2 grid of 32-byte vertices, word k = murmur3(id*16+k) & ((1<<(k+1))-1).
 * Written in grid order; the caller applies the vertex cache/fetch reorder (reference build only). */
HARNESS_API void harness_gen_grid(uint16_t* out, int side)
{
	size_t n = (size_t)(side + 1) * (side + 1);
	for (size_t id = 0; id < n; ++id)
		for (int k = 0; k < 16; ++k)
			out[id * 16 + k] = (uint16_t)(murmur3_fmix((uint32_t)(id * 16 + k)) & ((1u << (k + 1)) - 1));
}

/* C1b: js/benchmark.js stream, 16 bytes per vertex, Lehmer generator advanced on every byte */
HARNESS_API void harness_gen_js16(uint8_t* out, size_t vertex_count)
{
	uint64_t lcg = 1;
	size_t total = vertex_count * 16;
	for (size_t i = 0; i < total; ++i)
	{
		lcg = lcg * 48271 % 2147483647;
		unsigned k = (unsigned)(i % 16);
		out[i] = k <= 8 ? (uint8_t)(lcg & ((1u << k) - 1)) : (uint8_t)(i & ((1u << (k - 8)) - 1));
	}
}

/* C2: 32-byte vertices [first, first+count): words 0-11 as the grid, words 12-13 a 32-bit counter
 * i*3, words 14-15 the float sin(i*1e-3) with 12 mantissa bits. */
HARNESS_API void harness_gen_c2(uint16_t* out, uint64_t first, size_t count, int threads)
{
	size_t chunk = 65536;
	size_t chunks = (count + chunk - 1) / chunk;
	parallel_for(chunks, threads, [&](size_t c) {
		size_t lo = c * chunk, hi = lo + chunk < count ? lo + chunk : count;
		for (size_t j = lo; j < hi; ++j)
		{
			uint64_t i = first + j;
			uint16_t* v = out + j * 16;
			for (int k = 0; k < 12; ++k)
				v[k] = (uint16_t)(murmur3_fmix((uint32_t)(i * 16 + k)) & ((1u << (k + 1)) - 1));
			uint32_t counter = (uint32_t)(i * 3);
			v[12] = (uint16_t)counter;
			v[13] = (uint16_t)(counter >> 16);
			uint32_t fb = keep_mantissa_bits(sinf((float)i * 1e-3f), 12);
			v[14] = (uint16_t)fb;
			v[15] = (uint16_t)(fb >> 16);
		}
	});
}

#ifdef HARNESS_REF
/* ---------------------------------------------------------------------------------------------
 * reference encoder front-end (input generation only)
 * ------------------------------------------------------------------------------------------- */

/* Encode n independent segments of one vertex array in parallel.  Segment i covers vertices
 * [first[i], first[i]+count[i]).  Output i is written at out + out_offset[i] (capacity out_cap[i]);
 * sizes_out[i] receives the encoded size (0 on failure).  version is passed explicitly so the
 * reference's global (src/vertexcodec.cpp:125) is never touched from several threads. */
HARNESS_API void harness_encode_segments(const unsigned char* vertices, size_t vertex_size, const uint64_t* first, const uint64_t* count,
    size_t n, int level, int version, unsigned char* out, const uint64_t* out_offset, const uint64_t* out_cap, uint64_t* sizes_out, int threads)
{
	parallel_for(n, threads, [&](size_t i) {
		sizes_out[i] = meshopt_encodeVertexBufferLevel(out + out_offset[i], (size_t)out_cap[i], vertices + first[i] * vertex_size, (size_t)count[i], vertex_size, level, version);
	});
}

HARNESS_API size_t harness_encode_bound(size_t vertex_count, size_t vertex_size)
{
	return meshopt_encodeVertexBufferBound(vertex_count, vertex_size);
}

/* codecbench's reorder of the grid (tools/codecbench.cpp:69-71): two triangles per cell, vertex
 * cache optimisation then vertex fetch optimisation; `vertices` is rewritten in fetch order. */
HARNESS_API void harness_grid_reorder(uint16_t* vertices, int side)
{
	size_t n = (size_t)(side + 1) * (side + 1);
	std::vector<unsigned int> indices;
	indices.reserve((size_t)side * side * 6);
	for (int x = 0; x < side; ++x)
		for (int y = 0; y < side; ++y)
		{
			unsigned a = (unsigned)((x + 0) * (side + 1) + (y + 0));
			unsigned b = (unsigned)((x + 1) * (side + 1) + (y + 0));
			unsigned c = (unsigned)((x + 0) * (side + 1) + (y + 1));
			unsigned d = (unsigned)((x + 1) * (side + 1) + (y + 1));
			indices.push_back(a);
			indices.push_back(b);
			indices.push_back(c);
			indices.push_back(c);
			indices.push_back(b);
			indices.push_back(d);
		}
	std::vector<unsigned int> ib(indices.size());
	meshopt_optimizeVertexCache(ib.data(), indices.data(), indices.size(), n);
	std::vector<uint16_t> vb(n * 16);
	meshopt_optimizeVertexFetch(vb.data(), ib.data(), ib.size(), vertices, n, 32);
	memcpy(vertices, vb.data(), n * 32);
}
#endif
