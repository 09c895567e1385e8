// mob200_meshlet.cu -- meshlet decode on the device (reference src/meshletcodec.cpp:262-395,981-1051:
// meshopt_decodeMeshlet / meshopt_decodeMeshletRaw).
//
// A meshlet is at most 256 vertex references and 256 triangles; an asset holds hundreds of thousands of them,
// each encoded on its own.  One THREAD decodes one meshlet: the vertex references are a running value with
// zigzag deltas of 0-4 bytes, the triangles a 4-bit code per triangle over a FIFO of the last three triangles
// (kept in three registers in the packed form c|a|b|c, so that either reusable edge is one shift away).  Return
// codes follow the reference's x86 path: the overrun check sits in front of every four vertices (:760-765), every
// pair of packed triangles (:627-633) and every four byte triangles (:683-693); -2 for an overrun, -3 when the data
// section does not end at its boundary.
#include "mob200_host.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace mob200
{

constexpr int kMeshletThreads = 128;

struct DevMeshlet
{
	const uint8_t* src;
	uint8_t* vertices;
	uint8_t* triangles;
	uint32_t src_size;
	uint16_t vertex_count, triangle_count;
	uint8_t vertex_size, triangle_size; // 2|4, 3|4
};

__device__ int decode_meshlet(const DevMeshlet& m)
{
	const uint8_t* src = m.src;
	const uint32_t vc = m.vertex_count, tc = m.triangle_count;
	const uint32_t code_bytes = (tc + 1) / 2, ctrl_bytes = (vc + 3) / 4;
	const uint32_t gap = code_bytes + ctrl_bytes < 16 ? 16 - (code_bytes + ctrl_bytes) : 0;
	if (m.src_size < code_bytes + ctrl_bytes + gap)
		return -2;
	const uint32_t codes = m.src_size - code_bytes;
	const uint32_t ctrl = codes - ctrl_bytes;
	const uint32_t bound = ctrl - gap; // end of the data section (at least 16 bytes follow it inside the buffer)
	uint32_t data = 0;

	// vertex references
	uint32_t last = 0xffffffffu;
	for (uint32_t i = 0; i < vc; i += 4)
	{
		if (data > bound)
			return -2;
		const uint32_t c4 = __ldg(src + ctrl + (i >> 2));
		for (uint32_t k = 0; k < 4; ++k)
		{
			const uint32_t len = c4 == 0xffu ? 4u : (((c4 >> k) & 1u) | ((c4 >> (k + 3)) & 2u));
			uint32_t v = 0;
			for (uint32_t j = 0; j < len; ++j)
				v |= (uint32_t)__ldg(src + data + j) << (8 * j);
			data += len;
			last += ((v >> 1) ^ (0u - (v & 1u))) + 1u;
			if (i + k < vc)
			{
				if (m.vertex_size == 4)
					reinterpret_cast<uint32_t*>(m.vertices)[i + k] = last;
				else
					reinterpret_cast<uint16_t*>(m.vertices)[i + k] = (uint16_t)last;
			}
		}
	}

	// triangles: packed as c | a << 8 | b << 16 | c << 24
	const uint32_t check_mask = m.triangle_size == 3 ? 3u : 1u; // byte triangles go four at a time, packed ones in pairs
	uint32_t next = 0, f0 = 0, f1 = 0, f2 = 0;
	for (uint32_t t = 0; t < tc; ++t)
	{
		if ((t & check_mask) == 0 && data > bound)
			return -2;
		const uint32_t code = (__ldg(src + codes + (t >> 1)) >> ((t & 1u) * 4u)) & 15u;
		uint32_t tri;
		if (code < 12)
		{
			uint32_t edge = code < 4 ? f0 : (code < 8 ? f1 : f2);
			edge >>= (code << 3) & 16u;
			uint32_t c;
			if (code & 1u)
				c = __ldg(src + data++);
			else
				c = next++ & 0xffu;
			tri = ((edge & 0xffu) << 16) | (edge & 0xff00u) | c | (c << 24);
		}
		else
		{
			const uint32_t a = code > 12 ? (uint32_t)__ldg(src + data++) : (next++ & 0xffu);
			const uint32_t b = code > 13 ? (uint32_t)__ldg(src + data++) : (next++ & 0xffu);
			const uint32_t c = code > 14 ? (uint32_t)__ldg(src + data++) : (next++ & 0xffu);
			tri = c | (a << 8) | (b << 16) | (c << 24);
		}
		if (m.triangle_size == 4)
			reinterpret_cast<uint32_t*>(m.triangles)[t] = tri >> 8;
		else
		{
			uint8_t* o = m.triangles + t * 3;
			o[0] = (uint8_t)(tri >> 8);
			o[1] = (uint8_t)(tri >> 16);
			o[2] = (uint8_t)(tri >> 24);
		}
		f2 = f1;
		f1 = f0;
		f0 = tri;
	}
	return data == bound ? 0 : -3;
}

__global__ void __launch_bounds__(kMeshletThreads) meshlet_decode_kernel_thread(const DevMeshlet* meshlets, int32_t* status, uint32_t n)
{
	// (first form: one thread per meshlet; kept for A/B, MOB200_MESHLET_FORM=0)
	const uint32_t i = blockIdx.x * kMeshletThreads + threadIdx.x;
	if (i < n)
		status[i] = decode_meshlet(meshlets[i]);
}

// ------------------------------------------------------------------------------------------------
// second form: EIGHT LANES per meshlet, four meshlets per warp
// ------------------------------------------------------------------------------------------------
//
// Everything that is a prefix sum is done by the eight lanes in parallel: the control bytes are known up front, so
// the byte offset of every vertex delta (lengths 0-4) and of every triangle's extra bytes (0-3 by its 4-bit code),
// the running vertex value and the running "next new vertex" counter are segmented warp scans; the lanes fetch
// their own bytes, finish the vertex references and write them coalesced, and leave one 32-bit descriptor per
// triangle in shared memory (its literal corners and its code).  What remains serial is the triangle FIFO
// (a triangle reuses an edge of one of the three before it): one short register-only chain per triangle,
//   edge = f[code / 4] >> shift;  tri = code < 12 ? edge corners + new corner : literal corners
// run by the group's lanes in step from the descriptors, so that a warp instruction of the chain serves four
// meshlets.  The finished triangles go back through shared memory and leave coalesced.  Error checks have the
// granularity of the reference's x86 path (see the header of this file); offsets only grow, so "some check in front
// of a group fails" is the same as "the first one fails".
constexpr int kGroupLanes = 8;
constexpr int kMeshletsPerCta = kMeshletThreads / kGroupLanes; // 16
constexpr uint32_t kMaxMeshletBytes = 2048; // a larger buffer cannot end where its data section does (4 * 256 + 3 * 256 + 16 + 64 + 128 = 2000): -3

__device__ __forceinline__ uint32_t load_byte(const uint8_t* src, uint32_t off, uint32_t size)
{
	return off < size ? (uint32_t)__ldg(src + off) : 0u; // (a malformed stream may point anywhere: never outside the buffer)
}

// inclusive scan over the 8 lanes of a group
__device__ __forceinline__ uint32_t group_scan(uint32_t v, uint32_t gl)
{
#pragma unroll
	for (int d = 1; d < kGroupLanes; d <<= 1)
	{
		const uint32_t o = __shfl_up_sync(0xffffffffu, v, d, kGroupLanes);
		if (gl >= (uint32_t)d)
			v += o;
	}
	return v;
}

__global__ void __launch_bounds__(kMeshletThreads) meshlet_decode_kernel(const DevMeshlet* meshlets, int32_t* status, uint32_t n)
{
	__shared__ uint32_t desc_all[kMeshletsPerCta][256]; // triangle descriptors (parallel part -> chain)
	__shared__ uint32_t tri_all[kMeshletsPerCta][256];  // finished triangles (chain -> coalesced stores)
	const uint32_t lane = threadIdx.x & 31u, gl = lane & 7u;
	const uint32_t group = threadIdx.x / kGroupLanes;
	const uint32_t idx = blockIdx.x * kMeshletsPerCta + group;
	const bool have = idx < n;
	uint32_t* desc = desc_all[group];
	uint32_t* tris = tri_all[group];

	DevMeshlet m = {};
	if (have)
		m = meshlets[idx];
	const uint8_t* src = m.src;
	const uint32_t size = m.src_size;
	const uint32_t vc = m.vertex_count, tc = m.triangle_count;
	const uint32_t code_bytes = (tc + 1) / 2, ctrl_bytes = (vc + 3) / 4;
	const uint32_t gap = code_bytes + ctrl_bytes < 16 ? 16 - (code_bytes + ctrl_bytes) : 0;
	int rc = 0;
	if (have && size < code_bytes + ctrl_bytes + gap)
		rc = -2;
	else if (have && size > kMaxMeshletBytes)
		rc = -3;
	const bool run = have && rc == 0;
	const uint32_t codes = run ? size - code_bytes : 0;
	const uint32_t ctrl = codes - (run ? ctrl_bytes : 0);
	const uint32_t bound = ctrl - (run ? gap : 0);
	const uint32_t vcr = run ? vc : 0, tcr = run ? tc : 0;

	uint32_t data = 0;          // bytes of the data section consumed so far (group-uniform)
	uint32_t last = 0xffffffffu; // running vertex reference
	bool overrun = false;

	// ---- vertex references: 32 per pass, four (one control byte) per lane ----------------------------------------------
	const uint32_t vmax = __reduce_max_sync(0xffffffffu, vcr);
	for (uint32_t base = 0; base < vmax; base += 32)
	{
		const uint32_t i0 = base + gl * 4;
		const bool on = i0 < vcr;
		const uint32_t c4 = on ? load_byte(src, ctrl + (i0 >> 2), size) : 0u;
		uint32_t len[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
			len[k] = !on ? 0u : (c4 == 0xffu ? 4u : (((c4 >> k) & 1u) | ((c4 >> (k + 3)) & 2u)));
		const uint32_t mine = len[0] + len[1] + len[2] + len[3];
		const uint32_t incl = group_scan(mine, gl);
		uint32_t off = data + incl - mine;
		overrun |= on && off > bound;
		uint32_t delta[4], dsum = 0;
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			uint32_t v = 0;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j)
				if (j < len[k])
					v |= load_byte(src, off + j, size) << (8 * j);
			off += len[k];
			delta[k] = on ? ((v >> 1) ^ (0u - (v & 1u))) + 1u : 0u;
			dsum += delta[k];
		}
		const uint32_t dincl = group_scan(dsum, gl);
		uint32_t value = last + dincl - dsum;
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			value += delta[k];
			if (i0 + k < vcr)
			{
				if (m.vertex_size == 4)
					reinterpret_cast<uint32_t*>(m.vertices)[i0 + k] = value;
				else
					reinterpret_cast<uint16_t*>(m.vertices)[i0 + k] = (uint16_t)value;
			}
		}
		data += __shfl_sync(0xffffffffu, incl, 7, kGroupLanes);
		last += __shfl_sync(0xffffffffu, dincl, 7, kGroupLanes);
	}

	// ---- triangles, parallel part: offsets of the extra bytes, new-vertex numbers, one descriptor per triangle -------------
	// descriptor = c | a << 8 | b << 16 | code << 24 (a, b only for the literal codes 12..15)
	const uint32_t check_mask = m.triangle_size == 3 ? 3u : 1u; // byte triangles are checked four at a time, packed ones in pairs
	const uint32_t tmax = __reduce_max_sync(0xffffffffu, tcr);
	uint32_t next = 0;
	for (uint32_t base = 0; base < tmax; base += 32)
	{
		const uint32_t t0 = base + gl * 4;
		uint32_t code[4], nbytes = 0, nnew = 0;
		const uint32_t cb0 = t0 < tcr ? load_byte(src, codes + (t0 >> 1), size) : 0u;
		const uint32_t cb1 = t0 + 2 < tcr ? load_byte(src, codes + (t0 >> 1) + 1, size) : 0u;
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			code[k] = ((k < 2 ? cb0 : cb1) >> ((k & 1) * 4)) & 15u;
			if (t0 + k < tcr)
			{
				const uint32_t extra = code[k] < 12 ? (code[k] & 1u) : code[k] - 12u; // bytes taken from the data section
				nbytes += extra;
				nnew += (code[k] < 12 ? 1u : 3u) - extra;                             // corners numbered by the running counter
			}
		}
		const uint32_t bincl = group_scan(nbytes, gl), nincl = group_scan(nnew, gl);
		uint32_t off = data + bincl - nbytes, nx = next + nincl - nnew;
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			const uint32_t t = t0 + k;
			if (t < tcr)
			{
				if ((t & check_mask) == 0 && off > bound)
					overrun = true;
				const uint32_t cd = code[k];
				uint32_t a = 0, b = 0, c;
				if (cd < 12)
					c = (cd & 1u) ? load_byte(src, off++, size) : (nx++ & 0xffu);
				else
				{
					a = cd > 12 ? load_byte(src, off++, size) : (nx++ & 0xffu);
					b = cd > 13 ? load_byte(src, off++, size) : (nx++ & 0xffu);
					c = cd > 14 ? load_byte(src, off++, size) : (nx++ & 0xffu);
				}
				desc[t] = c | (a << 8) | (b << 16) | (cd << 24);
			}
		}
		data += __shfl_sync(0xffffffffu, bincl, 7, kGroupLanes);
		next += __shfl_sync(0xffffffffu, nincl, 7, kGroupLanes);
	}
	__syncwarp();

	// ---- triangles, serial part: the FIFO of the last three triangles (packed c | a << 8 | b << 16 | c << 24) ------------
	// every lane of the group runs the same chain; lane t % 8 keeps triangle t
	{
		uint32_t f0 = 0, f1 = 0, f2 = 0;
		for (uint32_t t = 0; t < tmax; ++t)
		{
			const uint32_t d = t < tcr ? desc[t] : 0u;
			const uint32_t cd = d >> 24, c = d & 0xffu;
			uint32_t edge = cd < 4 ? f0 : (cd < 8 ? f1 : f2);
			edge >>= (cd << 3) & 16u;
			const uint32_t reuse = ((edge & 0xffu) << 16) | (edge & 0xff00u) | c | (c << 24);
			const uint32_t lit = (d & 0x00ffffffu) | (c << 24);
			const uint32_t tri = cd < 12 ? reuse : lit;
			f2 = f1;
			f1 = f0;
			f0 = tri;
			if (t < tcr && (t & 7u) == gl)
				tris[t] = tri;
		}
	}
	__syncwarp();

	// ---- triangles out, coalesced -------------------------------------------------------------------------------------------
	if (m.triangle_size == 4)
	{
		for (uint32_t t = gl; t < tcr; t += kGroupLanes)
			reinterpret_cast<uint32_t*>(m.triangles)[t] = tris[t] >> 8;
	}
	else
	{
		for (uint32_t o = gl; o < tcr * 3; o += kGroupLanes)
		{
			const uint32_t t = o / 3, comp = o - t * 3;
			m.triangles[o] = (uint8_t)(tris[t] >> (8 + 8 * comp));
		}
	}

	// the overrun flag of a meshlet is the OR over its eight lanes
	const uint32_t ballot = __ballot_sync(0xffffffffu, overrun);
	const bool group_overrun = ((ballot >> (lane & 24u)) & 0xffu) != 0;
	if (have && gl == 0)
		status[idx] = rc != 0 ? rc : (group_overrun ? -2 : (data == bound ? 0 : -3));
}

} // namespace mob200

using namespace mob200;

static float g_last_kernel_ms = 0.f;

extern "C" float mob200_debug_last_kernel_ms(void)
{
	return g_last_kernel_ms;
}

namespace
{

bool meshlet_args_ok(const mob200_Meshlet& m)
{
	if (m.vertex_count > 256 || m.triangle_count > 256)
		return false;
	if (m.vertex_size != 2 && m.vertex_size != 4)
		return false;
	if (m.triangle_size != 3 && m.triangle_size != 4)
		return false;
	if ((m.vertex_count && !m.vertices) || (m.triangle_count && !m.triangles))
		return false;
	return m.src_size < 0xfffffff0ull;
}

int run_meshlet_batch(mob200_Meshlet* meshlets, size_t n, cudaStream_t st)
{
	std::vector<DevMeshlet> host;
	std::vector<size_t> map;
	host.reserve(n);
	map.reserve(n);
	for (size_t i = 0; i < n; ++i)
	{
		mob200_Meshlet& m = meshlets[i];
		if (!meshlet_args_ok(m))
		{
			m.status = MOB200_ERR_ARGUMENT;
			continue;
		}
		if (!m.src)
		{
			m.status = -2;
			continue;
		}
		DevMeshlet d;
		d.src = m.src;
		d.vertices = static_cast<uint8_t*>(m.vertices);
		d.triangles = static_cast<uint8_t*>(m.triangles);
		d.src_size = (uint32_t)m.src_size;
		d.vertex_count = (uint16_t)m.vertex_count;
		d.triangle_count = (uint16_t)m.triangle_count;
		d.vertex_size = (uint8_t)m.vertex_size;
		d.triangle_size = (uint8_t)m.triangle_size;
		host.push_back(d);
		map.push_back(i);
	}
	const size_t cnt = host.size();
	if (cnt)
	{
		AsyncScratch scratch; // (freed on every return below)
		const size_t desc_bytes = cnt * sizeof(DevMeshlet);
		if (scratch.alloc(desc_bytes + cnt * sizeof(int32_t), st))
			return MOB200_ERR_CUDA;
		void* d_desc = scratch.ptr;
		int32_t* d_status = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_desc) + desc_bytes);
		CUDA_TRY(cudaMemcpyAsync(d_desc, host.data(), desc_bytes, cudaMemcpyHostToDevice, st));
		static const int form = getenv("MOB200_MESHLET_FORM") ? atoi(getenv("MOB200_MESHLET_FORM")) : 1;
		static const bool timing = getenv("MOB200_TIMING") != nullptr; // diagnostics: kernel time of the last batch (mob200_debug_last_kernel_ms)
		cudaEvent_t ev0 = nullptr, ev1 = nullptr;
		if (timing)
		{
			cudaEventCreate(&ev0);
			cudaEventCreate(&ev1);
			cudaEventRecord(ev0, st);
		}
		if (form == 0)
			meshlet_decode_kernel_thread<<<(unsigned)((cnt + kMeshletThreads - 1) / kMeshletThreads), kMeshletThreads, 0, st>>>(static_cast<const DevMeshlet*>(d_desc), d_status, (uint32_t)cnt);
		else
			meshlet_decode_kernel<<<(unsigned)((cnt + kMeshletsPerCta - 1) / kMeshletsPerCta), kMeshletThreads, 0, st>>>(static_cast<const DevMeshlet*>(d_desc), d_status, (uint32_t)cnt);
		CUDA_TRY(cudaGetLastError());
		if (timing)
			cudaEventRecord(ev1, st);
		std::vector<int32_t> rc(cnt);
		CUDA_TRY(cudaMemcpyAsync(rc.data(), d_status, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		if (timing)
		{
			cudaEventElapsedTime(&g_last_kernel_ms, ev0, ev1);
			cudaEventDestroy(ev0);
			cudaEventDestroy(ev1);
		}
		for (size_t k = 0; k < cnt; ++k)
			meshlets[map[k]].status = rc[k];
	}
	int failed = 0;
	for (size_t i = 0; i < n; ++i)
		failed += meshlets[i].status != 0;
	return failed;
}

size_t align16(size_t v)
{
	return (v + 15) & ~size_t(15);
}

} // namespace

extern "C" int mob200_decode_meshlet_batch_device(mob200_Context* ctx, mob200_Meshlet* meshlets, size_t n, void* cuda_stream)
{
	if (!ctx || (!meshlets && n))
		return MOB200_ERR_ARGUMENT;
	if (set_device(ctx))
		return MOB200_ERR_CUDA;
	return run_meshlet_batch(meshlets, n, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int mob200_decode_meshlet_batch_host(mob200_Context* ctx, mob200_Meshlet* meshlets, size_t n)
{
	if (!ctx || (!meshlets && n))
		return MOB200_ERR_ARGUMENT;
	if (set_device(ctx))
		return MOB200_ERR_CUDA;
	std::lock_guard<std::mutex> lock(ctx->mu);

	std::vector<size_t> in_off(n), v_off(n), t_off(n);
	size_t in_bytes = 0, out_bytes = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_Meshlet& m = meshlets[i];
		in_off[i] = in_bytes;
		v_off[i] = t_off[i] = out_bytes;
		if (!meshlet_args_ok(m) || !m.src)
			continue;
		in_bytes += align16(m.src_size);
		out_bytes += align16(m.vertex_count * m.vertex_size);
		t_off[i] = out_bytes;
		out_bytes += align16(m.triangle_count * m.triangle_size);
	}
	if (ctx->d_in.reserve(in_bytes + 16) || ctx->d_out.reserve(out_bytes + 16) || ctx->h_in.reserve(in_bytes + 16) || ctx->h_out.reserve(out_bytes + 16))
		return MOB200_ERR_CUDA;

	std::vector<mob200_Meshlet> dev(meshlets, meshlets + n);
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_Meshlet& m = meshlets[i];
		if (!meshlet_args_ok(m) || !m.src)
			continue;
		memcpy(static_cast<uint8_t*>(ctx->h_in.ptr) + in_off[i], m.src, m.src_size);
		dev[i].src = static_cast<const unsigned char*>(ctx->d_in.ptr) + in_off[i];
		dev[i].vertices = static_cast<uint8_t*>(ctx->d_out.ptr) + v_off[i];
		dev[i].triangles = static_cast<uint8_t*>(ctx->d_out.ptr) + t_off[i];
	}
	cudaStream_t st = ctx->stream;
	if (in_bytes)
		CUDA_TRY(cudaMemcpyAsync(ctx->d_in.ptr, ctx->h_in.ptr, in_bytes, cudaMemcpyHostToDevice, st));
	const int rc = run_meshlet_batch(dev.data(), n, st);
	if (rc < 0)
		return rc;
	if (out_bytes)
	{
		CUDA_TRY(cudaMemcpyAsync(ctx->h_out.ptr, ctx->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	for (size_t i = 0; i < n; ++i)
	{
		meshlets[i].status = dev[i].status;
		const mob200_Meshlet& m = meshlets[i];
		// (only a meshlet that decoded: the staging area is shared with earlier calls, and a rejected meshlet leaves
		// the caller's arrays untouched)
		if (!meshlet_args_ok(m) || !m.src || m.status != 0)
			continue;
		if (m.vertex_count)
			memcpy(m.vertices, static_cast<uint8_t*>(ctx->h_out.ptr) + v_off[i], m.vertex_count * m.vertex_size);
		if (m.triangle_count)
			memcpy(m.triangles, static_cast<uint8_t*>(ctx->h_out.ptr) + t_off[i], m.triangle_count * m.triangle_size);
	}
	return rc;
}

// ------------------------------------------------------------------------------------------------
// drop-in symbols (reference src/meshoptimizer.h:349-350): host pointers, synchronous
// ------------------------------------------------------------------------------------------------


extern "C" int meshopt_decodeMeshlet(void* vertices, size_t vertex_count, size_t vertex_size, void* triangles, size_t triangle_count, size_t triangle_size, const unsigned char* buffer, size_t buffer_size)
{
	PoolLease lease;
	mob200_Context* ctx = lease.ctx;
	if (!ctx)
		return MOB200_ERR_CUDA;
	mob200_Meshlet m;
	m.src = buffer;
	m.src_size = buffer_size;
	m.vertices = vertices;
	m.vertex_count = vertex_count;
	m.vertex_size = vertex_size;
	m.triangles = triangles;
	m.triangle_count = triangle_count;
	m.triangle_size = triangle_size;
	m.status = 0;
	const int rc = mob200_decode_meshlet_batch_host(ctx, &m, 1);
	return rc < 0 ? rc : m.status;
}

// (the reference's raw form may also write the padding elements up to a multiple of four vertices / two
// triangles; only the counted elements are written here)
extern "C" int meshopt_decodeMeshletRaw(unsigned int* vertices, size_t vertex_count, unsigned int* triangles, size_t triangle_count, const unsigned char* buffer, size_t buffer_size)
{
	return meshopt_decodeMeshlet(vertices, vertex_count, 4, triangles, triangle_count, 4, buffer, buffer_size);
}
