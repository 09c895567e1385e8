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

__global__ void __launch_bounds__(kMeshletThreads) meshlet_decode_kernel(const DevMeshlet* meshlets, int32_t* status, uint32_t n)
{
	const uint32_t i = blockIdx.x * kMeshletThreads + threadIdx.x;
	if (i < n)
		status[i] = decode_meshlet(meshlets[i]);
}

} // namespace mob200

using namespace mob200;

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
		void* d_desc = nullptr;
		const size_t desc_bytes = cnt * sizeof(DevMeshlet);
		CUDA_TRY(cudaMallocAsync(&d_desc, desc_bytes + cnt * sizeof(int32_t), st));
		int32_t* d_status = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_desc) + desc_bytes);
		CUDA_TRY(cudaMemcpyAsync(d_desc, host.data(), desc_bytes, cudaMemcpyHostToDevice, st));
		meshlet_decode_kernel<<<(unsigned)((cnt + kMeshletThreads - 1) / kMeshletThreads), kMeshletThreads, 0, st>>>(static_cast<const DevMeshlet*>(d_desc), d_status, (uint32_t)cnt);
		CUDA_TRY(cudaGetLastError());
		std::vector<int32_t> rc(cnt);
		CUDA_TRY(cudaMemcpyAsync(rc.data(), d_status, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		CUDA_TRY(cudaFreeAsync(d_desc, st));
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
		if (!meshlet_args_ok(m) || !m.src)
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
