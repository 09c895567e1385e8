// mob200_index.cu -- index-stream decode on the device: the other two modes of a meshopt-compressed glTF
// bufferView (TRIANGLES -> meshopt_decodeIndexBuffer, INDICES -> meshopt_decodeIndexSequence; reference
// src/indexcodec.cpp:384-576 and :647-703, dispatched by gltf/parsegltf.cpp:584-594).
//
// Both formats are strictly sequential state machines (two 16-entry FIFOs, two running counters and a data
// cursor that advances by a data-dependent number of bytes per triangle), so the parallel unit is the stream:
// one THREAD per stream, FIFOs in shared memory in [entry][thread] order (the 32 threads of a warp hit 32
// banks whatever entries they address), code / data bytes through the read-only cache, triangles written
// straight to the destination.  Return codes are the reference's: -1 header, -2 truncated, -3 size mismatch.
// This is the batched form a loader needs next to the attribute streams (many meshes, one launch); a single
// long index stream is decoded by one thread and is latency-bound by construction.
#include "mob200_host.h"

#include <stdint.h>
#include <stdlib.h>

#include <vector>

namespace mob200
{

constexpr int kIndexThreads = 64; // streams per CTA: 64 x (16 + 32) x 4 bytes of FIFO state = 12 KB shared memory

struct DevIndexStream
{
	const uint8_t* src;
	uint8_t* dst;
	uint32_t src_size;
	uint32_t index_count;
	uint32_t index_size; // 2 or 4
	uint32_t kind;       // enum mob200_IndexKind
};

// LEB128-style group of up to five bytes (reference decodeVByte, :95-118): always terminates after 5 bytes
__device__ __forceinline__ uint32_t index_varint(const uint8_t* src, uint32_t& at)
{
	uint32_t lead = __ldg(src + at++);
	if (lead < 128)
		return lead;
	uint32_t result = lead & 127u, shift = 7;
	for (int i = 0; i < 4; ++i)
	{
		const uint32_t group = __ldg(src + at++);
		result |= (group & 127u) << shift;
		shift += 7;
		if (group < 128)
			break;
	}
	return result;
}

// zigzag delta against the previous free index (reference decodeIndex, :130-136)
__device__ __forceinline__ uint32_t index_delta(const uint8_t* src, uint32_t& at, uint32_t last)
{
	const uint32_t v = index_varint(src, at);
	return last + ((v >> 1) ^ (0u - (v & 1u)));
}

__device__ __forceinline__ void index_store3(uint8_t* dst, uint32_t index_size, size_t tri, uint32_t a, uint32_t b, uint32_t c)
{
	if (index_size == 2)
	{
		uint16_t* p = reinterpret_cast<uint16_t*>(dst) + tri * 3;
		p[0] = (uint16_t)a;
		p[1] = (uint16_t)b;
		p[2] = (uint16_t)c;
	}
	else
	{
		uint32_t* p = reinterpret_cast<uint32_t*>(dst) + tri * 3;
		p[0] = a;
		p[1] = b;
		p[2] = c;
	}
}

// FIFO state of one thread: entry e of the vertex FIFO at vf[e * stride], edge FIFO halves at ea / eb
struct IndexFifos
{
	uint32_t* vf;
	uint32_t* ea;
	uint32_t* eb;
	uint32_t stride;
	uint32_t vo, eo; // write positions

	__device__ __forceinline__ uint32_t vertex(uint32_t back) const { return vf[((vo - back) & 15u) * stride]; }
	__device__ __forceinline__ void push_vertex(uint32_t v, uint32_t advance)
	{
		vf[vo * stride] = v; // (the slot is written even when the position does not advance, as the reference does)
		vo = (vo + advance) & 15u;
	}
	__device__ __forceinline__ void push_edge(uint32_t a, uint32_t b)
	{
		ea[eo * stride] = a;
		eb[eo * stride] = b;
		eo = (eo + 1u) & 15u;
	}
};

// triangle list (reference meshopt_decodeIndexBuffer, :384-576)
__device__ int decode_triangles(const DevIndexStream& s, IndexFifos F)
{
	const uint8_t* src = s.src;
	const uint32_t size = s.src_size;
	const uint32_t tris = s.index_count / 3;

	if ((uint64_t)size < 1ull + tris + 16ull)
		return -2;
	const uint32_t head = __ldg(src);
	if ((head & 0xf0u) != 0xe0u)
		return -1;
	const uint32_t version = head & 0x0fu;
	if (version > 1)
		return -1;
	const uint32_t fecmax = version >= 1 ? 13u : 15u;

	for (uint32_t e = 0; e < 16; ++e)
	{
		F.vf[e * F.stride] = 0xffffffffu;
		F.ea[e * F.stride] = 0xffffffffu;
		F.eb[e * F.stride] = 0xffffffffu;
	}
	F.vo = F.eo = 0;

	uint32_t next = 0, last = 0;
	uint32_t data = 1 + tris;            // triangle data follows the code bytes
	const uint32_t safe_end = size - 16; // ... and ends where the 16-byte codeaux table starts
	const uint32_t table = safe_end;

	for (uint32_t t = 0; t < tris; ++t)
	{
		const uint32_t code = __ldg(src + 1 + t);
		uint32_t a, b, c;
		if (code < 0xf0u)
		{
			// edge reuse: (a, b) from the edge FIFO, c from the vertex FIFO, a counter, or the data stream
			const uint32_t slot = ((F.eo - 1u - (code >> 4)) & 15u) * F.stride;
			a = F.ea[slot];
			b = F.eb[slot];
			const uint32_t fec = code & 15u;
			if (fec < fecmax)
			{
				const uint32_t cf = F.vertex(1u + fec);
				const uint32_t fresh = fec == 0;
				c = fresh ? next : cf;
				next += fresh;
				F.push_vertex(c, fresh);
			}
			else
			{
				if (data > safe_end)
					return -2;
				c = fec != 15u ? last + (fec * 2u - 27u) : index_delta(src, data, last);
				last = c;
				F.push_vertex(c, 1);
			}
			F.push_edge(c, b);
			F.push_edge(a, c);
		}
		else
		{
			uint32_t feb, fec, adv_b, adv_c;
			if (code < 0xfeu)
			{
				const uint32_t aux = __ldg(src + table + (code & 15u));
				feb = aux >> 4;
				fec = aux & 15u;
				a = next++;
				const uint32_t bf = F.vertex(feb), cf = F.vertex(fec); // (read before any push of this triangle)
				adv_b = feb == 0;
				b = adv_b ? next : bf;
				next += adv_b;
				adv_c = fec == 0;
				c = adv_c ? next : cf;
				next += adv_c;
			}
			else
			{
				if (data > safe_end)
					return -2;
				const uint32_t aux = __ldg(src + data++);
				const uint32_t fea = code == 0xfeu ? 0u : 15u;
				feb = aux >> 4;
				fec = aux & 15u;
				if (aux == 0)
					next = 0; // reset marker
				a = fea == 0 ? next++ : 0u;
				b = feb == 0 ? next++ : F.vertex(feb);
				c = fec == 0 ? next++ : F.vertex(fec);
				if (fea == 15u)
					last = a = index_delta(src, data, last);
				if (feb == 15u)
					last = b = index_delta(src, data, last);
				if (fec == 15u)
					last = c = index_delta(src, data, last);
				adv_b = (feb == 0) | (feb == 15u);
				adv_c = (fec == 0) | (fec == 15u);
			}
			F.push_vertex(a, 1);
			F.push_vertex(b, adv_b);
			F.push_vertex(c, adv_c);
			F.push_edge(b, a);
			F.push_edge(c, b);
			F.push_edge(a, c);
		}
		index_store3(s.dst, s.index_size, t, a, b, c);
	}
	return data == safe_end ? 0 : -3;
}

// index sequence (reference meshopt_decodeIndexSequence, :647-703)
__device__ int decode_sequence(const DevIndexStream& s)
{
	const uint8_t* src = s.src;
	const uint32_t size = s.src_size;
	if ((uint64_t)size < 1ull + s.index_count + 4ull)
		return -2;
	const uint32_t head = __ldg(src);
	if ((head & 0xf0u) != 0xd0u)
		return -1;
	if ((head & 0x0fu) > 1)
		return -1;

	uint32_t data = 1;
	const uint32_t safe_end = size - 4;
	uint32_t last0 = 0, last1 = 0;
	for (uint32_t i = 0; i < s.index_count; ++i)
	{
		if (data >= safe_end)
			return -2;
		uint32_t v = index_varint(src, data);
		const uint32_t which = v & 1u; // baseline this index is a delta of
		v >>= 1;
		const uint32_t d = (v >> 1) ^ (0u - (v & 1u));
		const uint32_t index = (which ? last1 : last0) + d;
		if (which)
			last1 = index;
		else
			last0 = index;
		if (s.index_size == 2)
			reinterpret_cast<uint16_t*>(s.dst)[i] = (uint16_t)index;
		else
			reinterpret_cast<uint32_t*>(s.dst)[i] = index;
	}
	return data == safe_end ? 0 : -3;
}

// kSparse = false: one stream per THREAD (many streams: a warp's 32 lanes each follow their own stream, diverging freely;
//                  throughput comes from the number of streams).
// kSparse = true:  one stream per WARP, lane 0 alone runs the state machine (few long streams: the 32 lanes of a warp
//                  that each follow another branch of another stream execute one after the other, so a long list decoded
//                  next to 31 others runs at 1/32 of a warp's speed; alone on its warp it never diverges).
template <bool kSparse>
__global__ void __launch_bounds__(kIndexThreads) index_decode_kernel(const DevIndexStream* streams, int32_t* status, uint32_t n)
{
	__shared__ uint32_t fifo[48 * kIndexThreads];
	const uint32_t slot = kSparse ? threadIdx.x >> 5 : threadIdx.x;
	const uint32_t i = kSparse ? blockIdx.x * (kIndexThreads / 32) + slot : blockIdx.x * kIndexThreads + threadIdx.x;
	if (i >= n || (kSparse && (threadIdx.x & 31u)))
		return;
	const DevIndexStream s = streams[i];
	int rc;
	if (s.kind == MOB200_INDEX_TRIANGLES)
	{
		IndexFifos F;
		F.stride = kIndexThreads;
		F.vf = fifo + slot;
		F.ea = fifo + 16 * kIndexThreads + slot;
		F.eb = fifo + 32 * kIndexThreads + slot;
		F.vo = F.eo = 0;
		rc = decode_triangles(s, F);
	}
	else
		rc = decode_sequence(s);
	status[i] = rc;
}

} // namespace mob200

using namespace mob200;

namespace
{

bool index_args_ok(const mob200_IndexStream& s)
{
	if (s.index_size != 2 && s.index_size != 4)
		return false;
	if (s.kind != MOB200_INDEX_TRIANGLES && s.kind != MOB200_INDEX_SEQUENCE)
		return false;
	if (s.kind == MOB200_INDEX_TRIANGLES && s.index_count % 3 != 0)
		return false;
	if (s.index_count >= 0xfffffff0ull || s.src_size >= 0xfffffff0ull)
		return false;
	if (s.index_count && !s.dst)
		return false;
	return true;
}

// streams with device pointers -> one launch; status[] (host) receives the return codes
int run_index_batch(mob200_Context* ctx, mob200_IndexStream* streams, size_t n, cudaStream_t st)
{
	std::vector<DevIndexStream> host(n);
	std::vector<size_t> map;
	map.reserve(n);
	size_t m = 0;
	for (size_t i = 0; i < n; ++i)
	{
		mob200_IndexStream& s = streams[i];
		if (!index_args_ok(s))
		{
			s.status = MOB200_ERR_ARGUMENT;
			continue;
		}
		if (s.src_size == 0 || !s.src)
		{
			s.status = -2; // the reference's size check fails first (:390, :652)
			continue;
		}
		DevIndexStream& d = host[m++];
		d.src = s.src;
		d.dst = static_cast<uint8_t*>(s.dst);
		d.src_size = (uint32_t)s.src_size;
		d.index_count = (uint32_t)s.index_count;
		d.index_size = (uint32_t)s.index_size;
		d.kind = (uint32_t)s.kind;
		map.push_back(i);
	}
	if (m)
	{
		AsyncScratch scratch; // (freed on every return below)
		const size_t desc_bytes = m * sizeof(DevIndexStream);
		if (scratch.alloc(desc_bytes + m * sizeof(int32_t), st))
			return MOB200_ERR_CUDA;
		void* d_desc = scratch.ptr;
		int32_t* d_status = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_desc) + desc_bytes);
		CUDA_TRY(cudaMemcpyAsync(d_desc, host.data(), desc_bytes, cudaMemcpyHostToDevice, st));
		// few long streams: one warp each (at most one wave of warps on the device); else one thread each
		size_t total_indices = 0;
		for (size_t k = 0; k < m; ++k)
			total_indices += host[k].index_count;
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx ? ctx->device : 0);
		static const int form = getenv("MOB200_INDEX_FORM") ? atoi(getenv("MOB200_INDEX_FORM")) : -1;
		const bool sparse = form >= 0 ? form == 1 : (m <= (size_t)sms * 32 && total_indices / m >= 3 * 1024);
		if (sparse)
			index_decode_kernel<true><<<(unsigned)((m + kIndexThreads / 32 - 1) / (kIndexThreads / 32)), kIndexThreads, 0, st>>>(static_cast<const DevIndexStream*>(d_desc), d_status, (uint32_t)m);
		else
			index_decode_kernel<false><<<(unsigned)((m + kIndexThreads - 1) / kIndexThreads), kIndexThreads, 0, st>>>(static_cast<const DevIndexStream*>(d_desc), d_status, (uint32_t)m);
		CUDA_TRY(cudaGetLastError());
		std::vector<int32_t> rc(m);
		CUDA_TRY(cudaMemcpyAsync(rc.data(), d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		for (size_t k = 0; k < m; ++k)
			streams[map[k]].status = rc[k];
	}
	int failed = 0;
	for (size_t i = 0; i < n; ++i)
		failed += streams[i].status != 0;
	return failed;
}

} // namespace

extern "C" int mob200_decode_index_batch_device(mob200_Context* ctx, mob200_IndexStream* streams, size_t n, void* cuda_stream)
{
	if (!ctx || (!streams && n))
		return MOB200_ERR_ARGUMENT;
	if (set_device(ctx))
		return MOB200_ERR_CUDA;
	return run_index_batch(ctx, streams, n, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int mob200_decode_index_batch_host(mob200_Context* ctx, mob200_IndexStream* streams, size_t n)
{
	if (!ctx || (!streams && n))
		return MOB200_ERR_ARGUMENT;
	if (set_device(ctx))
		return MOB200_ERR_CUDA;
	std::lock_guard<std::mutex> lock(ctx->mu);

	// one staging round trip for the whole batch: encoded bytes in (16-byte aligned pieces), indices out
	std::vector<size_t> in_off(n), out_off(n);
	size_t in_bytes = 0, out_bytes = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_IndexStream& s = streams[i];
		in_off[i] = in_bytes;
		out_off[i] = out_bytes;
		if (!index_args_ok(s) || !s.src)
			continue;
		in_bytes += (s.src_size + 15) & ~size_t(15);
		out_bytes += (s.index_count * s.index_size + 15) & ~size_t(15);
	}
	if (ctx->d_in.reserve(in_bytes + 16) || ctx->d_out.reserve(out_bytes + 16) || ctx->h_in.reserve(in_bytes + 16) || ctx->h_out.reserve(out_bytes + 16))
		return MOB200_ERR_CUDA;

	std::vector<mob200_IndexStream> dev(streams, streams + n);
	for (size_t i = 0; i < n; ++i)
	{
		const mob200_IndexStream& s = streams[i];
		if (!index_args_ok(s) || !s.src)
			continue;
		memcpy(static_cast<uint8_t*>(ctx->h_in.ptr) + in_off[i], s.src, s.src_size);
		dev[i].src = static_cast<const unsigned char*>(ctx->d_in.ptr) + in_off[i];
		dev[i].dst = static_cast<uint8_t*>(ctx->d_out.ptr) + out_off[i];
	}
	cudaStream_t st = ctx->stream;
	if (in_bytes)
		CUDA_TRY(cudaMemcpyAsync(ctx->d_in.ptr, ctx->h_in.ptr, in_bytes, cudaMemcpyHostToDevice, st));
	const int rc = run_index_batch(ctx, dev.data(), n, st);
	if (rc < 0)
		return rc;
	if (out_bytes)
	{
		CUDA_TRY(cudaMemcpyAsync(ctx->h_out.ptr, ctx->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	for (size_t i = 0; i < n; ++i)
	{
		streams[i].status = dev[i].status;
		const mob200_IndexStream& s = streams[i];
		// (only a stream that decoded: the staging area is shared with earlier calls, and a rejected stream leaves
		// the caller's array untouched, which the reference's "may produce garbage" allows)
		if (index_args_ok(s) && s.src && s.dst && s.status == 0)
			memcpy(s.dst, static_cast<uint8_t*>(ctx->h_out.ptr) + out_off[i], s.index_count * s.index_size);
	}
	return rc;
}

// ------------------------------------------------------------------------------------------------
// drop-in symbols (reference src/meshoptimizer.h:344,351,376): host pointers, synchronous
// ------------------------------------------------------------------------------------------------


static int decode_index_dropin(int kind, void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size)
{
	PoolLease lease;
	mob200_Context* ctx = lease.ctx;
	if (!ctx)
		return MOB200_ERR_CUDA;
	mob200_IndexStream s;
	s.src = buffer;
	s.src_size = buffer_size;
	s.dst = destination;
	s.index_count = index_count;
	s.index_size = index_size;
	s.kind = kind;
	s.status = 0;
	const int rc = mob200_decode_index_batch_host(ctx, &s, 1);
	return rc < 0 ? rc : s.status;
}

extern "C" int meshopt_decodeIndexBuffer(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size)
{
	return decode_index_dropin(MOB200_INDEX_TRIANGLES, destination, index_count, index_size, buffer, buffer_size);
}

extern "C" int meshopt_decodeIndexSequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size)
{
	return decode_index_dropin(MOB200_INDEX_SEQUENCE, destination, index_count, index_size, buffer, buffer_size);
}

// (pure header inspection, reference src/indexcodec.cpp:364-382)
extern "C" int meshopt_decodeIndexVersion(const unsigned char* buffer, size_t buffer_size)
{
	if (buffer_size < 1)
		return -1;
	const unsigned char header = buffer[0];
	if ((header & 0xf0) != 0xe0 && (header & 0xf0) != 0xd0)
		return -1;
	const int version = header & 0x0f;
	return version > 1 ? -1 : version;
}
