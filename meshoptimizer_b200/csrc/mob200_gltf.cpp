// mob200_gltf.cpp -- glTF bufferView front-end of the decode path (host side, plain C++).
//
// The reference consumer of the codecs is a loop over the bufferViews of a glTF asset that carry the
// EXT_meshopt_compression / KHR_meshopt_compression extension (reference gltf/parsegltf.cpp:561-627,
// decompressMeshopt): one decode call per view by `mode`, one filter call by `filter`, the result lands
// where the view's own buffer/byteOffset/byteLength say (the "fallback" buffer gltfpack declares,
// gltf/write.cpp:739-790).  This file turns that loop into ONE batched device decode:
//
//   mob200_gltf_scan           finds the compressed views in a .glb container or bare .gltf JSON text
//                              (a small JSON scanner of our own: the reference uses the third-party cgltf
//                              parser, extern/cgltf.h:5045-5130,5225-5245; its validation rules :1645-1667
//                              are applied here as well)
//   mob200_gltf_decode_host    all views of an asset, host pointers in / out, synchronous
//   mob200_gltf_decode_device  same with device pointers
//
// ATTRIBUTES views (with their filters) go through the vertex-stream kernels as one batch; TRIANGLES and
// INDICES views go through the index-stream kernels (mob200_index.cu).  Nothing here decodes on the CPU.
#include "../../include/meshopt_b200.h"

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace
{

// ------------------------------------------------------------------------------------------------
// minimal JSON scanner (RFC 8259 subset that glTF uses): values are skipped structurally, numbers are
// read as unsigned integers where the schema asks for them
// ------------------------------------------------------------------------------------------------
struct Json
{
	const char* p;
	const char* end;
	bool ok;
	int depth; // nesting of the value being read: an untrusted asset must not be able to exhaust the host stack
	static const int kMaxDepth = 64;

	Json(const char* b, const char* e) : p(b), end(e), ok(true), depth(0) {}

	bool enter()
	{
		if (++depth > kMaxDepth)
		{
			fail();
			return false;
		}
		return true;
	}
	void leave() { --depth; }

	void ws()
	{
		while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r'))
			++p;
	}

	bool eat(char c)
	{
		ws();
		if (p < end && *p == c)
		{
			++p;
			return true;
		}
		return false;
	}

	bool peek(char c)
	{
		ws();
		return p < end && *p == c;
	}

	void fail()
	{
		ok = false;
		p = end;
	}

	// string without escape processing beyond skipping (keys and enum values of glTF are plain ASCII)
	bool string(std::string& out)
	{
		out.clear();
		if (!eat('"'))
		{
			fail();
			return false;
		}
		while (p < end && *p != '"')
		{
			if (*p == '\\')
			{
				if (p + 1 >= end)
					break;
				out.push_back(p[1]);
				p += (p[1] == 'u') ? 6 : 2;
				if (p > end)
					p = end;
				continue;
			}
			out.push_back(*p++);
		}
		if (p >= end)
		{
			fail();
			return false;
		}
		++p;
		return true;
	}

	bool number(double& out)
	{
		ws();
		const char* s = p;
		if (p < end && (*p == '-' || *p == '+'))
			++p;
		double v = 0;
		bool digits = false;
		while (p < end && *p >= '0' && *p <= '9')
		{
			v = v * 10 + (*p - '0');
			++p;
			digits = true;
		}
		if (p < end && *p == '.')
		{
			++p;
			double f = 0.1;
			while (p < end && *p >= '0' && *p <= '9')
			{
				v += f * (*p - '0');
				f *= 0.1;
				++p;
				digits = true;
			}
		}
		if (p < end && (*p == 'e' || *p == 'E'))
		{
			++p;
			bool neg = false;
			if (p < end && (*p == '-' || *p == '+'))
				neg = *p++ == '-';
			int ex = 0;
			while (p < end && *p >= '0' && *p <= '9')
				ex = ex * 10 + (*p++ - '0');
			for (int i = 0; i < ex && i < 400; ++i)
				v = neg ? v / 10 : v * 10;
		}
		if (!digits)
		{
			fail();
			return false;
		}
		out = (*s == '-') ? -v : v;
		return true;
	}

	bool size(size_t& out)
	{
		double v;
		if (!number(v) || v < 0 || v > 9.0e15)
		{
			fail();
			return false;
		}
		out = (size_t)(v + 0.5);
		return true;
	}

	void skip()
	{
		ws();
		if (p >= end)
		{
			fail();
			return;
		}
		if (*p == '"')
		{
			std::string s;
			string(s);
		}
		else if (*p == '{')
		{
			++p;
			if (eat('}'))
				return;
			if (!enter())
				return;
			do
			{
				std::string k;
				if (!string(k) || !eat(':'))
				{
					fail();
					return;
				}
				skip();
			} while (ok && eat(','));
			leave();
			if (!eat('}'))
				fail();
		}
		else if (*p == '[')
		{
			++p;
			if (eat(']'))
				return;
			if (!enter())
				return;
			do
				skip();
			while (ok && eat(','));
			leave();
			if (!eat(']'))
				fail();
		}
		else if (*p == 't' || *p == 'f' || *p == 'n')
		{
			while (p < end && *p >= 'a' && *p <= 'z')
				++p;
		}
		else
		{
			double v;
			number(v);
		}
	}

};

// walks `{ "k": v, ... }` calling f(key) for every member; f must consume the value
template <typename F>
void each_member(Json& j, F f)
{
	if (!j.eat('{'))
	{
		j.fail();
		return;
	}
	if (j.eat('}'))
		return;
	if (!j.enter())
		return;
	do
	{
		std::string key;
		if (!j.string(key) || !j.eat(':'))
		{
			j.fail();
			return;
		}
		f(key);
	} while (j.ok && j.eat(','));
	j.leave();
	if (!j.eat('}'))
		j.fail();
}

template <typename F>
void each_element(Json& j, F f)
{
	if (!j.eat('['))
	{
		j.fail();
		return;
	}
	if (j.eat(']'))
		return;
	if (!j.enter())
		return;
	size_t index = 0;
	do
		f(index++);
	while (j.ok && j.eat(','));
	j.leave();
	if (!j.eat(']'))
		j.fail();
}

struct ViewFields
{
	bool compressed = false;
	bool have_mode = false;
	size_t buffer = 0, offset = 0, length = 0;
	size_t src_buffer = 0, src_offset = 0, src_size = 0, stride = 0, count = 0;
	int mode = -1;
	int filter = MOB200_FILTER_NONE;
	bool bad = false;
};

void parse_extension(Json& j, ViewFields& v)
{
	v.compressed = true;
	each_member(j, [&](const std::string& key) {
		if (key == "buffer")
			j.size(v.src_buffer);
		else if (key == "byteOffset")
			j.size(v.src_offset);
		else if (key == "byteLength")
			j.size(v.src_size);
		else if (key == "byteStride")
			j.size(v.stride);
		else if (key == "count")
			j.size(v.count);
		else if (key == "mode")
		{
			std::string s;
			j.string(s);
			v.have_mode = true;
			v.mode = s == "ATTRIBUTES" ? MOB200_GLTF_ATTRIBUTES : (s == "TRIANGLES" ? MOB200_GLTF_TRIANGLES : (s == "INDICES" ? MOB200_GLTF_INDICES : -1));
		}
		else if (key == "filter")
		{
			std::string s;
			j.string(s);
			if (s == "NONE")
				v.filter = MOB200_FILTER_NONE;
			else if (s == "OCTAHEDRAL")
				v.filter = MOB200_FILTER_OCTAHEDRAL;
			else if (s == "QUATERNION")
				v.filter = MOB200_FILTER_QUATERNION;
			else if (s == "EXPONENTIAL")
				v.filter = MOB200_FILTER_EXPONENTIAL;
			else if (s == "COLOR")
				v.filter = MOB200_FILTER_COLOR;
			else
				v.bad = true;
		}
		else
			j.skip();
	});
}

// the structural rules cgltf_validate applies to a compressed view (extern/cgltf.h:1645-1667)
bool view_valid(const ViewFields& v)
{
	if (v.bad || !v.have_mode || v.mode < 0)
		return false;
	// (sizes come from untrusted JSON: no product may wrap, and the codecs address 32-bit counts and sizes)
	if (v.count >= 0xffffffffull || v.src_size >= 0xffffffffull || v.stride > 256)
		return false;
	if (v.length != v.count * v.stride)
		return false;
	if (v.mode == MOB200_GLTF_ATTRIBUTES && !(v.stride % 4 == 0 && v.stride <= 256 && v.stride > 0))
		return false;
	if (v.mode == MOB200_GLTF_TRIANGLES && v.count % 3 != 0)
		return false;
	if ((v.mode == MOB200_GLTF_TRIANGLES || v.mode == MOB200_GLTF_INDICES) && v.stride != 2 && v.stride != 4)
		return false;
	if ((v.mode == MOB200_GLTF_TRIANGLES || v.mode == MOB200_GLTF_INDICES) && v.filter != MOB200_FILTER_NONE)
		return false;
	if (v.filter == MOB200_FILTER_OCTAHEDRAL && v.stride != 4 && v.stride != 8)
		return false;
	if (v.filter == MOB200_FILTER_QUATERNION && v.stride != 8)
		return false;
	if (v.filter == MOB200_FILTER_COLOR && v.stride != 4 && v.stride != 8)
		return false;
	return true;
}

uint32_t rd32(const unsigned char* p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

} // namespace

extern "C" int mob200_gltf_scan(const void* data, size_t size, mob200_GltfView* views, size_t view_capacity, size_t* buffer_sizes, size_t buffer_capacity, mob200_GltfInfo* info)
{
	if (!data || !info)
		return MOB200_ERR_ARGUMENT;
	memset(info, 0, sizeof(*info));
	const unsigned char* bytes = static_cast<const unsigned char*>(data);

	// .glb container: 12-byte header, then chunks (length, type, payload): JSON first, BIN second
	size_t json_off = 0, json_size = size;
	if (size >= 12 && rd32(bytes) == 0x46546c67u) // "glTF"
	{
		if (rd32(bytes + 4) != 2 || rd32(bytes + 8) > size)
			return MOB200_ERR_ARGUMENT;
		size_t total = rd32(bytes + 8), at = 12;
		bool have_json = false;
		while (at + 8 <= total)
		{
			const size_t len = rd32(bytes + at);
			const uint32_t type = rd32(bytes + at + 4);
			if (len > total - at - 8)
				return MOB200_ERR_ARGUMENT;
			if (type == 0x4e4f534au && !have_json) // "JSON"
			{
				json_off = at + 8;
				json_size = len;
				have_json = true;
			}
			else if (type == 0x004e4942u && info->bin_size == 0) // "BIN\0"
			{
				info->bin_offset = at + 8;
				info->bin_size = len;
			}
			at += 8 + ((len + 3) & ~size_t(3));
		}
		if (!have_json)
			return MOB200_ERR_ARGUMENT;
	}
	info->json_offset = json_off;
	info->json_size = json_size;

	Json j(reinterpret_cast<const char*>(bytes) + json_off, reinterpret_cast<const char*>(bytes) + json_off + json_size);
	size_t n_views = 0, n_buffers = 0, n_invalid = 0;

	each_member(j, [&](const std::string& key) {
		if (key == "buffers")
		{
			each_element(j, [&](size_t index) {
				size_t length = 0;
				each_member(j, [&](const std::string& k) {
					if (k == "byteLength")
						j.size(length);
					else
						j.skip();
				});
				if (buffer_sizes && index < buffer_capacity)
					buffer_sizes[index] = length;
				n_buffers = index + 1;
			});
		}
		else if (key == "bufferViews")
		{
			each_element(j, [&](size_t index) {
				ViewFields v;
				each_member(j, [&](const std::string& k) {
					if (k == "buffer")
						j.size(v.buffer);
					else if (k == "byteOffset")
						j.size(v.offset);
					else if (k == "byteLength")
						j.size(v.length);
					else if (k == "extensions")
					{
						each_member(j, [&](const std::string& ext) {
							if (ext == "EXT_meshopt_compression" || ext == "KHR_meshopt_compression")
								parse_extension(j, v);
							else
								j.skip();
						});
					}
					else
						j.skip();
				});
				if (!v.compressed)
					return;
				const bool valid = view_valid(v);
				n_invalid += !valid;
				if (views && n_views < view_capacity)
				{
					mob200_GltfView& o = views[n_views];
					o.view = index;
					o.mode = v.mode;
					o.filter = v.filter;
					o.src_buffer = v.src_buffer;
					o.src_offset = v.src_offset;
					o.src_size = v.src_size;
					o.count = v.count;
					o.stride = v.stride;
					o.dst_buffer = v.buffer;
					o.dst_offset = v.offset;
					o.dst_size = v.length;
					o.status = valid ? 0 : MOB200_ERR_ARGUMENT;
				}
				++n_views;
			});
		}
		else
			j.skip();
	});
	if (!j.ok)
		return MOB200_ERR_ARGUMENT;

	// a view that names a buffer the asset does not have, or a range outside its buffers, is invalid (cgltf_validate
	// rejects the asset, extern/cgltf.h:1645-1667); buffers[] may follow bufferViews[] in the text, hence this pass
	if (views)
		for (size_t k = 0; k < n_views && k < view_capacity; ++k)
		{
			mob200_GltfView& o = views[k];
			if (o.status != 0)
				continue;
			bool bad = o.src_buffer >= n_buffers || o.dst_buffer >= n_buffers;
			if (!bad && buffer_sizes && o.src_buffer < buffer_capacity && o.dst_buffer < buffer_capacity)
			{
				const size_t sb = buffer_sizes[o.src_buffer], db = buffer_sizes[o.dst_buffer];
				bad = o.src_offset > sb || o.src_size > sb - o.src_offset || o.dst_offset > db || o.dst_size > db - o.dst_offset;
			}
			if (bad)
			{
				o.status = MOB200_ERR_ARGUMENT;
				++n_invalid;
			}
		}

	info->buffer_count = n_buffers;
	info->view_count = n_views;
	info->invalid_views = n_invalid;
	return 0;
}

namespace
{

// shared by the host and device entry points: views -> stream descriptors of the two codec families
int decode_views(mob200_Context* ctx, mob200_GltfView* views, size_t n, size_t buffer_count, const void* const* buffers, const size_t* buffer_sizes, void* const* outputs, const size_t* output_sizes,
    bool device, void* cuda_stream)
{
	if (!ctx || (!views && n) || (n && (!buffers || !buffer_sizes || !outputs || !output_sizes)))
		return MOB200_ERR_ARGUMENT;

	std::vector<mob200_Stream> vstreams;
	std::vector<mob200_IndexStream> istreams;
	std::vector<size_t> vmap, imap;
	for (size_t i = 0; i < n; ++i)
	{
		mob200_GltfView& v = views[i];
		if (v.status == MOB200_ERR_ARGUMENT) // rejected by the scan
			continue;
		v.status = MOB200_ERR_ARGUMENT;
		// every index and range of a view comes from the asset's JSON: nothing is dereferenced before it is bounded
		if (v.src_buffer >= buffer_count || v.dst_buffer >= buffer_count || !buffers[v.src_buffer] || !outputs[v.dst_buffer])
			continue;
		if (v.src_offset > buffer_sizes[v.src_buffer] || v.src_size > buffer_sizes[v.src_buffer] - v.src_offset)
			continue;
		if (v.count >= 0xffffffffull || v.stride == 0 || v.stride > 256 || v.dst_size != v.count * v.stride)
			continue;
		if (v.dst_offset > output_sizes[v.dst_buffer] || v.dst_size > output_sizes[v.dst_buffer] - v.dst_offset)
			continue;
		const unsigned char* src = static_cast<const unsigned char*>(buffers[v.src_buffer]) + v.src_offset;
		unsigned char* dst = static_cast<unsigned char*>(outputs[v.dst_buffer]) + v.dst_offset;
		if (v.mode == MOB200_GLTF_ATTRIBUTES)
		{
			mob200_Stream s;
			s.src = src;
			s.src_size = v.src_size;
			s.dst = dst;
			s.vertex_count = v.count;
			s.vertex_size = v.stride;
			s.filter = v.filter;
			s.status = 0;
			vstreams.push_back(s);
			vmap.push_back(i);
		}
		else
		{
			mob200_IndexStream s;
			s.src = src;
			s.src_size = v.src_size;
			s.dst = dst;
			s.index_count = v.count;
			s.index_size = v.stride;
			s.kind = v.mode == MOB200_GLTF_TRIANGLES ? MOB200_INDEX_TRIANGLES : MOB200_INDEX_SEQUENCE;
			s.status = 0;
			istreams.push_back(s);
			imap.push_back(i);
		}
	}

	int rc = 0;
	if (!vstreams.empty())
		rc = device ? mob200_decode_batch_device(ctx, vstreams.data(), vstreams.size(), cuda_stream) : mob200_decode_batch_host(ctx, vstreams.data(), vstreams.size());
	if (rc < 0 && rc <= MOB200_ERR_CUDA)
		return rc;
	int rci = 0;
	if (!istreams.empty())
		rci = device ? mob200_decode_index_batch_device(ctx, istreams.data(), istreams.size(), cuda_stream) : mob200_decode_index_batch_host(ctx, istreams.data(), istreams.size());
	if (rci < 0 && rci <= MOB200_ERR_CUDA)
		return rci;

	for (size_t k = 0; k < vmap.size(); ++k)
		views[vmap[k]].status = vstreams[k].status;
	for (size_t k = 0; k < imap.size(); ++k)
		views[imap[k]].status = istreams[k].status;
	int failed = 0;
	for (size_t i = 0; i < n; ++i)
		failed += views[i].status != 0;
	return failed;
}

} // namespace

extern "C" int mob200_gltf_decode_host(mob200_Context* ctx, mob200_GltfView* views, size_t n, size_t buffer_count, const void* const* buffers, const size_t* buffer_sizes, void* const* outputs, const size_t* output_sizes)
{
	return decode_views(ctx, views, n, buffer_count, buffers, buffer_sizes, outputs, output_sizes, false, nullptr);
}

extern "C" int mob200_gltf_decode_device(mob200_Context* ctx, mob200_GltfView* views, size_t n, size_t buffer_count, const void* const* device_buffers, const size_t* buffer_sizes, void* const* device_outputs, const size_t* output_sizes, void* cuda_stream)
{
	return decode_views(ctx, views, n, buffer_count, device_buffers, buffer_sizes, device_outputs, output_sizes, true, cuda_stream);
}
