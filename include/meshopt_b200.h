/*
 * meshopt_b200.h -- C ABI of the B200-native meshoptimizer vertex-buffer decode path.
 *
 * Groups of entry points, all exported from meshoptimizer_b200/lib/libmeshopt_b200.so (1 and 2 are the hot
 * path; 3 and 4 are its callers' side: the index streams and the glTF bufferView loop around it):
 *
 *  1. DROP-IN symbols: exactly the signatures of the reference C API for this path, so an
 *     application (or an FFI binding) that links meshoptimizer can link this library instead:
 *       meshopt_decodeVertexBuffer   reference src/meshoptimizer.h:396  (impl. src/vertexcodec.cpp:1799-1872)
 *       meshopt_decodeVertexVersion  reference src/meshoptimizer.h:403  (impl. src/vertexcodec.cpp:1782-1797)
 *       meshopt_decodeFilterOct      reference src/meshoptimizer.h:421  (impl. src/vertexfilter.cpp:1211-1228)
 *       meshopt_decodeFilterQuat     reference src/meshoptimizer.h:422  (impl. src/vertexfilter.cpp:1230-1242)
 *       meshopt_decodeFilterExp      reference src/meshoptimizer.h:423  (impl. src/vertexfilter.cpp:1244-1255)
 *       meshopt_decodeFilterColor    reference src/meshoptimizer.h:424  (impl. src/vertexfilter.cpp:1257-1274)
 *     They take HOST pointers, are synchronous and thread-safe, and run on the GPU (there is no CPU
 *     fallback: without a usable CUDA device meshopt_decodeVertexBuffer returns MOB200_ERR_CUDA and
 *     the filters abort()).
 *
 *  2. DEVICE / BATCHED variant (new; prefix mob200_): many independent streams -- one descriptor per
 *     glTF bufferView, the shape gltf/parsegltf.cpp:561-627 iterates over -- decoded by one launch
 *     sequence with the decode filter fused, inputs and outputs resident in HBM.
 *
 * Plain C89-compatible declarations: pointers and sizes only, no CUDA or torch types.  A CUDA
 * stream is passed as void* (cudaStream_t); NULL means the default stream.
 */
#ifndef MESHOPT_B200_H
#define MESHOPT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MESHOPTIMIZER_API
#if defined(_WIN32)
#define MESHOPTIMIZER_API __declspec(dllexport)
#else
#define MESHOPTIMIZER_API __attribute__((visibility("default")))
#endif
#endif

/* ---- 1. drop-in symbols (reference src/meshoptimizer.h:396,403,421-424) ------------------------ */

/* Returns 0 on success, -1 bad header/version, -2 truncated or malformed data, -3 trailing bytes
 * (same codes as the reference, src/vertexcodec.cpp:1827-1869); MOB200_ERR_* (< -3) for failures
 * that have no reference equivalent. */
MESHOPTIMIZER_API int meshopt_decodeVertexBuffer(void* destination, size_t vertex_count, size_t vertex_size, const unsigned char* buffer, size_t buffer_size);
MESHOPTIMIZER_API int meshopt_decodeVertexVersion(const unsigned char* buffer, size_t buffer_size);

MESHOPTIMIZER_API void meshopt_decodeFilterOct(void* buffer, size_t count, size_t stride);
MESHOPTIMIZER_API void meshopt_decodeFilterQuat(void* buffer, size_t count, size_t stride);
MESHOPTIMIZER_API void meshopt_decodeFilterExp(void* buffer, size_t count, size_t stride);
MESHOPTIMIZER_API void meshopt_decodeFilterColor(void* buffer, size_t count, size_t stride);

/* ---- 2. device / batched variant --------------------------------------------------------------- */

#define MOB200_ERR_CUDA (-100)        /* no device, launch failure, out of memory */
#define MOB200_ERR_ARGUMENT (-101)    /* vertex_size not in (0,256] or not a multiple of 4, count >= 2^32, ... */

enum mob200_Filter
{
	MOB200_FILTER_NONE = 0,
	MOB200_FILTER_OCTAHEDRAL = 1, /* vertex_size 4 or 8  (meshopt_decodeFilterOct)   */
	MOB200_FILTER_QUATERNION = 2, /* vertex_size 8       (meshopt_decodeFilterQuat)  */
	MOB200_FILTER_EXPONENTIAL = 3, /* vertex_size % 4 == 0 (meshopt_decodeFilterExp)  */
	MOB200_FILTER_COLOR = 4       /* vertex_size 4 or 8  (meshopt_decodeFilterColor) */
};

/* One independent encoded stream = one call of meshopt_decodeVertexBuffer (+ optional filter).
 * src/dst are DEVICE pointers for the mob200_*_device entry points and HOST pointers for
 * mob200_decode_batch_host.  Device requirements: src must be readable in
 * [src & ~15, (src + src_size + 15) & ~15) (any cudaMalloc'ed range is); dst should be 16-byte
 * aligned for full-width stores (4-byte and 1-byte aligned destinations take slower paths). */
typedef struct mob200_Stream
{
	const unsigned char* src;
	size_t src_size;
	void* dst;
	size_t vertex_count;
	size_t vertex_size;
	int filter; /* enum mob200_Filter */
	int status; /* out: reference return code for this stream (written by *_status / *_host calls) */
} mob200_Stream;

typedef struct mob200_Context mob200_Context; /* per-device state: scratch arenas, staging, streams */
typedef struct mob200_Plan mob200_Plan;       /* a prepared batch: descriptors + scratch in HBM   */

/* device < 0 selects the current CUDA device.  Returns 0 or MOB200_ERR_CUDA. */
MESHOPTIMIZER_API int mob200_context_create(mob200_Context** out, int device);
/* A plan points at its context: destroy the plans of a context before the context. */
MESHOPTIMIZER_API void mob200_context_destroy(mob200_Context* ctx);

/* Prepare a batch of n streams whose src/dst are device pointers: uploads the descriptor table and
 * sizes the offset / look-back scratch.  Nothing is decoded yet. */
MESHOPTIMIZER_API int mob200_plan_create(mob200_Context* ctx, const mob200_Stream* streams, size_t n, mob200_Plan** out);
MESHOPTIMIZER_API void mob200_plan_destroy(mob200_Plan* plan);

/* Enqueue the fused walk + decode(+filter) kernel for the whole batch on `cuda_stream`; asynchronous.
 * May be called repeatedly (every call walks and decodes the batch again).
 * The kernel is persistent and its roles wait for each other across CTAs: all its CTAs must be resident at the same
 * time.  The launch therefore takes at most one CTA per SM of the device, and runs of plans that may overlap in time
 * (different streams) are serialised by the library with an event chain per context. */
MESHOPTIMIZER_API int mob200_plan_run(mob200_Plan* plan, void* cuda_stream);

/* Wait for the last run on `cuda_stream` and copy one reference return code per stream to
 * status[n] (host).  Returns the number of streams whose code is non-zero, or MOB200_ERR_CUDA. */
MESHOPTIMIZER_API int mob200_plan_status(mob200_Plan* plan, int* status, void* cuda_stream);

/* ---- 2b. block-offset sidecar: walk parallelism for few long streams ---------------------------------------------
 *
 * A stream stores no index, so finding where block b starts means walking blocks 0 .. b-1 (reference
 * src/vertexcodec.cpp:1857-1866 advances one pointer); one long stream is therefore one serial chain, however many
 * SMs there are.  The SIDECAR of a stream is the table of those offsets: mob200_sidecar_entries() = nblocks + 1
 * unsigned ints, entry b = byte offset of block b from the start of the stream (entry 0 is 1: the header byte),
 * entry nblocks = where the last block ends (= buffer_size - padded tail).  4 bytes per <= 8 KB of vertices.
 * With it every block is walked by its own GPU lane ("block mode"): one monolithic stream decodes as fast as
 * thousands of short ones.  A sidecar is never trusted: each lane checks that its block ends exactly where the
 * next one is said to start, block 0 at byte 1, the last one at the tail -- the verified blocks then are the very
 * chain the serial walk follows.  A stream whose sidecar does not fit (stale, corrupt, or the stream itself is
 * malformed) gets status MOB200_ERR_SIDECAR and must be decoded again without it to obtain the reference code.
 *
 * Where a sidecar comes from: (a) mob200_plan_export_sidecar after a run of the serial walk (store it next to the
 * asset; decode with it ever after); (b) the encoder-side helper mob200_encode_segments, which emits it;
 * (c) implicitly: any plan that has run once keeps its offsets, and mob200_plan_run_ex(.., MOB200_RUN_BLOCK_PARALLEL)
 * decodes the same plan again in block mode (re-verifying them: changed input bytes are caught).
 */
#define MOB200_ERR_SIDECAR (-102)         /* per-stream status: the block-offset sidecar does not describe this stream */
#define MOB200_RUN_BLOCK_PARALLEL 1       /* mob200_plan_run_ex flag: walk every block from the plan's offset table */

/* nblocks + 1 (0 for an empty stream or an illegal vertex size) */
MESHOPTIMIZER_API size_t mob200_sidecar_entries(size_t vertex_count, size_t vertex_size);

/* mob200_plan_create with one HOST sidecar pointer per stream (sidecars[i] has mob200_sidecar_entries(..) entries; it
 * may be NULL only for streams without vertices -- if any other stream lacks one the plan is created without
 * offsets).  The offsets are uploaded once; mob200_plan_run_ex(.., MOB200_RUN_BLOCK_PARALLEL) uses them. */
MESHOPTIMIZER_API int mob200_plan_create_sidecar(mob200_Context* ctx, const mob200_Stream* streams, size_t n, const unsigned int* const* sidecars, mob200_Plan** out);

/* mob200_plan_run with flags.  MOB200_RUN_BLOCK_PARALLEL needs offsets in the plan (mob200_plan_has_offsets), else
 * MOB200_ERR_ARGUMENT.  Status codes of a block-mode run: 0, the framing codes -1 / -2 / -3 of streams whose
 * header, size or tail is wrong, and MOB200_ERR_SIDECAR. */
MESHOPTIMIZER_API int mob200_plan_run_ex(mob200_Plan* plan, void* cuda_stream, int flags);
MESHOPTIMIZER_API int mob200_plan_has_offsets(const mob200_Plan* plan);

/* Copy the sidecar of stream `stream_index` (caller order) out of the plan after a run on `cuda_stream`
 * (synchronises on it).  Returns the number of entries written, MOB200_ERR_ARGUMENT if capacity is too small, or
 * MOB200_ERR_SIDECAR if that stream's walk failed (its offsets are not a sidecar). */
MESHOPTIMIZER_API int mob200_plan_export_sidecar(mob200_Plan* plan, size_t stream_index, unsigned int* out, size_t capacity, void* cuda_stream);

/* Number of kernels one mob200_plan_run enqueues (for launch accounting). */
MESHOPTIMIZER_API int mob200_plan_launches(const mob200_Plan* plan);

/* Convenience: create plan, run, fetch status into streams[i].status, destroy.  Synchronous. */
MESHOPTIMIZER_API int mob200_decode_batch_device(mob200_Context* ctx, mob200_Stream* streams, size_t n, void* cuda_stream);

/* Same contract with HOST pointers in src/dst: compressed bytes go host->device, decoded (and
 * filtered) vertices come device->host, pipelined in chunks over several CUDA streams.  Pinned
 * (page-locked) caller buffers are transferred in place, pageable ones through pinned staging.
 * Synchronous; fills streams[i].status (a stream with illegal arguments gets MOB200_ERR_ARGUMENT and the
 * rest of the batch is still decoded); returns the number of failed streams or MOB200_ERR_*.  Nothing is
 * left in flight when the call returns, also on failure. */
MESHOPTIMIZER_API int mob200_decode_batch_host(mob200_Context* ctx, mob200_Stream* streams, size_t n);

/* mob200_decode_batch_host with one host block-offset sidecar per stream (section 2b; NULL array = none): chunks
 * whose streams all carry one are decoded in block mode. */
MESHOPTIMIZER_API int mob200_decode_batch_host_sidecar(mob200_Context* ctx, mob200_Stream* streams, size_t n, const unsigned int* const* sidecars);

/* ---- several GPUs of one box: streams are independent, nothing is exchanged -------------- */

/* Greedy longest-processing-time partition: costs[i] (algorithmic bytes of stream i) -> rank_of[i] in [0, world). */
MESHOPTIMIZER_API int mob200_shard_streams(const size_t* costs, size_t n, int world, int* rank_of);

/* mob200_decode_batch_host_sidecar over n_devices GPUs: the batch is partitioned by mob200_shard_streams on
 * src_size + vertex_count * vertex_size, every device gets its own host thread and context (kept for later calls) and
 * decodes its shard concurrently.  device_ms (optional, n_devices entries): wall time of each device's shard.
 * Returns the number of failed streams or MOB200_ERR_*. */
MESHOPTIMIZER_API int mob200_decode_batch_multi_host(const int* devices, int n_devices, mob200_Stream* streams, size_t n, const unsigned int* const* sidecars, float* device_ms);

/* In-place decode filter on a DEVICE buffer of count elements (asynchronous on cuda_stream).
 * filter is enum mob200_Filter (not NONE); stride rules as the reference asserts them. */
MESHOPTIMIZER_API int mob200_filter_device(int filter, void* device_buffer, size_t count, size_t stride, void* cuda_stream);

/* Library / device introspection used by tests and bench: number of SMs of the context's device,
 * and a version string. */
MESHOPTIMIZER_API int mob200_context_sm_count(const mob200_Context* ctx);
MESHOPTIMIZER_API const char* mob200_version(void);

/* Duration in milliseconds of the most recent mob200_plan_run: ONE fused persistent kernel (walker, producer and
 * decoder warps), measured with a pair of CUDA events recorded on the launching stream around the launch.
 * Synchronises on the end event.  Returns 0 or MOB200_ERR_CUDA. */
MESHOPTIMIZER_API int mob200_plan_last_timing(mob200_Plan* plan, float* ms);

/* Same for the most recent runs (at most 64 are remembered, oldest first): fills up to max_runs entries of ms and
 * returns how many were written.  Nothing is synchronised until this call, so a timed region of repeated
 * mob200_plan_run calls stays asynchronous. */
MESHOPTIMIZER_API int mob200_plan_timing_history(mob200_Plan* plan, int max_runs, float* ms);

/* Host milliseconds mob200_plan_create spent on this plan (validation, sort, decode order, table uploads): the
 * one-off cost a single-shot decode pays in front of the kernel. */
MESHOPTIMIZER_API float mob200_plan_create_ms(const mob200_Plan* plan);

/* Diagnostics: cycle counters the kernel accumulates over all CTAs since the last reset -- [0] decoder
 * warps total, [1..3] of which waiting for staged data / the cross-block carry / the output tile, [4]
 * producer warps total, [5..7] of which in block metadata / waiting for a free slot / in the look-back,
 * [8] walker warps total, [9..10] of which waiting for ring refills, [11..12] refill points / forced waits.
 * count <= 16.
 * Synchronises the device.  Returns 0 or MOB200_ERR_*. */
MESHOPTIMIZER_API int mob200_plan_debug_counters(mob200_Plan* plan, unsigned long long* out, int count, int reset);

/* ---- 2c. encoder-side helper: the segmenter -----------------------------------------
 *
 * Host code for the asset pipeline that feeds the GPU decoder: vertex arrays are encoded in the reference wire format
 * (reference src/vertexcodec.cpp:1615-1693; every stream made here is a valid input of the unmodified
 * meshopt_decodeVertexBuffer, and has the size the reference encoder gives it at the same level) as independently
 * decodable segments, each with its block-offset sidecar (section 2b) -- the two things the decoder's parallelism
 * comes from.  No CUDA is involved. */

/* Upper bound of the encoded size of one stream (the reference's meshopt_encodeVertexBufferBound, :1748-1768). */
MESHOPTIMIZER_API size_t mob200_encode_vertex_bound(size_t vertex_count, size_t vertex_size);

/* Encode one stream: version 0 or 1, level 0..3 (meaning as meshopt_encodeVertexBufferLevel, :1615).  sidecar: NULL, or
 * mob200_sidecar_entries(vertex_count, vertex_size) entries that receive the stream's block offsets.  Returns the encoded
 * size, 0 if buffer_size is too small or an argument is illegal. */
MESHOPTIMIZER_API size_t mob200_encode_vertex_buffer(unsigned char* buffer, size_t buffer_size, const void* vertices, size_t vertex_count, size_t vertex_size, int level, int version, unsigned int* sidecar);

typedef struct mob200_Segment
{
	size_t first_vertex, vertex_count; /* the vertices this stream holds */
	size_t offset, size;               /* the stream inside the output blob (offset is a multiple of 16) */
	size_t sidecar_offset, sidecar_entries; /* its block offsets inside the sidecar array (entries, not bytes) */
} mob200_Segment;

/* ceil(vertex_count / segment_vertices); segment_vertices 0 = one segment */
MESHOPTIMIZER_API size_t mob200_segment_count(size_t vertex_count, size_t segment_vertices);
/* capacities mob200_encode_segments needs: bytes of `out`, entries of `sidecars` */
MESHOPTIMIZER_API size_t mob200_encode_segments_bound(size_t vertex_count, size_t vertex_size, size_t segment_vertices);
MESHOPTIMIZER_API size_t mob200_segments_sidecar_entries(size_t vertex_count, size_t vertex_size, size_t segment_vertices);

/* Split a vertex array into segments of segment_vertices vertices and encode each as its own stream on `threads` host
 * threads (0 = all): streams packed back to back on 16-byte boundaries in `out` (*out_size bytes used), one
 * mob200_Segment per stream, and (sidecars != NULL) every stream's block offsets.  segments[i] maps directly onto a
 * mob200_Stream {out + offset, size, dst + first_vertex * vertex_size, vertex_count, vertex_size} and onto the sidecar
 * pointer sidecars + sidecar_offset.  Returns the number of segments or MOB200_ERR_ARGUMENT. */
MESHOPTIMIZER_API int mob200_encode_segments(const void* vertices, size_t vertex_count, size_t vertex_size, size_t segment_vertices, int level, int version, int threads,
    unsigned char* out, size_t out_capacity, size_t* out_size, mob200_Segment* segments, size_t segment_capacity, unsigned int* sidecars, size_t sidecar_capacity);

/* ---- 3. index streams (the other two modes of a compressed glTF bufferView) ---------------------- */

/* Drop-in symbols, reference src/meshoptimizer.h:344 (meshopt_decodeIndexBuffer, impl. src/indexcodec.cpp:384-576),
 * :351 (meshopt_decodeIndexVersion, impl. :364-382) and :376 (meshopt_decodeIndexSequence, impl. :647-703).
 * HOST pointers, synchronous; 0 / -1 / -2 / -3 as the reference, MOB200_ERR_* otherwise.  index_size is 2 or 4;
 * index_count of a triangle list is a multiple of 3 (the reference asserts both; here MOB200_ERR_ARGUMENT). */
MESHOPTIMIZER_API int meshopt_decodeIndexBuffer(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);
MESHOPTIMIZER_API int meshopt_decodeIndexVersion(const unsigned char* buffer, size_t buffer_size);
MESHOPTIMIZER_API int meshopt_decodeIndexSequence(void* destination, size_t index_count, size_t index_size, const unsigned char* buffer, size_t buffer_size);

enum mob200_IndexKind
{
	MOB200_INDEX_TRIANGLES = 0, /* meshopt_decodeIndexBuffer   (glTF mode "TRIANGLES") */
	MOB200_INDEX_SEQUENCE = 1   /* meshopt_decodeIndexSequence (glTF mode "INDICES")   */
};

/* One independent index stream.  The formats are sequential state machines, so a batch is decoded with one
 * thread per stream: throughput comes from the number of streams (meshes, meshlets), not from their length. */
typedef struct mob200_IndexStream
{
	const unsigned char* src;
	size_t src_size;
	void* dst; /* index_count * index_size bytes, aligned to index_size */
	size_t index_count;
	size_t index_size;
	int kind;   /* enum mob200_IndexKind */
	int status; /* out: reference return code */
} mob200_IndexStream;

/* src/dst are device pointers; one kernel launch on cuda_stream; synchronises on it to fetch the codes into
 * streams[i].status.  Returns the number of failed streams or MOB200_ERR_*. */
MESHOPTIMIZER_API int mob200_decode_index_batch_device(mob200_Context* ctx, mob200_IndexStream* streams, size_t n, void* cuda_stream);
/* Same with host pointers (one staging round trip for the batch).  Synchronous and thread-safe. */
MESHOPTIMIZER_API int mob200_decode_index_batch_host(mob200_Context* ctx, mob200_IndexStream* streams, size_t n);

/* ---- 3b. meshlets (reference src/meshoptimizer.h:349-350; impl. src/meshletcodec.cpp:981-1051) --------- */

/* Drop-in symbols: HOST pointers, synchronous; 0 / -2 / -3 as the reference.  vertex_count, triangle_count <= 256,
 * vertex_size 2 or 4, triangle_size 3 or 4 (the reference asserts these; here MOB200_ERR_ARGUMENT).  The raw form
 * writes exactly vertex_count / triangle_count 32-bit elements (the reference may also write the padding elements). */
MESHOPTIMIZER_API int meshopt_decodeMeshlet(void* vertices, size_t vertex_count, size_t vertex_size, void* triangles, size_t triangle_count, size_t triangle_size, const unsigned char* buffer, size_t buffer_size);
MESHOPTIMIZER_API int meshopt_decodeMeshletRaw(unsigned int* vertices, size_t vertex_count, unsigned int* triangles, size_t triangle_count, const unsigned char* buffer, size_t buffer_size);

/* One encoded meshlet; decoded by one GPU thread, so batches of many meshlets are the intended use. */
typedef struct mob200_Meshlet
{
	const unsigned char* src;
	size_t src_size;
	void* vertices; /* vertex_count * vertex_size bytes, aligned to vertex_size */
	size_t vertex_count;
	size_t vertex_size;
	void* triangles; /* triangle_count * triangle_size bytes (4-byte aligned when triangle_size is 4) */
	size_t triangle_count;
	size_t triangle_size;
	int status; /* out: reference return code */
} mob200_Meshlet;

MESHOPTIMIZER_API int mob200_decode_meshlet_batch_device(mob200_Context* ctx, mob200_Meshlet* meshlets, size_t n, void* cuda_stream);
MESHOPTIMIZER_API int mob200_decode_meshlet_batch_host(mob200_Context* ctx, mob200_Meshlet* meshlets, size_t n);

/* Diagnostics (environment MOB200_TIMING=1): milliseconds of the decode kernel of the most recent meshlet batch of this
 * process (CUDA events around the launch), without the descriptor upload and the status read-back. */
MESHOPTIMIZER_API float mob200_debug_last_kernel_ms(void);

/* ---- 4. glTF bufferView front-end (reference gltf/parsegltf.cpp:561-627, decompressMeshopt) ------ */

enum mob200_GltfMode
{
	MOB200_GLTF_ATTRIBUTES = 0, /* -> meshopt_decodeVertexBuffer + filter */
	MOB200_GLTF_TRIANGLES = 1,  /* -> meshopt_decodeIndexBuffer */
	MOB200_GLTF_INDICES = 2     /* -> meshopt_decodeIndexSequence */
};

/* One bufferView that carries EXT_meshopt_compression / KHR_meshopt_compression (field names as written by
 * gltf/write.cpp:739-790): the compressed bytes are src_size bytes at src_offset of buffers[src_buffer]; the
 * decompressed view is dst_size = count * stride bytes at dst_offset of buffers[dst_buffer] (the buffer the
 * view itself names: gltfpack's "fallback" buffer). */
typedef struct mob200_GltfView
{
	size_t view; /* index in bufferViews[] */
	int mode;    /* enum mob200_GltfMode */
	int filter;  /* enum mob200_Filter */
	size_t src_buffer, src_offset, src_size;
	size_t count, stride;
	size_t dst_buffer, dst_offset, dst_size;
	int status; /* scan: 0 or MOB200_ERR_ARGUMENT (violates the extension's rules, extern/cgltf.h:1645-1667);
	               decode: reference return code of the view's codec */
} mob200_GltfView;

typedef struct mob200_GltfInfo
{
	size_t json_offset, json_size; /* JSON text inside the input */
	size_t bin_offset, bin_size;   /* BIN chunk of a .glb (the bytes of buffers[0]); 0, 0 for bare JSON */
	size_t buffer_count;           /* length of buffers[] */
	size_t view_count;             /* compressed views found (may exceed view_capacity: call again) */
	size_t invalid_views;          /* of which rejected by the extension's rules */
} mob200_GltfInfo;

/* Find the compressed bufferViews of a .glb container or of bare .gltf JSON text.  Writes up to view_capacity
 * views and up to buffer_capacity byteLength values of buffers[] (either array may be NULL).  Views that break the
 * extension's rules, name a buffer the asset does not have, or (when buffer_sizes is given) reach outside their
 * buffers are returned with status MOB200_ERR_ARGUMENT.  JSON nested deeper than 64 levels is rejected.  Host only,
 * no device work.  Returns 0, or MOB200_ERR_ARGUMENT for a malformed container / JSON. */
MESHOPTIMIZER_API int mob200_gltf_scan(const void* data, size_t size, mob200_GltfView* views, size_t view_capacity, size_t* buffer_sizes, size_t buffer_capacity, mob200_GltfInfo* info);

/* Decode all views in one batch per codec family.  The asset has buffer_count buffers; buffers[i] (buffer_sizes[i]
 * bytes) is where buffers[i] of the asset is loaded, outputs[i] (output_sizes[i] bytes) is where decompressed views
 * that belong to buffers[i] go.  All four arrays have buffer_count entries and are mandatory; NULL pointers in
 * buffers[] / outputs[] make the views that need them fail.  Every view is checked against these bounds before
 * anything is read or written: a view that names a buffer >= buffer_count, or whose source or destination range
 * does not lie inside its buffer, gets MOB200_ERR_ARGUMENT (views come from untrusted JSON).  Host pointers,
 * synchronous.  Returns the number of views whose status is non-zero, or MOB200_ERR_*. */
MESHOPTIMIZER_API int mob200_gltf_decode_host(mob200_Context* ctx, mob200_GltfView* views, size_t n, size_t buffer_count, const void* const* buffers, const size_t* buffer_sizes, void* const* outputs, const size_t* output_sizes);
/* Same with device pointers in buffers[] / outputs[]. */
MESHOPTIMIZER_API int mob200_gltf_decode_device(mob200_Context* ctx, mob200_GltfView* views, size_t n, size_t buffer_count, const void* const* device_buffers, const size_t* buffer_sizes, void* const* device_outputs, const size_t* output_sizes, void* cuda_stream);

#ifdef __cplusplus
}
#endif

#endif /* MESHOPT_B200_H */
