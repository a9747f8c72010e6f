/* vrq.h - C ABI of libvrq.so: B200 (sm_100a) kernels for the embedding-quantisation and multi-phase
 * search hot path of aitrailblazer/VectorRAGQuantization.
 *
 * The reference is pure Python; it has no FFI of its own.  The boundary a maintainer binds is therefore
 *   (1) the private static NumPy kernels of the VectorDB* classes            -> vrq_quantize_* / vrq_to_binary_* / vrq_dequantize_*
 *   (2) the faiss surface those classes call (IndexBinaryIDMap2(IndexBinaryFlat)) -> vrq_index_*
 *   (3) the per-candidate Python rescoring loops inside search()             -> vrq_index_search3 / vrq_index_search2
 * Each entry point cites the reference lines (file:line under /root/reference) it replaces.
 * INTEGRATION.md shows the ctypes stubs that slot these into the reference classes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every data pointer may be a HOST pointer or a DEVICE pointer (all data pointers of one call in the
 *     same space; the library asks the CUDA runtime which).  Host buffers: the call runs a chunked two-stream pipeline
 *     (H2D of chunk i+1 and D2H of chunk i-1 around the kernel of chunk i) and returns when the outputs are complete.  The
 *     copies are issued straight from / to the caller's buffers: they overlap the kernels when those buffers are pinned
 *     (cudaHostAlloc, torch pin_memory); with pageable memory the runtime stages them itself and the steps serialise.
 *     Device buffers: the call only enqueues work on the context's stream (vrq_ctx_set_stream) and returns.
 *   - Return value: 0 = OK, > 0 = a cudaError_t, < 0 = VRQ_ERR_*.  vrq_last_error() gives the text.
 *   - One vrq_ctx per GPU; a ctx (and the indexes made from it) must not be used from two threads at once.
 *   - Rows are C-contiguous.  d must be a multiple of 8 (faiss's own requirement for binary indexes).
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef VRQ_H
#define VRQ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRQ_VERSION 100

#define VRQ_ERR_ARG (-1)         /* bad argument (null pointer, d % 8 != 0, k <= 0, mixed host/device ...) */
#define VRQ_ERR_UNSUPPORTED (-2) /* valid request outside what the kernels cover (d > 8192, k too large)   */
#define VRQ_ERR_IO (-3)          /* index file could not be read / written / parsed                       */
#define VRQ_ERR_STATE (-4)       /* e.g. search3 on an index without an int8 payload                      */
#define VRQ_ERR_NOMEM (-5)

typedef struct vrq_ctx vrq_ctx;
typedef struct vrq_index vrq_index;

int vrq_version(void);
const char* vrq_last_error(void);

/* ---------------------------------------------------------------- context ------------------------- */
int vrq_ctx_create(int device, vrq_ctx** out);
int vrq_ctx_destroy(vrq_ctx* ctx);
/* Run all subsequent work of this ctx on the given cudaStream_t (e.g. torch's current stream; NULL = CUDA's legacy
 * default stream, which is torch's default stream).  A new ctx runs on its own non-blocking stream;
 * vrq_ctx_reset_stream goes back to it. */
int vrq_ctx_set_stream(vrq_ctx* ctx, void* cuda_stream);
int vrq_ctx_reset_stream(vrq_ctx* ctx);
int vrq_ctx_sync(vrq_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t vrq_ctx_launch_count(const vrq_ctx* ctx);
int vrq_ctx_device(const vrq_ctx* ctx);

/* ---------------------------------------------------------------- encoders ------------------------
 * x: float32[n, d].  ubin (nullable): uint8[n, d/8] = np.packbits(x > np.mean(x)) of the same row, fused
 * into the same pass over x (the reference calls _quantize_* and _to_binary back to back on every
 * embedding: VectorDBInt8.py:103-106).  Device pointers: x and q 16-byte aligned, ubin 4-byte aligned (any torch /
 * cudaMalloc allocation is; VRQ_ERR_ARG otherwise).  Host pointers: no alignment requirement. */

/* VectorDBInt8._quantize_to_int8 (VectorDBInt8.py:114-126): scale = f32(127)/max(|min|,|max|) (f32 divide),
 * q = trunc(x*scale); max == min -> zeros.  q int8[n,d]; mn, mx float32[n]. */
int vrq_quantize_int8_perdoc(vrq_ctx*, const float* x, int64_t n, int d, int8_t* q, float* mn, float* mx, uint8_t* ubin);
/* VectorDBInt8Global._quantize_to_int8 (VectorDBInt8Global.py:130-142): clip(+-f32(limit)), * f32(127.0/limit),
 * round-half-even, clip +-127. */
int vrq_quantize_int8_global(vrq_ctx*, const float* x, int64_t n, int d, double limit, int8_t* q, uint8_t* ubin);
/* VectorDBInt16Global._quantize_to_int16 (VectorDBInt16Global.py:130-142), 32767. */
int vrq_quantize_int16_global(vrq_ctx*, const float* x, int64_t n, int d, double limit, int16_t* q, uint8_t* ubin);
/* VectorDBInt4._quantize_to_int4 (VectorDBInt4.py:116-154) and VectorDBInt4Global._quantize_to_int4
 * (VectorDBInt4Global.py:129-164 - its limit argument is unused by the reference, so there is none here):
 * scale = f32(7.0/f64(max(|min|,|max|))), rint, clip [-8,7], byte i = ((s[2i]+8)<<4)|(s[2i+1]+8).
 * packed int8[n, d/2]; mn, mx (nullable) float64[n] = float(np.min), float(np.max). */
int vrq_quantize_int4(vrq_ctx*, const float* x, int64_t n, int d, int8_t* packed, double* mn, double* mx, uint8_t* ubin);

/* _to_binary (VectorDBInt8.py:140-146 + 4 siblings): packbits(x > mean_f32_pairwise(x)).  ge != 0 selects
 * CohereVectorDBBinary's '>=' (CohereVectorDBBinary.py:133-151). */
int vrq_to_binary_f32(vrq_ctx*, const float* x, int64_t n, int d, int ge, uint8_t* ubin);
/* _to_binary on integer embeddings (CohereVectorDBInt8.py:130-135, VectorDBInt16.py:148-157,
 * CohereEnhancedVectorDB.py:130-134): exact integer form d*x > sum(x). */
int vrq_to_binary_i8(vrq_ctx*, const int8_t* x, int64_t n, int d, int ge, uint8_t* ubin);
int vrq_to_binary_i16(vrq_ctx*, const int16_t* x, int64_t n, int d, int ge, uint8_t* ubin);

/* ---------------------------------------------------------------- decoders ------------------------ */
/* VectorDBInt8._dequantize_int8 (VectorDBInt8.py:128-138): f32(q) * (max(|min|,|max|)/127 in f32); zeros if max==min */
int vrq_dequantize_int8_perdoc(vrq_ctx*, const int8_t* q, int64_t n, int d, const float* mn, const float* mx, float* out);
/* VectorDBInt8Global._dequantize_int8 (VectorDBInt8Global.py:144-152): f32(q) * f32(limit/127.0) */
int vrq_dequantize_int8_global(vrq_ctx*, const int8_t* q, int64_t n, int d, double limit, float* out);
/* VectorDBInt16Global._dequantize_int16 (VectorDBInt16Global.py:144-152) */
int vrq_dequantize_int16_global(vrq_ctx*, const int16_t* q, int64_t n, int d, double limit, float* out);
/* VectorDBInt4._dequantize_int4 (VectorDBInt4.py:156-184): f32((nibble-8) * (max(|min|,|max|)/7.0 in f64)); zeros if
 * max==min.  (The reference loop raises OverflowError on NumPy >= 2; this is its NumPy-1.x result.) */
int vrq_dequantize_int4_perdoc(vrq_ctx*, const int8_t* packed, int64_t n, int d, const double* mn, const double* mx, float* out);
/* VectorDBInt4Global._dequantize_int4 (VectorDBInt4Global.py:166-188): f32((nibble-8) * (limit/7.0)) */
int vrq_dequantize_int4_global(vrq_ctx*, const int8_t* packed, int64_t n, int d, double limit, float* out);

/* ---------------------------------------------------------------- binary index --------------------
 * Device-resident equivalent of faiss.IndexBinaryIDMap2(faiss.IndexBinaryFlat(d)) as constructed at
 * CohereEnhancedVectorDB.py:126 / VectorDBInt8.py:69: codes uint8[ntotal, d/8] + int64 id per position. */
int vrq_index_create(vrq_ctx*, int d, vrq_index** out);
int vrq_index_free(vrq_index*);
int64_t vrq_index_ntotal(const vrq_index*); /* .ntotal  (CohereEnhancedVectorDB.py:247,267) */
int vrq_index_d(const vrq_index*);
int vrq_index_reserve(vrq_index*, int64_t capacity_rows);
/* Device addresses of the resident arrays (codes u8[ntotal,d/8], ids i64[ntotal] or NULL while ids are implicit,
 * payload rows, aux pairs) for callers that launch the stand-alone kernels on them; invalidated by add / remove. */
int vrq_index_device_ptrs(vrq_index*, void** codes, void** ids, void** payload, void** aux);

/* Optional per-position payload stored beside the codes (what the reference keeps in RocksDB and fetches one
 * pickle at a time inside its rescoring loops: CohereEnhancedVectorDB.py:303, VectorDBInt8.py:228). */
#define VRQ_PAYLOAD_NONE 0
#define VRQ_PAYLOAD_INT8_RAW 1     /* int8[d]              CohereEnhancedVectorDB {"int8"}        */
#define VRQ_PAYLOAD_INT8_PERDOC 2  /* int8[d] + f32 min,max  VectorDBInt8 {"emb_int8","min_max"}    */
#define VRQ_PAYLOAD_INT8_GLOBAL 3  /* int8[d], limit         VectorDBInt8Global                     */
#define VRQ_PAYLOAD_INT16_GLOBAL 4 /* int16[d], limit        VectorDBInt16Global                    */
#define VRQ_PAYLOAD_INT4_PERDOC 5  /* int8[d/2] + f64 min,max VectorDBInt4                          */
#define VRQ_PAYLOAD_INT4_GLOBAL 6  /* int8[d/2], limit       VectorDBInt4Global                     */
#define VRQ_PAYLOAD_F32 7          /* float32[d]             the float_embeddings dict (compare_float32=True) */
#define VRQ_PAYLOAD_CODES_PM1 8    /* no extra rows: the 1-bit code itself unpacked to +-1.0f (CohereVectorDBBinary.py:153-159, :227) */
int vrq_index_set_payload(vrq_index*, int kind, double global_limit); /* only while ntotal == 0 */
int vrq_index_payload_kind(const vrq_index*);

/* .add_with_ids (CohereEnhancedVectorDB.py:217, VectorDBInt8.py:175).  payload / aux are required iff a payload
 * kind is set: payload = n rows of the kind's row type; aux = n x {min,max} (f32 pairs for INT8_PERDOC, f64
 * pairs for INT4_PERDOC), NULL otherwise. */
int vrq_index_add_with_ids(vrq_index*, int64_t n, const uint8_t* codes, const int64_t* ids, const void* payload,
                           const void* aux);
/* .search (CohereEnhancedVectorDB.py:268, VectorDBInt8.py:218): the k codes with smallest (hamming, position),
 * ascending; dist int32[nq,k], labels int64[nq,k]; ntotal < k pads with (INT32_MAX, -1). */
int vrq_index_search(vrq_index*, int64_t nq, const uint8_t* q, int k, int32_t* dist, int64_t* labels);
/* Every Hamming distance, dist int32[nq, ntotal] (what IndexBinaryFlat computes internally before its heap), straight
 * from the tensor-core scan's accumulators.  d must be 1024; VRQ_ERR_UNSUPPORTED otherwise.  Used by the parity tests to
 * check the int8 MMA contraction element by element; small ntotal only (the output is nq * ntotal * 4 bytes). */
int vrq_index_distances(vrq_index*, int64_t nq, const uint8_t* q, int32_t* dist);
/* .reconstruct(id) (CohereEnhancedVectorDB.py:286); last added wins on duplicate ids; host output. */
int vrq_index_reconstruct(vrq_index*, int64_t id, uint8_t* code_out);
/* .remove_ids (CohereEnhancedVectorDB.py:334): returns the number of rows removed (>= 0; every row carrying a listed id,
 * as IDMap2 does).  The removal is recorded and the device arrays are compacted once - in place, order-preserving like
 * faiss, through a 64 MB bounce buffer - before the next call that reads them, so the reference's one-id-per-call
 * remove_document loop costs one pass instead of one O(N) pass per id.  vrq_index_ntotal reflects it immediately. */
int64_t vrq_index_remove_ids(vrq_index*, int64_t n, const int64_t* ids);
/* faiss.write_index_binary / read_index_binary (CohereEnhancedVectorDB.py:346,123): byte-compatible "IBM2"/"IBxF".
 * Both stream between the file and device memory in 64 MB chunks (13.6 GB at 100 M codes never sits in host memory);
 * the writer goes through "<path>.tmp" + fsync + rename, so a crash leaves the old file or the new one. */
int vrq_index_write(vrq_index*, const char* path);
int vrq_index_read(vrq_ctx*, const char* path, vrq_index** out);
/* The quantised vectors the reference keeps in RocksDB pickles beside index.bin (CohereEnhancedVectorDB.py:221 {"int8"},
 * VectorDBInt8.py:179 {"emb_int8","min_max"}) as one flat sidecar file ("VRQP": header, payload rows, aux rows), streamed
 * the same way.  read_payload attaches the rows to an index that came from vrq_index_read (same ntotal, no payload yet). */
int vrq_index_write_payload(vrq_index*, const char* path);
int vrq_index_read_payload(vrq_index*, const char* path);
/* Give a payload-less index (e.g. one read from a reference-written index.bin) zero-filled payload rows of `kind`, to be
 * filled with vrq_index_write_rows: how the importer of the reference's docs/ store re-attaches its int8 vectors. */
int vrq_index_attach_payload(vrq_index*, int kind, double global_limit);
/* `count` consecutive rows starting at position `offset` of one resident array, to / from a host or device buffer. */
#define VRQ_ROWS_CODES 0   /* uint8[d/8]              */
#define VRQ_ROWS_IDS 1     /* int64                   */
#define VRQ_ROWS_PAYLOAD 2 /* row type of the payload kind */
#define VRQ_ROWS_AUX 3     /* {min,max} pairs         */
int vrq_index_read_rows(vrq_index*, int which, int64_t offset, int64_t count, void* out);
int vrq_index_write_rows(vrq_index*, int which, int64_t offset, int64_t count, const void* src); /* codes / payload / aux */
/* Copy payload rows (and aux) of the given positions to host/device buffers (doc_db.get replacement). */
int vrq_index_get_payload(vrq_index*, int64_t m, const int64_t* positions, void* payload_out, void* aux_out);
int64_t vrq_index_position_of(vrq_index*, int64_t id); /* -1 if absent */

/* CohereEnhancedVectorDB.search phases I-III (CohereEnhancedVectorDB.py:267-322) for nq queries at once.
 * Needs VRQ_PAYLOAD_INT8_RAW.  q_float float32[nq,d], q_ubin uint8[nq,d/8].  Outputs, k entries per query
 * (first out_count[q] valid, rest -1 / -inf): labels int64, hamming int32, score_binary f64, score_cosine f64. */
int vrq_index_search3(vrq_index*, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k, int binary_oversample,
                      int int8_oversample, int64_t* labels, int32_t* hamming, double* score_binary, double* score_cosine,
                      int32_t* out_count);
/* The 2-phase search of the six VectorDB* classes (VectorDBInt8.py:213-242): Hamming top min(k*oversample, ntotal),
 * float32 dot(q, dequantised payload) for every hit, stable sort descending, [:k].  score float32[nq,k].
 * q_ubin may be NULL: the 1-bit code of every query is then computed on the device as the classes do it,
 * query_bin = _to_binary(query float) = packbits(q > mean(q)) (VectorDBInt8.py:213, :140-146). */
int vrq_index_search2(vrq_index*, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k, int binary_oversample,
                      int64_t* labels, float* score, int32_t* out_count);

/* CohereVectorDBFloat (the float32 recall baseline): faiss.IndexIDMap(faiss.IndexFlatIP(d)) = an index whose payload kind is
 * VRQ_PAYLOAD_F32 (codes may then be NULL in vrq_index_add_with_ids).  .search (CohereVectorDBFloat.py:156): the k rows with the
 * largest float32 inner product with each query, descending (ties: lower position first); scores float32[nq,k], labels
 * int64[nq,k], padded with (-inf, -1).  k <= 4096. */
int vrq_index_search_ip(vrq_index*, int64_t nq, const float* q_float, int k, float* scores, int64_t* labels);
/* faiss.write_index / read_index of that index (CohereVectorDBFloat.py:184,58): byte-compatible "IxMp" wrapping "IxFI". */
int vrq_index_write_float(vrq_index*, const char* path);
int vrq_index_read_float(vrq_ctx*, const char* path, vrq_index** out);

/* ---------------------------------------------------------------- multi-GPU pieces (device pointers) ----
 * Row-sharded database: every rank runs search3_local on its shard, the host all-gathers the three arrays over
 * NCCL, every rank (or rank 0) runs merge3.  pos_base = global position of this shard's row 0. */
int vrq_index_search3_local(vrq_index*, int64_t nq, const float* q_float, const uint8_t* q_ubin, int binary_k,
                            int64_t pos_base, uint64_t* keys /*[nq,binary_k] (hamming<<40 | global pos), ~0 = none*/,
                            int64_t* labels /*[nq,binary_k]*/, double* score_binary, double* score_cosine);
/* rank_stride: elements between consecutive ranks' [nq,binary_k] blocks inside each of the four input arrays
 * (0 = nq*binary_k, i.e. dense [world,nq,binary_k]); lets one packed all-gather buffer feed the merge directly. */
int vrq_merge3(vrq_ctx*, int world, int64_t nq, int binary_k, int64_t rank_stride, const uint64_t* keys,
               const int64_t* labels, const double* score_binary, const double* score_cosine, int k, int k2,
               int64_t* out_labels, int32_t* out_hamming, double* out_score_binary, double* out_score_cosine,
               int32_t* out_count);

/* The same exchange inside the library (sharded.cu), for hosts without torch: NCCL is loaded with dlopen at first use
 * (libnccl.so.2), libvrq.so itself links against libcudart only.
 *   one process per GPU:  vrq_nccl_unique_id on rank 0 -> ship the 128 bytes -> vrq_nccl_init_rank on every rank ->
 *                         vrq_ctx_set_nccl -> vrq_index_search3_sharded on every rank (same queries everywhere);
 *   one process, all GPUs: vrq_nccl_init_all -> vrq_ctx_set_nccl per context -> vrq_search3_sharded_group.
 * All data pointers are DEVICE pointers on the rank's GPU; work is enqueued on the context's stream.  pos_base = global
 * position of the shard's row 0, ntotal_global = rows over all shards (binary_k = min(k * oversample, ntotal_global)). */
int vrq_nccl_unique_id(char* id128);
int vrq_nccl_init_rank(int device, int world, const char* id128, int rank, void** comm_out);
int vrq_nccl_init_all(int ndev, const int* devices, void** comms_out);
int vrq_nccl_destroy(void* comm);
int vrq_ctx_set_nccl(vrq_ctx*, void* nccl_comm, int rank, int world);
int vrq_index_search3_sharded(vrq_index*, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k, int binary_oversample,
                              int int8_oversample, int64_t pos_base, int64_t ntotal_global, int64_t* labels, int32_t* hamming,
                              double* score_binary, double* score_cosine, int32_t* out_count);
int vrq_search3_sharded_group(int world, vrq_index* const* ixs, int64_t nq, const float* const* q_float, const uint8_t* const* q_ubin, int k,
                              int binary_oversample, int int8_oversample, const int64_t* pos_base, int64_t ntotal_global,
                              int64_t* const* labels, int32_t* const* hamming, double* const* score_binary, double* const* score_cosine,
                              int32_t* const* out_count);

/* ---------------------------------------------------------------- stand-alone rescoring kernels --- */
/* Phase II (CohereEnhancedVectorDB.py:283-293): score[q,i] = sum_j qf[q,j] * (2*bit_j(codes[pos[q,i]]) - 1), f64. */
int vrq_rescore_binary(vrq_ctx*, const uint8_t* codes, int64_t n, int d, const int64_t* pos, int64_t nq, int m,
                       const float* q_float, double* score);
/* Phase III (CohereEnhancedVectorDB.py:302-318): score = dot(qf, f(int8 row)) / ||row||_2 ; -inf if the norm is 0. */
int vrq_rescore_int8cos(vrq_ctx*, const int8_t* rows, int64_t n, int d, const int64_t* pos, int64_t nq, int m,
                        const float* q_float, double* score);

/* ---------------------------------------------------------------- synthetic data (no network) ----- */
/* Counter-based generator shared bit-for-bit with the oracle: float rows, and the
 * Cohere-like (ubinary = x > 0, int8 = clip(rint(1259 x - 0.69))) pair derived from the same rows. */
int vrq_synth_f32(vrq_ctx*, uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale, float* out);
int vrq_synth_codes_int8(vrq_ctx*, uint64_t seed, int64_t row0, int64_t nrows, int d, uint8_t* codes, int8_t* int8_rows);
/* Append nrows synthetic rows straight into a device-resident index (ids = id0 + i), payload INT8_RAW or NONE. */
int vrq_index_add_synthetic(vrq_index*, uint64_t seed, int64_t row0, int64_t nrows, int64_t id0);

/* Benchmark-only stand-in for an INT8_RAW payload that does not fit in HBM (1 billion rows x 1 KB on fewer than 8 GPUs,
 * SURVEY.md H6): nothing is stored, Phase III regenerates the int8 row of position p from the counter-based generator
 * (seed, row0 + p) - the bytes vrq_index_add_synthetic would have stored.  Only on an empty index whose payload kind is
 * INT8_RAW; afterwards the index accepts vrq_index_add_synthetic(seed, row0 + ntotal, ...) only, and no removals. */
int vrq_index_set_synthetic_payload(vrq_index*, uint64_t seed, int64_t row0);

/* CUDA-event timing of the library's own kernels on the launching stream (bench.py's roofline.achieved).
 * After vrq_ctx_enable_timing(ctx, 1) every scan / encode / rescore / merge launch group is bracketed by events;
 * vrq_ctx_timing_ms returns the summed device milliseconds of category `which` ("scan" | "encode" | "rescore" |
 * "merge") recorded since the previous query, and how many regions that was. */
int vrq_ctx_enable_timing(vrq_ctx*, int on);
double vrq_ctx_timing_ms(vrq_ctx*, const char* which, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* VRQ_H */
