/* tfsc_b200.h -- C ABI of libtfsc_b200.so, the B200-native replacement for the
 * route -> ensure-resident -> predict path of mKaloer/TFServingCache.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no
 * C++/torch types, callable from cgo, ctypes or JNI, from any thread, re-entrant.  Every entry
 * point cites the reference interface (file:line under the reference repo) it replaces.
 *
 * Conventions
 *   - return value: >= 0 success (meaning documented per call), < 0 one of TFSC_E_*.
 *   - tfsc_last_error() returns a thread-local message for the last failing call.
 *   - strings are NUL-terminated UTF-8; out buffers are caller-owned with an explicit capacity;
 *     a too-small buffer yields TFSC_E_BUFFER.
 *   - buffers returned through `void**` are library-owned and released with tfsc_free().
 *   - there is NO CPU fallback: every compute entry fails with TFSC_E_NO_DEVICE when no
 *     sm_100-class device is usable.
 */
#ifndef TFSC_B200_H_
#define TFSC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFSC_ABI_VERSION 2

/* error codes; the gRPC status each one maps to is given in parentheses */
#define TFSC_OK 0
#define TFSC_E_INVALID (-3)       /* INVALID_ARGUMENT (3)  */
#define TFSC_E_TIMEOUT (-4)       /* DEADLINE_EXCEEDED (4): "Timeout: Model did not load in time" */
#define TFSC_E_NOT_FOUND (-5)     /* NOT_FOUND (5)         */
#define TFSC_E_EXHAUSTED (-8)     /* RESOURCE_EXHAUSTED (8)*/
#define TFSC_E_UNIMPLEMENTED (-12)/* UNIMPLEMENTED (12): MultiInference, tfservingproxy.go:215-217 */
#define TFSC_E_INTERNAL (-13)     /* INTERNAL (13)         */
#define TFSC_E_NO_DEVICE (-14)    /* UNAVAILABLE (14): CUDA device / extension missing */
#define TFSC_E_EMPTY_RING (-20)   /* consistent.ErrEmptyCircle, cluster.go:118-120 */
#define TFSC_E_BUFFER (-21)       /* caller buffer too small */

/* ModelVersionStatus_State, pkg/cachemanager/servingcontroller.go:29-54 */
#define TFSC_STATE_UNKNOWN 0
#define TFSC_STATE_START 10
#define TFSC_STATE_LOADING 20
#define TFSC_STATE_AVAILABLE 30
#define TFSC_STATE_UNLOADING 40
#define TFSC_STATE_END 50

/* fetchModel outcome, pkg/cachemanager/cachemanager.go:103-150 */
#define TFSC_FETCH_HIT 0     /* cached and resident: cache_hits_total++            */
#define TFSC_FETCH_RELOAD 1  /* in the host tier but not HBM-resident (:133-143)    */
#define TFSC_FETCH_MISS 2    /* not cached: provider load, cache_misses_total++     */

/* tensorflow.DataType subset, proto/tensorflow/core/framework/types.pb.go:30-55 */
#define TFSC_DT_FLOAT 1
#define TFSC_DT_INT32 3
#define TFSC_DT_INT64 9

int tfsc_abi_version(void);
const char* tfsc_last_error(void);
const char* tfsc_strerror(int code);
void tfsc_free(void* p);

/* ---------------------------------------------------------------- a3-a5: routing ring ------
 * Bit-exact restatement of stathat.com/c/consistent v1.0.0 as driven by
 * pkg/taskhandler/cluster.go:55 (New), :111 (Set), :117 (GetN). */
typedef struct tfsc_ring tfsc_ring;

uint32_t tfsc_crc32_ieee(const void* data, size_t len); /* Go crc32.ChecksumIEEE */
tfsc_ring* tfsc_ring_new(void);                         /* consistent.New(), cluster.go:55 */
void tfsc_ring_free(tfsc_ring* r);
/* consistent.Set(members): cluster.go:104-113 (clusterUpdated). member = "host:rest:grpc". */
int tfsc_ring_set(tfsc_ring* r, const char* const* members, int n_members);
int tfsc_ring_members(const tfsc_ring* r);              /* member count */
int tfsc_ring_points(const tfsc_ring* r);               /* ring points (<= 20 * members) */
/* consistent.GetN(key, n): cluster.go:116-130 (FindNodeForKey). Writes up to n member strings,
 * '\n'-separated, clockwise order, into buf. Returns the number of members written. */
int tfsc_ring_getn(const tfsc_ring* r, const char* key, int n, char* buf, size_t cap);
/* Replica choice among the GetN candidates: "random" = the reference (taskhandler.go:91), "first" = primary,
 * "hot-spread" = primary unless the key's recent request share exceeds hot_fraction / members (then random),
 * "balanced" = hot-spread + least-loaded-replica binding (tfsc_picker_pick_ids), "hash" = crc32(key) mod replicas
 * (stateless: independent processes agree on the replica of a key).
 * Deterministic for a given seed and call sequence. tfsc_picker_pick returns an index in [0, n_replicas). */
typedef struct tfsc_picker tfsc_picker;
tfsc_picker* tfsc_picker_new(const char* policy, uint64_t seed, double hot_fraction);
void tfsc_picker_free(tfsc_picker* p);
int tfsc_picker_pick(tfsc_picker* p, const char* key, int n_replicas, int members);
/* same with a stable integer id per candidate; policy "balanced" = hot-spread + sticky power-of-two-choices
 * (a cold key binds to the candidate that holds the fewest keys) */
int tfsc_picker_pick_ids(tfsc_picker* p, const char* key, const int* member_ids, int n_replicas, int members);
/* key = modelName + "##" + version (taskhandler.go:85). Returns strlen. */
int tfsc_model_key(const char* model_name, const char* version, char* buf, size_t cap);

/* ---------------------------------------------------------------- a9: LRU model cache ------
 * pkg/cachemanager/lrucache.go:20-105 (ModelCache interface :11-18). Not internally
 * synchronised, exactly like the reference (the cache manager holds the lock). */
typedef struct tfsc_lru tfsc_lru;

tfsc_lru* tfsc_lru_new(const char* base_dir, int64_t capacity_bytes); /* NewLRUCache :28 */
void tfsc_lru_free(tfsc_lru* c);
/* Put :54-65. Returns the number of entries evicted to make room. */
int tfsc_lru_put(tfsc_lru* c, const char* model_name, int64_t version, const char* path, int64_t size_on_disk);
/* Get :43-51 (touches recency). Returns 1 if present (fills size/path), 0 if not. */
int tfsc_lru_get(tfsc_lru* c, const char* model_name, int64_t version, int64_t* size_on_disk, char* path, size_t cap);
int tfsc_lru_ensure_free_bytes(tfsc_lru* c, int64_t bytes);           /* :68-87; returns #evicted */
int64_t tfsc_lru_current_size(const tfsc_lru* c);
int64_t tfsc_lru_capacity(const tfsc_lru* c);
int tfsc_lru_len(const tfsc_lru* c);
/* ListModels :89-97, MRU -> LRU, one "name\tversion\tsize\tpath\n" line per model. Returns count. */
int tfsc_lru_list(const tfsc_lru* c, char* buf, size_t cap);

/* ---------------------------------------------------------------- a1/a2/a7: request parsing -
 * tfServingRestURLMatch (tfservingproxy.go:24) + RestProxy.Serve status logic (:93-129).
 * Returns 200 (name+version filled, version verbatim incl. leading zeros), 404 or 400. */
int tfsc_rest_match_url(const char* url, char* model_name, size_t name_cap, char* version, size_t version_cap);
/* exact JSON error body json.NewEncoder would emit for 404 / 400 (tfservingproxy.go:99-124) */
const char* tfsc_rest_error_body(int http_status);
/* strconv.ParseInt(version, 10, 64), cachemanager.go:297. 0 or TFSC_E_INVALID. */
int tfsc_parse_version(const char* version, int64_t* out);
/* clientForSpec (tfservingproxy.go:246-250): scan a serialized ModelSpec-bearing request
 * (PredictRequest/ClassificationRequest/...: model_spec is field 1) and return name + version
 * string ("0" when absent). */
int tfsc_grpc_model_spec(const void* req, size_t len, char* model_name, size_t name_cap, char* version, size_t version_cap);

/* ---------------------------------------------------------------- a11: disk model provider --
 * diskmodelprovider.go:46-69 findSrcPathForModel (numeric version match), :71-83 ModelSize
 * (fixed: recursive byte size). */
int tfsc_disk_find_version_dir(const char* base_dir, const char* model_name, int64_t version, char* buf, size_t cap);
int64_t tfsc_disk_model_size(const char* base_dir, const char* model_name, int64_t version);
/* TensorFlow SavedModel ingestion without TensorFlow. The disk provider imports a version directory that
 * holds `saved_model.pb` + `variables/variables.{index,data-*}` (and no tfsc_model.json) on the fly when the model is
 * fetched; this entry does the same conversion offline and writes tfsc_model.json + weights.bin into out_dir (may equal
 * version_dir). Recognised graphs: y = a*x + b (half_plus_two) and MatMul + BiasAdd (+ Relu) chains; anything else is
 * TFSC_E_INVALID with the offending node in tfsc_last_error(). Needs no GPU. */
int tfsc_savedmodel_convert(const char* version_dir, const char* out_dir);
uint32_t tfsc_crc32c(const void* data, size_t len);   /* CRC-32C (Castagnoli), the tensor-bundle / table checksum */

/* ---------------------------------------------------------------- server (a6,a8,a10,X) ------
 * One server = the cache tier + proxy tier of cmd/taskhandler/main.go:45-113 for the GPUs of
 * this process: one "node" per GPU (ring member), each with its own LRU host tier, HBM arena,
 * residency table, copy stream and batcher. Configuration is a flat JSON object whose keys are
 * the reference's viper keys, e.g.
 *   {"modelProvider.type":"diskProvider","modelProvider.diskProvider.baseDir":"/model_repo",
 *    "modelCache.size":68719476736,"serving.maxConcurrentModels":32,"proxy.replicasPerModel":2,
 *    "gpu.devices":[0,1],"gpu.arenaBytes":171798691840,"gpu.maxBatch":8,"gpu.members":["gpu0:0:0","gpu1:0:0"]}
 */
typedef struct tfsc_server tfsc_server;

typedef struct tfsc_tensor {
  const char* name;   /* signature key ("x", "y", ...); may be NULL for the only input/output */
  int32_t dtype;      /* TFSC_DT_* */
  int32_t rank;
  int64_t shape[8];
  void* data;         /* host pointer (tfsc_predict) or device pointer (tfsc_predict_device) */
  size_t nbytes;
} tfsc_tensor;

typedef struct tfsc_stats {
  /* names kept from cachemanager.go:24-43 / tfservingproxy.go:25-32 */
  int64_t cache_total, cache_hits_total, cache_misses_total;
  int64_t proxy_requests_rest, proxy_requests_grpc, proxy_failures_rest, proxy_failures_grpc;
  int64_t evictions_host, evictions_hbm;
  int64_t h2d_weight_bytes, h2d_input_bytes, d2h_output_bytes;
  int64_t kernel_launches, batches, batched_rows;
  int64_t arena_bytes_used, arena_bytes_capacity, resident_models, host_models;
  double cache_duration_seconds_sum, cache_fetch_duration_seconds_sum;
  /* forward hop (a6): requests sent to / received from other ranks, bytes the owner moved over NVLink */
  int64_t fwd_out_requests, fwd_in_requests, fwd_out_failures, fwd_peer_bytes_read, fwd_peer_bytes_written;
  double fwd_rtt_seconds_sum;
  /* HBM arena defragmentation: passes that moved resident blocks (device-to-device) and the bytes they moved */
  int64_t arena_compactions, arena_compacted_bytes;
} tfsc_stats;

tfsc_server* tfsc_server_create(const char* config_json); /* main.go:45-113 */
void tfsc_server_destroy(tfsc_server* s);
int tfsc_server_num_nodes(const tfsc_server* s);
/* DiscoveryService member update (cluster.go:25-30,104-113). Default members: "gpu<i>:0:0". */
int tfsc_server_set_members(tfsc_server* s, const char* const* members, int n);
/* nodeForKey (taskhandler.go:84-92): ring lookup + replica pick. Writes the ordered replica
 * set as local node indices (or -1 for non-local members) into nodes[0..cap) and returns the
 * picked index into that list via *picked. Return value = replica count. */
int tfsc_route(tfsc_server* s, const char* model_name, const char* version, int* nodes, int cap, int* picked);
/* fetchModel (cachemanager.go:91-152) on one node: ensure the model is HBM-resident.
 * Returns TFSC_FETCH_* or an error. Blocks only this caller (per-model load lock). */
int tfsc_model_ensure(tfsc_server* s, int node, const char* model_name, int64_t version);
/* Same decision and bookkeeping as tfsc_model_ensure, but returns as soon as the page-in is queued on the copy
 * stream (state LOADING); launches on the model wait for it on-device (event), the host never blocks. */
int tfsc_model_ensure_async(tfsc_server* s, int node, const char* model_name, int64_t version);
/* GetModelStatus (servingcontroller.go:114-138): TFSC_STATE_* or TFSC_E_NOT_FOUND. */
int tfsc_model_status(tfsc_server* s, int node, const char* model_name, int64_t version);
/* Resident set, MRU first: "name\tversion\tbytes\tstate\n" lines. Returns count. */
int tfsc_resident_list(tfsc_server* s, int node, char* buf, size_t cap);
/* LRU host tier listing (LocalCache.ListModels), same line format as tfsc_lru_list. */
int tfsc_host_list(tfsc_server* s, int node, char* buf, size_t cap);

/* proxyServiceServer.Predict (tfservingproxy.go:201-212) with the forward replaced by on-GPU
 * execution: route -> ensure-resident -> batch -> kernels. Host tensors in, host tensors out;
 * out[i].data/nbytes must be a caller buffer large enough for the result (shape is filled).
 * `version` is the verbatim string ("00000123" routes differently from "123": reference quirk). */
int tfsc_predict(tfsc_server* s, const char* model_name, const char* version,
                 const tfsc_tensor* in, int n_in, tfsc_tensor* out, int n_out);
/* tfsc_predict with a deadline (absolute, on the clock of tfsc_now_ns() = CLOCK_MONOTONIC; 0 = none): a request still
 * queued when its deadline passes is answered TFSC_E_TIMEOUT without being launched (grpc deadline / proxy.grpcTimeout). */
int tfsc_predict_deadline(tfsc_server* s, const char* model_name, const char* version,
                          const tfsc_tensor* in, int n_in, tfsc_tensor* out, int n_out, int64_t deadline_ns);
int64_t tfsc_now_ns(void);
/* The cache tier of one member, without the ring lookup (the reference's second tier: cachemanager.ServeRest / ServeGrpc on
 * cacheRestPort / cacheGrpcPort, cmd/taskhandler/main.go:60-84 -- a request that reaches a cache node is served there):
 * `member` indexes the current member list ("gpu.members" / tfsc_server_set_members order). A local member runs on its node;
 * a member of another rank takes the forward hop. For callers that route themselves (tfsc_route, or a front load balancer). */
int tfsc_predict_member(tfsc_server* s, int member, const char* model_name, const char* version,
                        const tfsc_tensor* in, int n_in, tfsc_tensor* out, int n_out, int64_t deadline_ns);
/* Asynchronous Predict: no OS thread is parked per in-flight request (a Go handler keeps a goroutine, not an M).
 *   submit: route -> ensure-resident (may block on a cold load of THIS model only) -> signature checks -> the input rows are
 *           copied to pinned staging, so `in` may be reused at once. out[0].data / nbytes is the caller's result buffer.
 *   wait  : timeout_ns < 0 blocks; >= 0 waits at most that long and answers TFSC_E_TIMEOUT while the request is still in
 *           flight (the ticket stays valid). On success out[0] holds dtype / shape / nbytes and the data.
 *   release: frees the ticket (waits for the request to retire first if it is still in flight). */
typedef struct tfsc_ticket tfsc_ticket;
int tfsc_predict_submit(tfsc_server* s, const char* model_name, const char* version, const tfsc_tensor* in, int n_in,
                        tfsc_tensor* out, int n_out, int64_t deadline_ns, tfsc_ticket** ticket);
int tfsc_predict_wait(tfsc_ticket* ticket, int64_t timeout_ns);
void tfsc_predict_release(tfsc_ticket* ticket);
/* Same, wire level: serialized tensorflow.serving.PredictRequest in, PredictResponse out
 * (library-owned; tfsc_free). This is what a cgo Predict handler calls. */
int tfsc_grpc_predict(tfsc_server* s, const void* req, size_t req_len, void** resp, size_t* resp_len);
/* proxyServiceServer.Classify / Regress (tfservingproxy.go:173-198) and SessionRun (:233-244), wire level like
 * tfsc_grpc_predict: serialized ClassificationRequest / RegressionRequest / SessionRunRequest in, the matching response out
 * (library-owned; tfsc_free). tf.Example inputs are mapped onto the model's input rows through the classify / regress
 * signatures the bundle declares ("extra_signatures" of the manifest; a SavedModel import carries over the signatures of
 * saved_model.pb, e.g. half_plus_two's regress_x_to_y / classify_x_to_y on feature "x"). A model without such a signature
 * answers INVALID_ARGUMENT with TF-Serving's message. MultiInference stays an error (tfservingproxy.go:215-217). */
int tfsc_grpc_classify(tfsc_server* s, const void* req, size_t req_len, void** resp, size_t* resp_len);
int tfsc_grpc_regress(tfsc_server* s, const void* req, size_t req_len, void** resp, size_t* resp_len);
int tfsc_grpc_session_run(tfsc_server* s, const void* req, size_t req_len, void** resp, size_t* resp_len);
/* RestProxy.Serve (tfservingproxy.go:93-129) with on-GPU execution: GET status / POST :predict.
 * Returns 0 and fills *http_status + body (library-owned; tfsc_free). */
int tfsc_rest_handle(tfsc_server* s, const char* method, const char* url, const void* body, size_t body_len,
                     int* http_status, void** resp, size_t* resp_len);

/* Device-resident predict on an explicit node/stream: x and y are DEVICE pointers (possibly
 * peer memory of another GPU: the forward hop a6 becomes NVLink loads/stores inside the first /
 * last kernel). rows = batch rows. stream = cudaStream_t or NULL for the node's compute stream.
 * The model must have been made resident (tfsc_model_ensure); it is pinned for the launch. */
int tfsc_predict_device(tfsc_server* s, int node, const char* model_name, int64_t version,
                        const void* x, int64_t rows, void* y, void* stream);
int tfsc_node_sync(tfsc_server* s, int node);
/* serving.maxConcurrentModels of one node at run time (takes effect at the next reload: cachemanager.go:167-170) */
int tfsc_node_set_max_resident(tfsc_server* s, int node, int max_concurrent_models);

/* ---------------------------------------------------------------- a6 / X7: forward hop between processes ------
 * One process per GPU (torchrun): config keys "cluster.rank", "cluster.endpoints" (one unix-socket path per entry of
 * "gpu.members", same order), "cluster.slotBytes", "cluster.windowSlots". A Predict whose ring owner is another rank is
 * forwarded there like restDirector / grpcDirector do (taskhandler.go:95-147), except that only a ~100-byte control message
 * crosses the socket: the request rows sit in this rank's FORWARD WINDOW (HBM exported with CUDA IPC), the owner's gather /
 * scatter kernels (or its first / last layer, with tfsc_predict_device) read x and write y there over NVLink.
 * tfsc_fwd_window: this rank's window (device pointer, bytes, slot size); returns the rank.
 * tfsc_fwd_peer_window: rank `peer_rank`'s window mapped into this process (dials the peer on first use). */
int tfsc_fwd_window(tfsc_server* s, void** dev_ptr, size_t* bytes, size_t* slot_bytes);
int tfsc_fwd_peer_window(tfsc_server* s, int peer_rank, void** dev_ptr, size_t* bytes);
/* synchronous copy between any two addresses of the unified address space (host, this GPU, a mapped peer window): how a
 * host program without its own CUDA binding fills / reads window slots for tfsc_predict_device */
int tfsc_device_memcpy(void* dst, const void* src, size_t nbytes);
int tfsc_get_stats(tfsc_server* s, int node, tfsc_stats* out); /* node = -1: sum over nodes */
/* number of kernels launched by this library since load (bench gpu_launches) */
int64_t tfsc_kernel_launches(void);

/* ---------------------------------------------------------------- raw kernels (X rows) ------
 * Direct launches on caller-provided device memory for parity tests and roofline timing.
 * y[rows,n] = act(x[rows,k] W[k,n] + b[n]); fp32; W row-major [k,n] (TF dense kernel layout). */
int tfsc_k_affine(const float* x, float* y, int64_t n, const float* a, const float* b, void* stream);      /* X1 */
int tfsc_k_dense(const float* x, const float* w, const float* b, float* y, int rows, int k, int n, int relu,
                 float* workspace, size_t workspace_bytes, void* stream);                                    /* X2 */
size_t tfsc_k_dense_workspace(int rows, int k, int n);
/* tfsc_k_dense with an explicit kernel choice: 0 auto (<= 8 rows: cluster-pair kernel with programmatic dependent launch,
 * csrc/dense_cluster.cu -- two CTAs split K and meet in distributed shared memory; more rows: tensor cores), 1 LDG-stream
 * SIMT kernel with split-K workspace (the round-1 default; still the fallback for shapes the cluster kernel does not take),
 * 2 / 4 bulk-copy (TMA) ring with 8 / 4 k-lanes, 3 tensor cores for every row count, 5 cluster-pair kernel for <= 8 rows and
 * the SIMT fallback otherwise. Same arguments, workspace and results (the fp32 summation order differs between variants; each
 * variant is bit-reproducible). */
int tfsc_k_dense_variant(int variant, const float* x, const float* w, const float* b, float* y, int rows, int k, int n,
                         int relu, float* workspace, size_t workspace_bytes, void* stream);
/* X3: the tcgen05/TMEM (3xTF32) path alone, rows <= 64, n % 32 == 0, k % 4 == 0. tfsc_k_dense picks it
 * automatically for more than 8 rows; this entry exists for parity tests and roofline timing. */
int tfsc_k_dense_tc(const float* x, const float* w, const float* b, float* y, int rows, int k, int n, int relu,
                    float* workspace, size_t workspace_bytes, void* stream);

/* X6 (+ X7): batch gather / scatter as one kernel over a table of segments. src / dst may be pinned host memory, local HBM or
 * a peer's forward window (NVLink): this is the kernel the batcher uses to assemble a batch from its requests' rows and to
 * hand the result rows back (csrc/nn_kernels.cu copy_segments_kernel). */
typedef struct tfsc_copy_seg {
  const void* src;
  void* dst;
  uint64_t bytes;
} tfsc_copy_seg;
int tfsc_k_copy_segments(const tfsc_copy_seg* segs, int n, void* stream);

/* X4/X5 building blocks of the graph executor (conv nets), fp32, row-major / NHWC. act: 0 none, 1 relu, 2 gelu.
 * C[M,N] = act(A[M,K] (row stride lda) * B[K,N] + bias[N] (+ R[M,N])); bias / R may be NULL. */
int tfsc_k_gemm(const float* a, const float* b, const float* bias, const float* r, float* c, int m, int n, int k, int lda,
                int act, void* stream);
/* the same GEMM on tcgen05 / TMEM (3xTF32, fp32-accurate): m >= 64, n >= 64, n % 32 == 0, k >= 32, lda % 4 == 0 */
int tfsc_k_gemm_tc(const float* a, const float* b, const float* bias, const float* r, float* c, int m, int n, int k, int lda,
                   int act, void* stream);
/* X4: implicit-GEMM convolution on tcgen05 / TMEM (3xTF32): y[B,OH,OW,cout] = act(conv2d(x[B,H,W,C] NHWC, w[KH,KW,C,cout] HWIO)
 * + bias (+ r)); the patch tiles are gathered from x by TMA im2col tensor maps, no patch matrix is materialised.
 * c % 32 == 0, cout >= 64 and % 32 == 0, batch*OH*OW >= 64. act: 0 none, 1 relu, 2 gelu, 3 tanh. */
int tfsc_k_conv_tc(const float* x, const float* w, const float* bias, const float* r, float* y, int batch, int h, int wd, int c,
                   int kh, int kw, int stride, int pad, int cout, int act, void* stream);
/* debugging aid: with TFSC_GT_TRACE=1 in the environment, 16 clock64 stamps of CTA 0 of the most recent persistent tcgen05
 * GEMM launch (entry, setup done, first TMA, first tile landed, converted, first MMA, per-tile commit / epilogue, exit) */
int tfsc_debug_gemm_trace(long long* out16);
/* col[(b*OH+oh)*OW+ow][(kh*KW+kw)*C+c] patch matrix with row stride ldc >= KH*KW*C (zero padded) */
int tfsc_k_im2col(const float* x, float* col, int batch, int h, int w, int c, int kh, int kw, int stride, int pad, int ldc,
                  void* stream);
int tfsc_k_maxpool(const float* x, float* y, int batch, int h, int w, int c, int kh, int kw, int stride, int pad, void* stream);
int tfsc_k_avgpool(const float* x, float* y, int batch, int hw, int c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TFSC_B200_H_ */
