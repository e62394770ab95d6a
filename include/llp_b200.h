/*
 * llp_b200.h — C-ABI of the B200-native LLP hot path (libllp_b200.so).
 *
 * The reference (snap-research/linkless-link-prediction) is pure Python with NO FFI /
 * plugin layer of its own: its hot-path arithmetic lives in the CUDA
 * kernels of third-party wheels reached through Python call sites.  Each entry point below
 * therefore cites the reference CALL SITE (file:line under /root/reference) whose
 * third-party kernel it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * would add on the reference side.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory unless named host_*.
 *   - the caller owns all memory including workspaces (sizes via *_workspace_bytes);
 *     kernels never allocate or free.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - stateless and re-entrant; one process per GPU.
 *   - return 0 on success, >0 = cudaError_t, <0 = LLP_E_* below.  There is NO CPU
 *     fallback: a non-sm_100 device yields LLP_E_DEVICE.
 *   - row-major matrices with explicit leading dimensions (ld*, in elements).
 */
#ifndef LLP_B200_H_
#define LLP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define LLP_F32 0
#define LLP_BF16 1

#define LLP_E_BADARG (-1)    /* null pointer / negative size / unknown enum            */
#define LLP_E_ALIGN (-2)     /* pointer or leading dimension not aligned as documented */
#define LLP_E_WORKSPACE (-3) /* workspace too small                                    */
#define LLP_E_DEVICE (-4)    /* current device is not sm_100                           */
#define LLP_E_SHAPE (-5)     /* shape not supported by the selected backend            */

#define LLP_GEMM_AUTO 0    /* bf16 -> tcgen05 kind::f16, f32 -> tcgen05 3xTF32 (SIMT only when TMA cannot address an operand) */
#define LLP_GEMM_SIMT 1    /* CUDA-core FFMA, fp32 accumulate                            */
#define LLP_GEMM_TCGEN05 2 /* tcgen05.mma + TMA + TMEM (bf16 operands, fp32 accumulate)  */
#define LLP_GEMM_TF32X3 3  /* tcgen05.mma kind::tf32, 3-term split accumulation of fp32 operands (fp32-grade results:
                              the reference's GEMMs are true fp32, src/models.py:48,143,146) */

int llp_version(void);
const char* llp_error_string(int code);
/* 1 if the CURRENT device is compute capability 10.x, 0 if not, <0/>0 on error. */
int llp_device_supported(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t llp_launch_count(void);
/* development knob for benchmark sweeps: selects between kernel variants that compute the SAME result (key 0 = SpMM
 * occupancy/register variant, 10/12/14/16-20 = GEMM pipeline variants, 15 = instrumentation).  The work-skipping
 * experiment keys (1, 2, 11, 13) are compiled only into -DLLP_EXPERIMENT builds; the shipped library ignores them. */
void llp_set_tuning(int key, int value);
/* development aid: copy n (<= 4096) int64 of the instrumentation scratch to the host (synchronises the device) */
int llp_debug_read(int64_t* host_out, int n);

/* ---------------------------------------------------------------------------------------
 * Graph structure.  Replaces: PyG MessagePassing.__collect__/aggregate over a dense [2,E]
 * edge_index (models.py:113,118 -> SAGEConv.propagate; train_teacher_gnn.py:317,331) and
 * torch_sparse's CSR conversion (sageconv_updated.py:86-89).
 * Messages flow key-side <- value-side: pass (edge_val=src, edge_key=dst) for the forward
 * CSR (rows = destinations) and swap them for the transpose.  Stable in the original edge
 * order, so results are bit-reproducible.  int32 indices: N, E < 2^31.
 * ------------------------------------------------------------------------------------- */
size_t llp_csr_build_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int llp_csr_build(const int64_t* edge_val, const int64_t* edge_key, int64_t num_edges, int64_t num_nodes,
                  int32_t* rowptr /*[N+1]*/, int32_t* col /*[E]*/, int32_t* perm /*[E] or NULL*/,
                  float* inv_deg /*[N] 1/max(deg,1) or NULL*/, void* workspace, size_t workspace_bytes,
                  void* stream);

/* Edge-balanced work plan for llp_spmm: chunk c owns rows [first_row[c], first_row[c+1]).  The plan buffer holds
 * llp_spmm_plan_ints(num_edges) int32 (32-byte aligned): the first-row table [num_chunks+1], then one 32-byte
 * descriptor per chunk (its edge range and first rows, so the kernel starts without chasing the tables). */
int64_t llp_spmm_num_chunks(int64_t num_edges);
int64_t llp_spmm_plan_ints(int64_t num_edges);
int llp_spmm_plan(const int32_t* rowptr, int64_t num_nodes, int64_t num_edges,
                  int32_t* chunk_first_row /*[llp_spmm_plan_ints(num_edges)]*/,
                  int32_t* hub_list /*[4*(num_chunks+1)] out, 16-byte aligned: hub table = header {n_small, n_big, index of
                                       the first big record, -} + one record {chunk, row, begin, end} per split row; small
                                       hubs fill records 1.., big ones the last records downwards.  A compacted copy
                                       (header, small records, big records; header[2] = n_small + n_big) works as well */,
                  int32_t* num_hubs /*[1] device out: rows longer than the split threshold (small + big)*/, void* stream);
size_t llp_spmm_workspace_bytes(int64_t num_edges, int64_t feat);

/* CSR gather-reduce SpMM: out[r,:] = (mean ? 1/max(deg r,1) : 1) * sum_{e in row r} scale[col e] * x[col e,:]
 * Replaces torch_scatter.scatter(reduce='mean') + index_select (SAGEConv.propagate, models.py:113) and
 * its autograd transpose (index_add_).  dtype in {LLP_F32, LLP_BF16}; fp32 accumulation in CSR order.
 * x/out rows must be 4-byte aligned at least; 16-byte aligned rows (ld*elt % 16 == 0) take the
 * 128-bit path.  src_scale may be NULL. */
int llp_spmm(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* chunk_first_row,
             int64_t num_rows, int64_t num_edges, const void* x, int64_t ldx, int64_t feat,
             const float* src_scale, int mean, void* out, int64_t ldo, void* workspace,
             const int32_t* hub_list /* the hub table of llp_spmm_plan */, int64_t num_hubs /* host copy of *num_hubs */,
             void* stream);

/* ---------------------------------------------------------------------------------------
 * Node-partitioned encoder over peer memory (the reference is single-GPU, its aggregation call
 * sites are models.py:113,118).  One process per GPU: every rank exports its activation block (llp_ipc_export), maps the
 * peers' blocks (llp_ipc_open) and aggregates its own rows with llp_spmm_peer, whose gathers of remote rows are plain
 * loads over NVLink / NVSwitch - no staged all-gather.  llp_peer_barrier orders "blocks in place" / "blocks read".
 * ------------------------------------------------------------------------------------- */
/* IPC handle (64 bytes) of the device allocation that contains `ptr`, and ptr's offset inside it. */
int llp_ipc_export(const void* ptr, void* handle64, int64_t* offset);
/* Map a peer's allocation; *ptr_out corresponds to the exporter's `ptr`, *base_out is what llp_ipc_close takes. */
int llp_ipc_open(const void* handle64, int64_t offset, void** ptr_out, void** base_out);
int llp_ipc_close(void* base);
/* Barrier of `world` ranks on `stream` (CUDA-graph capturable): flags[r] = rank r's zero-initialised uint64[world + 2]
 * flag array as mapped into this process (HOST array of pointers).  Bounded wait (~2 s): on a timeout slot world + 1 of
 * this rank's array becomes non-zero and the kernel returns. */
int llp_peer_barrier(void* const* flags, int rank, int world, void* stream);
/* dst[dst_rows ? dst_rows[i] : i, :] = row (src[i] & ((1 << shift) - 1)) of rank (src[i] >> shift)'s block, i < n_rows
 * (row_bytes % 16 == 0, <= 4096): every remote row a rank's local messages reference, fetched once over NVLink into a
 * local staging matrix that an ordinary llp_spmm then reads; with dst_rows, the embedding rows an edge batch scores,
 * dropped into their places of an otherwise untouched [N, feat] matrix.  peer_x = DEVICE table of block base pointers. */
int llp_peer_gather_rows(const void* const* peer_x, const int32_t* src, const int32_t* dst_rows /* or NULL */, int shift,
                         int64_t n_rows, int64_t row_bytes, void* dst, void* stream);
/* Sparse return of the embedding gradient (transpose of the pull by node id): every rank publishes {count, ids...} of the
 * nodes its edge shard scored and its gradient rows for them inside an exported [N_padded, ld] matrix; the owner of a node
 * block marks which of its rows each rank touched (llp_peer_mark_rows; `mark` = zeroed uint8[world * n_loc]) and pulls +
 * adds the marked rows in rank order with fp32 accumulation (llp_peer_reduce_rows; rows nobody marked become zero).
 * Replaces a dense reduce-scatter of all N rows (the reference is single-GPU: autograd's index_put_ of models.py:140). */
int llp_peer_mark_rows(const void* const* ids_table /* DEVICE [world] */, int world, int64_t lo, int64_t n_loc, int64_t max_ids,
                       void* mark, void* stream);
int llp_peer_reduce_rows(int dtype, const void* const* g_table /* DEVICE [world] */, int world, const void* mark, int64_t lo,
                         int64_t n_loc, int64_t feat, int64_t ld, void* out, int64_t ldo, void* stream);
/* llp_spmm with the source rows in peer-mapped blocks: peer_x = DEVICE array of `world` base pointers (rank r's
 * [peer_nloc, ldx] block), col[e] = (owner rank << peer_shift) | row inside the owner's block, src_scale (optional)
 * indexed by owner * peer_nloc + row.  Row widths of 256 or 512 bytes; LLP_E_SHAPE otherwise (use llp_spmm on an
 * all-gathered matrix then). */
int llp_spmm_peer(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* chunk_first_row,
                  int64_t num_rows, int64_t num_edges, const void* const* peer_x, int world, int peer_shift,
                  int64_t peer_nloc, int64_t ldx, int64_t feat, const float* src_scale, int mean, void* out, int64_t ldo,
                  void* workspace, const int32_t* hub_list, int64_t num_hubs, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense layers.  Replaces cuBLAS SGEMM behind F.linear (sageconv_updated.py:71,76; PyG
 * SAGEConv lin_l/lin_r; models.py:48,143,146) plus the relu/dropout/bias ATen kernels
 * (models.py:116-117,144-145).
 * ------------------------------------------------------------------------------------- */
typedef struct llp_gemm_nt_args {
  int dtype;     /* operand dtype of A*, B* (LLP_F32 | LLP_BF16)                              */
  int out_dtype; /* dtype of D and addend                                                     */
  int backend;   /* LLP_GEMM_*                                                                */
  int relu;      /* apply max(.,0) after bias/addend                                          */
  int64_t M, N, K1, K2;
  const void* A1; int64_t lda1; /* [M,K1]                                                     */
  const void* B1; int64_t ldb1; /* [N,K1]   D = A1 * B1^T                                     */
  const void* A2; int64_t lda2; /* [M,K2] or NULL                                             */
  const void* B2; int64_t ldb2; /* [N,K2]     + A2 * B2^T                                     */
  const float* bias;            /* [N] or NULL                                                */
  const void* addend; int64_t ldadd; /* [M,N] out_dtype or NULL: + addend                     */
  const void* gate; int64_t ldgate;  /* [M,N] out_dtype or NULL: D = gate>0 ? D*gate_scale : 0
                                        (backward of relu+dropout from the saved output)      */
  float gate_scale;
  float dropout_p;              /* 0 = off; keep-prob 1-p, survivors scaled by 1/(1-p)        */
  uint64_t seed, offset;        /* Philox4x32-10 stream for the dropout mask (offset < 2^44)  */
  const uint64_t* rng_state;    /* device {seed, step} or NULL: mixed into the stream, so a
                                   captured CUDA graph draws a fresh mask on every replay       */
  void* D; int64_t ldd;         /* [M,N]                                                      */
} llp_gemm_nt_args;
int llp_gemm_nt(const llp_gemm_nt_args* host_args, void* stream);

/* Weight gradient: D[N1,N2] (fp32) = A[M,N1]^T * B[M,N2], reduction over the M rows, split
 * across CTAs with a deterministic fixed-order reduce (autograd of F.linear). */
size_t llp_gemm_tn_workspace_bytes(int64_t M, int64_t N1, int64_t N2);
int llp_gemm_tn(int dtype, int backend, int64_t M, int64_t N1, int64_t N2, const void* A, int64_t lda,
                const void* B, int64_t ldb, float* D, int64_t ldd, int accumulate, void* workspace,
                size_t workspace_bytes, void* stream);

/* Fused weight gradient of one layer, everything the backward of `lin_l(agg) + lin_r(x)` (PyG SAGEConv;
 * sageconv_updated.py:71-76) or of one nn.Linear (models.py:48,143) needs from the output gradient G[M,N1]:
 *     dWa[N1,N2a] (+)= G^T A[M,N2a]     dWb[N1,N2b] (+)= G^T B[M,N2b]  (N2b = 0: none)     dbias[N1] (+)= colsum(G)
 * bf16 with N2a, N2b <= 256: ONE tcgen05 kernel that reads G, A and B once (split over M, deterministic reduce);
 * otherwise the separate kernels above.  dbias may be NULL.  accumulate != 0 adds into the outputs. */
size_t llp_wgrad_workspace_bytes(int64_t M, int64_t N1, int64_t N2a, int64_t N2b);
int llp_wgrad(int dtype, int backend, int64_t M, int64_t N1, const void* G, int64_t ldg, int64_t N2a, const void* A,
              int64_t lda, float* dWa, int64_t ldwa, int64_t N2b, const void* B, int64_t ldb, float* dWb, int64_t ldwb,
              float* dbias, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Column sums (bias gradient): out[n] (+)= sum_m A[m,n].  Deterministic. */
size_t llp_colsum_workspace_bytes(int64_t N);
int llp_colsum(int dtype, const void* A, int64_t lda, int64_t M, int64_t N, float* out, int accumulate,
               void* workspace, void* stream);
/* dst[c,r] = (dst dtype) src[r,c]; also plain cast when transpose == 0. */
int llp_cast2d(int src_dtype, int dst_dtype, const void* src, int64_t lds, int64_t rows, int64_t cols,
               void* dst, int64_t ldd, int transpose, void* stream);
/* bf16 working copies of fp32 master weights, all matrices in ONE launch: dst[rows, cols] (ld) = bf16(src) and/or
 * dst_t[cols, rows] (ld_t) = bf16(src)^T.  src is contiguous [rows, cols].  Run once per optimiser step instead of a
 * cast / transpose launch per weight and use (F.linear / its autograd read the same weights several times per step). */
typedef struct llp_weight_desc {
  const float* src; int64_t rows, cols;
  void* dst; int64_t ld;        /* may be NULL */
  void* dst_t; int64_t ld_t;    /* may be NULL */
} llp_weight_desc;
int llp_weights_prep(int count, const llp_weight_desc* host_descs, void* stream);
/* y = dropout(relu(a + addend + bias)): the llp_gemm_nt epilogue as a stand-alone pass, same dropout stream
 * (sageconv_updated.py:71-81 + models.py:116-117: the pre-activation is aggregate(lin_l x) + lin_r x). */
int llp_add_act(int dtype, const void* a, int64_t lda, const void* addend, int64_t ldadd, const float* bias, int64_t M,
                int64_t N, int relu, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* rng_state, void* y,
                int64_t ldy, void* stream);
/* y = gate>0 ? g*scale : 0  (relu/dropout backward from the saved forward output). */
int llp_gate(int dtype, const void* g, int64_t ldg, const void* gate, int64_t ldgate, int64_t M, int64_t N,
             float scale, void* y, int64_t ldy, void* stream);

/* ---------------------------------------------------------------------------------------
 * Edge scoring.  Replaces the two advanced-index gathers + mul feeding LinkPredictor
 * (train_teacher_gnn.py:58,97,103,109,115; main.py:186,214; models.py:140) and their
 * autograd (index_put_ with atomics).
 * ------------------------------------------------------------------------------------- */
/* z[m,:] = h[u[m],:] * h[v[m],:] */
int llp_edge_hadamard(int dtype, const void* h, int64_t ldh, int64_t feat, const int64_t* u, const int64_t* v,
                      int64_t num_edges, void* z, int64_t ldz, void* stream);
/* Incidence plan of an edge batch: the 2*num_edges (node, edge) incidences stably sorted by node (own counting sort:
 * per-node counts -> exclusive scan -> scatter -> rows put back into incidence order; no library sort).
 * rowptr[n] = first sorted position of node n (int32[num_nodes+1]); meta[2p], meta[2p+1] = (edge m, OTHER endpoint of m)
 * of sorted position p (int32[4*num_edges]).  Depends only on u and v: callers run it early / on a side stream. */
size_t llp_edge_plan_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int llp_edge_plan(const int64_t* u, const int64_t* v, int64_t num_edges, int64_t num_nodes, int32_t* rowptr, int32_t* meta,
                  void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the gather: gh[n,:] = sum_{u[m]==n} dz[m,:]*h[v[m],:] + sum_{v[m]==n} dz[m,:]*h[u[m],:] for EVERY node row
 * n < num_nodes (rows no edge touches are written as zeros), gh in the activation dtype.  Gather-reduce over the plan:
 * no atomics, bit-reproducible. */
size_t llp_edge_hadamard_bwd_workspace_bytes(int64_t num_edges);
int llp_edge_hadamard_bwd(int dtype, const void* h, int64_t ldh, int64_t feat, int64_t num_edges, const void* dz,
                          int64_t lddz, int64_t num_nodes, const int32_t* rowptr, const int32_t* meta, void* gh,
                          int64_t ldgh, void* workspace, size_t workspace_bytes, void* stream);
/* Fused edge scorer, bf16 (the gather-Hadamard is the A-operand producer of the first predictor GEMM):
 *   z[m,:] = h[u[m],:]*h[v[m],:] ; y[m,:] = dropout(relu(z[m,:] W1^T + bias1)) ; prob[m] = sigmoid(y[m,:].w2 + b2)
 * (models.py:139-150 with num_layers == 2 and out_channels == 1) in one tcgen05 kernel.  z and y are optional outputs
 * (training keeps them for the backward pass; evaluation leaves them NULL: two row gathers in, four bytes out per
 * edge); prob is optional when y is wanted.  The dropout stream is the one llp_gemm_nt draws (same seed / offset /
 * rng_state => same mask).  K % 64 == 0, N % 8 == 0, N <= 256, W1 column tile <= 128 KB (llp_edge_mlp_supported). */
typedef struct llp_edge_mlp_args {
  const void* h; int64_t ldh;           /* bf16 [num_nodes, K] */
  const int64_t* u; const int64_t* v;   /* [M] */
  int64_t M, K, N;
  const void* W1; int64_t ldw1;         /* bf16 [N, K] */
  const float* bias1;                   /* [N] or NULL */
  int relu; float dropout_p;
  uint64_t seed, offset; const uint64_t* rng_state;
  void* z; int64_t ldz;                 /* bf16 [M, K] or NULL */
  void* y; int64_t ldy;                 /* bf16 [M, N] or NULL */
  const float* w2; const float* b2;     /* [N], [1] (b2 may be NULL) */
  float* prob;                          /* [M] or NULL */
} llp_edge_mlp_args;
int llp_edge_mlp_supported(int64_t K, int64_t N);
int llp_edge_mlp_fused(const llp_edge_mlp_args* host_args, void* stream);
/* Final predictor layer with one output: logit[m] = y[m,:].w + b ; p = sigmoid(logit).
 * (models.py:146,150 with out_channels == 1.) */
int llp_score_head(int dtype, const void* y, int64_t ldy, int64_t M, int64_t H, const float* w, const float* b,
                   float* prob, void* stream);
/* backward of llp_score_head: dlogit = dprob*p*(1-p); gy[m,:] = dlogit[m]*w ; gw = sum_m dlogit*y ; gb = sum dlogit.
 * gate_scale > 0 additionally applies the relu/dropout backward of the layer that produced y:
 * gy = y > 0 ? gy*gate_scale : 0 (models.py:144-145). */
int llp_score_head_bwd(int dtype, const void* y, int64_t ldy, int64_t M, int64_t H, const float* w,
                       const float* prob, const float* dprob, float gate_scale, void* gy, int64_t ldgy, float* gw,
                       float* gb, void* workspace, size_t workspace_bytes, void* stream);
size_t llp_score_head_bwd_workspace_bytes(int64_t M, int64_t H);

/* ---------------------------------------------------------------------------------------
 * Losses.  BCE: nn.BCELoss (train_teacher_gnn.py:33,59; main.py:162,215).  LLP_D: kl_loss
 * (main.py:27-31,188).  LLP_R: the rank block (main.py:190-203).  Each writes the scalar
 * loss (deterministic reduction) and the gradient w.r.t. its student input for
 * upstream-gradient 1.
 * ------------------------------------------------------------------------------------- */
size_t llp_loss_workspace_bytes(int64_t rows);
/* labels: first n_pos entries are 1, the rest 0 (train_teacher_gnn.py:57). */
int llp_bce(const float* prob, int64_t n, int64_t n_pos, float* loss, float* dprob, void* workspace,
            void* stream);
int llp_kd_d(const float* s, const float* t, int64_t rows, int64_t K, float T, float* loss, float* ds,
             void* workspace, void* stream);
int llp_kd_r(const float* s, const float* t, int64_t rows, int64_t K, float margin, float* loss, float* ds,
             void* workspace, void* stream);

/* LLP_D and LLP_R of the same score rows in one pass (main.py:188,190-203 read the same s_r / t_r):
 * losses[0] = LLP_D, losses[1] = LLP_R (the arithmetic and reduction tree of llp_kd_d / llp_kd_r), losses[2] = w_d*LLP_D + w_r*LLP_R,
 * ds = d losses[2] / d s (or NULL).  1 < K <= 1024. */
size_t llp_kd_fused_workspace_bytes(int64_t rows);
int llp_kd_fused(const float* s, const float* t, int64_t rows, int64_t K, float T, float margin, float w_d, float w_r,
                 float* losses /*[3]*/, float* ds, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------
 * Hits@K.  Replaces ogb Evaluator._eval_hits = torch.topk on CPU tensors + compare + sum
 * (train_teacher_gnn.py:120-145,226-249).  llp_topk_desc returns the kmax largest negative
 * scores in descending order (padded with -inf when n < kmax) — the per-rank candidate list
 * of the multi-GPU exchange; llp_count_greater counts positives strictly above each threshold.
 * ------------------------------------------------------------------------------------- */
size_t llp_topk_workspace_bytes(int64_t n, int64_t kmax);
int llp_topk_desc(const float* scores, int64_t n, int64_t kmax, float* out /*[kmax]*/, void* workspace,
                  size_t workspace_bytes, void* stream);
int llp_count_greater(const float* pos, int64_t n_pos, const float* thresholds, int64_t n_thr,
                      int64_t* counts /*[n_thr]*/, void* stream);

/* ---------------------------------------------------------------------------------------
 * ROC-AUC.  Replaces sklearn.metrics.roc_auc_score on host copies of the scores
 * (train_teacher_gnn.py:147-153,251-266).  pairs[0] = #{(p,n): neg_n < pos_p}, pairs[1] = #{(p,n): neg_n == pos_p}
 * (-0.0 == +0.0); AUC = (pairs[0] + 0.5*pairs[1]) / (n_pos*n_neg).  Integer counts: bit-exact, additive over shards of
 * the positives (the multi-GPU exchange all-gathers the negatives and all-reduces `pairs`).  n_neg < 2^31.
 * ------------------------------------------------------------------------------------- */
size_t llp_auc_workspace_bytes(int64_t n_neg);
int llp_auc_pairs(const float* pos, int64_t n_pos, const float* neg, int64_t n_neg, int64_t* pairs /*[2]*/,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Negative sampling (host).  PyG 2.2.0 `negative_sampling` draws its candidate ids with CPython's
 * `random.sample(range(n), k)` (train_teacher_gnn.py:50-51; main.py:81-82,206-207 -> torch_geometric/utils/
 * negative_sampling.py: sample()).  This is that algorithm in C++ on CPython's own MT19937 state, so the sampled
 * indices stay bit-exact while the ~0.4 us/draw Python loop disappears: mt_state = the 625 uint32 of
 * `random.getstate()[1]` (624 words + position), advanced in place; out[k] = the sample.  Pure host code, no GPU needed.
 * ------------------------------------------------------------------------------------- */
int llp_py_random_sample(uint32_t* host_mt_state /*[625] in/out*/, uint64_t n, int64_t k, int64_t* host_out /*[k]*/);

/* Device half of PyG 2.2.0 negative_sampling(method='dense') (train_teacher_gnn.py:50-51; main.py:81-82,206-207): of the k
 * candidate ids (the host's random.sample), keep — in order — those that are not in taken_sorted (the sorted linearised
 * ids row*(N-1)+col' of the existing non-self-loop edges; what indexing PyG's N*N-N mask answers), write the first max_out
 * to kept[max_out] and, de-linearised (r = id/(N-1), c = id%(N-1), c += r <= c), to edges[2,max_out] (optional);
 * *count = how many were kept in total (device int32; may exceed max_out). */
size_t llp_negative_filter_workspace_bytes(int64_t k);
int llp_negative_filter(const int64_t* cand, int64_t k, const int64_t* taken_sorted, int64_t n_taken, int64_t num_nodes,
                        int64_t max_out, int64_t* kept, int64_t* edges, int32_t* count, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Context sampling.  Replaces torch_cluster.random_walk(coalesced=False) uniform kernel
 * (main.py:37,43,45).  rand is the [B,L] fp32 torch.rand tensor; out is [B,L+1] int64.
 * ------------------------------------------------------------------------------------- */
int llp_random_walk(const int64_t* rowptr, const int64_t* col, const int64_t* start, const float* rand,
                    int64_t num_walks, int64_t walk_length, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimiser tail.  Replaces clip_grad_norm_ x2 + torch.optim.Adam foreach kernels
 * (train_teacher_gnn.py:63-67; main.py:226-230) on flat fp32 buffers.  Groups are
 * contiguous ranges [group_begin[g], group_begin[g+1]) clipped separately to max_norm.
 * ------------------------------------------------------------------------------------- */
/* state[1] += 1 (one launch per training step; llp_gemm_nt_args.rng_state points at `state`). */
int llp_rng_advance(uint64_t* state, void* stream);
size_t llp_clip_adam_workspace_bytes(int num_groups);
int llp_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  const int64_t* host_group_begin, int num_groups, float max_norm, float grad_scale, float lr,
                  float beta1, float beta2, float eps, int64_t step, int64_t* device_step /*NULL, or a device counter
                  that is incremented and used instead of `step` (CUDA-graph safe)*/, void* bf16_copy /*[n] or NULL*/,
                  float* group_norms /*[num_groups] out*/, void* workspace, void* stream);

/* Deterministic sum of n floats (double accumulate): out[0] = scale * sum(in). */
int llp_sum(const float* in, int64_t n, float scale, float* out, void* workspace /* >= 8 KiB */, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* LLP_B200_H_ */
