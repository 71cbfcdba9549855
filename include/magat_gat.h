/*
 * magat_gat.h -- C ABI of the B200-native batched graph-attention layer.
 *
 * Drop-in boundary for ONE hot path of proroklab/magat_pathplanning:
 *   utils/graphUtils/graphML.py:4506-4685  class GraphFilterBatchAttentional
 *   utils/graphUtils/graphML.py:1724-1827  graphAttentionLSIGFBatch_{KeyQuery,modified}
 *   utils/graphUtils/graphML.py:1180-1286  learnAttentionGSOBatch_KeyQuery
 *   utils/graphUtils/graphML.py:713-823    learnAttentionGSOBatch (GAT_modified)
 * The reference has no FFI of its own (pure PyTorch); the entry points below are what a
 * ctypes/cffi binding for that path binds (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless stated;
 *   - the library never allocates device memory: the caller passes every buffer;
 *   - return value 0 = success, non-zero = error code (MAGAT_E_*), message through
 *     magat_last_error(); nothing throws or exits across this boundary;
 *   - E (edge features) is fixed to 1, the only value the planners use
 *     (graphs/models/decentralplanner_GAT.py:179).
 *
 * Tensor layouts (fp32 unless noted; B batch, N nodes, G in-features, F out-features per
 * head, K taps, P heads, D = ELL width >= max in/out degree over the batch, W = ceil(N/32)):
 *   x        [B][N][G]   node-major rows (the planner's memory layout, decentralplanner_GAT.py:304);
 *                        row stride x_sn >= G elements, batch stride x_sb, feature stride 1
 *   S        [B][1][N][N] dense GSO as the reference passes it, fp32 or fp64; only
 *                        |S| > 1e-9 is used (graphML.py:1274-1276)
 *   y        element (b, n, c) at y[b*y_sb + n*y_sn + c*y_sc]; c = p*F+f when concatenating
 *                        (graphML.py:4656-4662), c = f when averaging heads (:4665-4667)
 *   rowbits  [B][N][W]   uint32, bit j%32 of word j/32 of row i set iff edge (i,j)
 *   colbits  [B][N][W]   uint32, bit i%32 of word i/32 of row j set iff edge (i,j)
 *   nbr_out  [B][N][D]   int32, receivers j of sender-row i, ascending, -1 padded
 *   nbr_in   [B][N][D]   int32, senders i of column j, ascending, -1 padded
 *   slot_in  [B][N][D]   int32, position of j inside nbr_out[b][i][:] for the same entry of nbr_in
 *   slot_out [B][N][D]   int32, position of i inside nbr_in[b][j][:] for the same entry of nbr_out
 *   att      [B][N][D][P] attention value A_p[i, nbr_out[i][s]] (row-softmax, graphML.py:1284)
 *   taps     [B][N][P][K-1][G]  u_k = u_{k-1} A for k = 1..K-1 (graphML.py:1756-1759)
 *   sproj    KeyQuery: [B][N][P][G], R_i^p = W_p^T x_i so that e_p[i,j] = R_i^p . x_j (:1257-1262);
 *            GAT_modified: [B][N][P][2], {a1_p.z_n, a2_p.z_n} with z = W_p x + wb_p (:777-789)
 */
#ifndef MAGAT_GAT_H_
#define MAGAT_GAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGAT_ABI_VERSION 13

enum {
  MAGAT_OK = 0,
  MAGAT_E_BAD_ARG = 1,      /* shape / mode / null pointer; mirrors the reference's asserts */
  MAGAT_E_UNSUPPORTED = 2,  /* e.g. KeyQuery with F != G (graphML.py:1728,1765), E != 1 */
  MAGAT_E_ALIGN = 3,        /* pointer or stride not aligned as required */
  MAGAT_E_CUDA = 4,         /* launch / runtime failure (cudaGetLastError) */
  MAGAT_E_DEVICE = 5        /* not an sm_100 device */
};

/* MAGAT_MODE_GSO_VALUES: the non-attentional graph filter GraphFilterBatch / BatchLSIGF (graphML.py:5485-5700, SURVEY 8f
 * row f2): the caller fills att[B][N][D][1] with the GSO's own values (magat_gso_edge_values) and magat_gat_forward /
 * magat_gat_backward skip the score, softmax and attention-parameter parts; P must be 1. */
enum { MAGAT_MODE_KEYQUERY = 0, MAGAT_MODE_GAT_MODIFIED = 1, MAGAT_MODE_GSO_VALUES = 2 };
enum { MAGAT_DT_F32 = 0, MAGAT_DT_F64 = 1 };
/* which implementation magat_gat_forward uses for the dense projections */
enum { MAGAT_PATH_AUTO = 0, MAGAT_PATH_SIMT = 1, MAGAT_PATH_TCGEN05 = 2 };

int magat_abi_version(void);
const char* magat_last_error(void);

/* Device check: 0 when the current device can run this library's sm_100a code. */
int magat_device_check(void);

/* ---- GSO -> adjacency (replaces graphML.py:1274-1278 / :808-812 mask construction) ----
 * One pass over S.  stats[0] = max out-degree, stats[1] = max in-degree, stats[2] = total
 * number of edges (saturating), stats[3] = 1 iff the mask is symmetric.  stats must be
 * zero-initialised by the caller except stats[3] = 1. */
int magat_gso_scan(const void* S, int s_dtype, int B, int N,
                   uint32_t* rowbits, uint32_t* colbits, int32_t* stats, void* stream);

/* The same scan with the predicate of the NON-attentional filter: BatchLSIGF multiplies by S itself
 * (graphML.py:5569-5572), so every entry that is not exactly zero is an edge (NaN included). */
int magat_gso_scan_nonzero(const void* S, int s_dtype, int B, int N,
                           uint32_t* rowbits, uint32_t* colbits, int32_t* stats, void* stream);

/* att[b][i][s] = (float) S[b][i][nbr_out[b][i][s]] (0 beyond the degree): the edge weights MAGAT_MODE_GSO_VALUES uses in
 * place of an attention. */
int magat_gso_edge_values(const void* S, int s_dtype, const int32_t* nbr_out, int B, int N, int D, float* att,
                          void* stream);

/* ---- GSO in HOST memory ----
 * The reference builds the GSO on the CPU (dataloader / simulator); testing 4 N^2 bytes per instance on the device means
 * shipping them over PCIe first.  magat_gso_pack_host (plain C++, no CUDA; multi-threaded, AVX2 when the CPU has it) makes
 * the row mask on the host cores -- rowbits_host[rows = B * N][W], same layout and predicate as magat_gso_scan -- so
 * N^2 / 8 bytes cross the link; after the caller's H2D copy magat_gso_from_rowbits transposes it into colbits on the
 * device and fills stats like magat_gso_scan (stats zero-initialised by the caller except stats[3] = 1).  Continue
 * with magat_gso_build_ell.  threads <= 0: one per hardware thread. */
int magat_gso_pack_host(const void* S_host, int s_dtype, long rows, int N, uint32_t* rowbits_host, int threads);
int magat_gso_from_rowbits(const uint32_t* rowbits, int B, int N, uint32_t* colbits, int32_t* stats, void* stream);

/* GAT_origin (graphML.py:964-1070) tests the edges of S + I (:1019): after magat_gso_scan, sets bit (i, i) of both masks to
 * |float(S_ii) + 1| > 1e-9 and recomputes stats.  Then magat_gso_build_ell as usual. */
int magat_gso_self_loops(const void* S, int s_dtype, int B, int N, uint32_t* rowbits, uint32_t* colbits, int32_t* stats,
                         void* stream);

/* Bit masks -> padded neighbour lists of width D (D >= max(stats[0], stats[1]), D >= 1). */
int magat_gso_build_ell(const uint32_t* rowbits, const uint32_t* colbits, int B, int N, int D,
                        int32_t* nbr_out, int32_t* nbr_in, int32_t* slot_in, int32_t* slot_out /* may be NULL */,
                        void* stream);

/* ---- the step BEFORE the path.  Edge mask straight from agent positions, replacing the CPU
 * squareform(pdist(pos)) < commR + zero diagonal of utils/new_simulator.py:823-827 and the 4N^2-byte dense GSO that
 * then crosses PCIe.  pos: [B][N][2] fp32 or fp64 (device); the predicate is evaluated in fp64 like scipy's.  Writes the
 * same rowbits / colbits / stats as magat_gso_scan (stats zero-initialised by the caller except stats[3] = 1);
 * continue with magat_gso_build_ell.  N <= 3072.  Neighbour search through a cell list (cells of the radius, 3 x 3 per
 * agent; an instance with non-finite positions or a bounding box of more than 4096 cells is tested all pairs): the
 * result is that of the all-pairs formula bit for bit. */
int magat_gso_from_positions(const void* pos, int pos_dtype, int B, int N, double comm_radius,
                             uint32_t* rowbits, uint32_t* colbits, int32_t* stats, void* stream);

/* ---- forward (replaces GraphFilterBatchAttentional.forward, graphML.py:4636-4667) ---- */
typedef struct magat_gat_fwd_args {
  int32_t B, N, G, F, K, P, D;
  int32_t mode;          /* MAGAT_MODE_* */
  int32_t concat;        /* 1: activation then concat heads; 0: mean over heads then activation */
  int32_t relu;          /* 1: fused ReLU (the reference default, graphML.py:4560); 0: identity */
  int32_t path;          /* MAGAT_PATH_* */
  int32_t reserved;
  /* inputs */
  const float* x; int64_t x_sb, x_sn;
  const int32_t* nbr_out; const int32_t* nbr_in; const int32_t* slot_in;
  const int32_t* slot_out;    /* may be NULL (then ain is filled by the gather kernel instead of the attention kernel) */
  /* parameters, in the reference's shapes (graphML.py:4579-4597) */
  const float* weight;        /* KeyQuery [P][1][G][G]; GAT_modified [P][1][F][G] */
  const float* mixer;         /* [P][1][2F] (GAT_modified only; may be NULL for KeyQuery) */
  const float* weight_bias;   /* [P][1][F]  (GAT_modified only) */
  const float* filterWeight;  /* [P][F][1][K][G] */
  const float* bias;          /* [F][1] or NULL */
  /* outputs */
  float* y; int64_t y_sb, y_sn, y_sc;
  float* att;                 /* [B][N][D][P] */
  float* ain;                 /* [B][N][P][D] receiver-major copy of att: only the general (non-vectorised) gather kernels
                                 use it, and only when slot_out is given; may be NULL */
  float* taps;                /* [B][N][P][K-1][G] (unused when K == 1) */
  /* scratch (caller allocated) */
  float* wprep;               /* magat_gat_wprep_floats(...) floats */
  float* sproj;               /* KeyQuery: [B][N][P][G]; GAT_modified: [B][N][P][2] */
  /* optional (may be NULL): the ReLU mask of y, one bit per element, for magat_gat_backward -- word [m >> 5][c] holds
   * (y[m][c] > 0) for the 32 node rows m = 32 (m >> 5) .. + 31 (m = b * N + n) of channel c = p * F + f.
   * magat_gat_relu_bits_words(B, N, P, F) words; written only when magat_gat_forward_relu_bits_valid(a) says so. */
  uint32_t* relu_bits;
} magat_gat_fwd_args;

size_t magat_gat_wprep_floats(int G, int F, int K, int P, int mode);
int magat_gat_forward(const magat_gat_fwd_args* a, void* stream);
/* How many tap planes (k = 1..) of a->taps the forward call leaves valid for these arguments (K-1 today).  Pass it on
 * as bwd.taps_valid; magat_gat_backward rebuilds the planes above it. */
int magat_gat_forward_taps_valid(const magat_gat_fwd_args* a);
/* 1 when magat_gat_forward fills a->relu_bits for these arguments (the tcgen05 K-tap projection runs and relu = 1);
 * only then may the buffer be passed on as bwd.relu_bits. */
int magat_gat_forward_relu_bits_valid(const magat_gat_fwd_args* a);
size_t magat_gat_relu_bits_words(int B, int N, int P, int F);


/* ---- heads averaged on top of the concat layout (graphML.py:4665-4667: y = act(mean_p y_p), contiguous [B][F][N]) ----
 * ycat / dycat: [B*N][P*F] (what magat_gat_forward writes with concat = 1, relu = 0, unit channel stride); y, dy: contiguous
 * [B][F][N].  forward: y = act(mean over heads); backward: dycat[.][p*F + f] = dy * act'(y) / P for every head.
 * F a multiple of 32. */
int magat_head_mean_forward(const float* ycat, int B, int N, int P, int F, int relu, float* y, void* stream);
int magat_head_mean_backward(const float* dy, const float* y, int B, int N, int P, int F, int relu, float* dycat,
                             void* stream);

/* ---- layer + linear action head in one pass (inference) ----
 * The planner feeds the layer's output straight into actionsMLP (graphs/models/decentralplanner_GAT.py:329-334; one
 * nn.Linear(P*F -> 5) in the published configurations) and decodes the action as argmax softmax
 * (utils/new_simulator.py:863-869).  This call runs magat_gat_forward with the head folded into the epilogue of the K-tap
 * projection: y (4*P*F bytes per agent) is never written.  a->y is ignored (may be NULL), a->relu / a->bias apply as usual.
 * head_weight [A][P*F] and head_bias [A] (or NULL) are the nn.Linear parameters, A <= 8; partial is P * B*N * 8 floats of
 * scratch (16 B aligned); logits [B*N][A]; actions_or_null [B*N] receives argmax_a (first maximum on ties, as torch.max).
 * Covers what the tcgen05 K-tap projection covers: concatenated heads, F = 128, G a multiple of 128, K <= 3
 * (magat_gat_actions_supported); MAGAT_E_UNSUPPORTED otherwise. */
int magat_gat_actions_supported(const magat_gat_fwd_args* a, int A);
int magat_gat_forward_actions(const magat_gat_fwd_args* a, const float* head_weight, const float* head_bias, int A,
                              float* partial, float* logits, int32_t* actions_or_null, void* stream);


/* ---- fused forward: ONE launch from the dense GSO to y (graphML.py:4636-4667 over :1724-1827, :1180-1286, :713-823) ----
 * Covers G = F = 128, K <= 3, P in {1,2,4}, heads concatenated, N % 4 == 0, N >= 64, both attention modes
 * (magat_gat_fused_supported).  Reads S once and never materialises anything N x N; the per-instance intermediates
 * (bit masks, operand images, receiver-major attention, and -- unless save = 1 -- R and the taps) live in a per-team
 * scratch inside `workspace` that stays in L2.  D is the caller's cap on the in/out degree (multiple of 4, <= 32):
 * after the call workspace[0..3] (int32, device) hold {max out-degree, max in-degree, number of (node, direction)
 * lists longer than D, 0}; when the third is non-zero the outputs are INVALID and the caller must redo the call with a
 * larger D or through magat_gso_scan / magat_gso_build_ell / magat_gat_forward.
 * Always written: y, nbr_out / nbr_in / slot_in [B][N][D] and att [B][N][D][P] (what returnAttentionGSO and
 * magat_gat_backward need; slot_out is not used by this path and may be NULL).  save = 1 (training) additionally writes taps [B][N][P][K-1][G], sproj and wprep
 * (GAT_modified: magat_gat_wprep_floats floats) exactly as magat_gat_forward leaves them for magat_gat_backward. */
typedef struct magat_gat_fused_args {
  int32_t B, N, G, F, K, P, D;
  int32_t mode, concat, relu;
  int32_t s_dtype;            /* MAGAT_DT_* of S */
  int32_t save;               /* 1: keep taps / sproj / wprep for magat_gat_backward */
  int32_t team;               /* CTAs per planning instance: 0 = default (8), or 8 / 16.  16 halves the instances in flight:
                                 the per-team scratch then fits L2 with room to spare (DRAM traffic 4.0 instead of 8.0 GB at
                                 B = 512, N = 1000) at ~25 % more time */
  int32_t reserved;
  const void* S;              /* [B][1][N][N] */
  const float* x; int64_t x_sb, x_sn;
  const float* weight; const float* mixer; const float* weight_bias; const float* filterWeight; const float* bias;
  float* y; int64_t y_sb, y_sn, y_sc;
  int32_t* nbr_out; int32_t* nbr_in; int32_t* slot_in; int32_t* slot_out;
  float* att;
  float* taps; float* sproj; float* wprep;      /* save = 1 only; may be NULL otherwise */
  void* workspace; size_t ws_bytes;             /* magat_gat_fused_workspace_bytes(...), 1024 B aligned */
} magat_gat_fused_args;

int magat_gat_fused_supported(int N, int G, int F, int K, int P, int D, int mode, int concat);
size_t magat_gat_fused_workspace_bytes(int B, int N, int K, int P, int D, int mode, int save, int team);
int magat_gat_forward_fused(const magat_gat_fused_args* a, void* stream);

/* ---- backward (what autograd does over graphML.py:1180-1286,713-823,1724-1827) ---- */
typedef struct magat_gat_bwd_args {
  int32_t B, N, G, F, K, P, D;
  int32_t mode, concat, relu, path;
  int32_t need_dx, need_dweight, need_dfilter, need_dbias, need_dmixer;   /* requires_grad gating */
  int32_t taps_valid;   /* tap planes k = 1..taps_valid of `taps` hold data (magat_gat_forward_taps_valid) */
  int32_t reserved;
  /* saved from forward */
  const float* x; int64_t x_sb, x_sn;
  const int32_t* nbr_out; const int32_t* nbr_in; const int32_t* slot_in;
  const float* weight; const float* mixer; const float* weight_bias; const float* filterWeight;
  const float* y; int64_t y_sb, y_sn, y_sc;   /* forward output (post activation) */
  const float* att;
  float* taps;                /* planes above taps_valid are (re)computed here */
  const float* wprep; const float* sproj;
  /* incoming gradient, same logical shape as y, own strides */
  const float* dy; int64_t dy_sb, dy_sn, dy_sc;
  /* outputs: written (not accumulated); any may be NULL when the matching need_* is 0 */
  float* dx;                  /* [B][N][G] contiguous */
  float* dweight; float* dmixer; float* dweight_bias; float* dfilterWeight; float* dbias;
  /* scratch */
  float* gz;                  /* [B][N][P][K][G] */
  float* datt;                /* [B][N][D][P] */
  float* rc;                  /* KeyQuery: [B][N][P][G]; GAT_modified: [B][N][P][2] */
  float* partial;             /* magat_gat_bwd_partial_floats(...) floats */
  /* optional (may be NULL): the forward's bit mask of y > 0 (fwd.relu_bits, only when ..._relu_bits_valid): the two
   * dense kernels that need relu'(y) then read 1/32 of the bytes of y */
  const uint32_t* relu_bits;
} magat_gat_bwd_args;

size_t magat_gat_bwd_partial_floats(int B, int N, int G, int F, int K, int P, int mode);
int magat_gat_backward(const magat_gat_bwd_args* a, void* stream);

/* The same call with ONE scratch buffer: a->gz, a->datt, a->rc and a->partial are ignored and carved out of `workspace`
 * (256 B aligned, magat_gat_backward_workspace_bytes(...) bytes) -- their layouts then are no part of the caller's
 * contract.  This is the form the host mirror uses. */
size_t magat_gat_backward_workspace_bytes(int B, int N, int G, int F, int K, int P, int D, int mode);
int magat_gat_backward_ws(const magat_gat_bwd_args* a, void* workspace, size_t ws_bytes, void* stream);

/* ---- lazy dense attention for returnAttentionGSO (graphML.py:4623-4634, :4650) ----
 * Expands att into aij[B][P][1][N][N] (mean_heads == 0), or into the head-mean [B][1][N][N]
 * that returnAttentionGSO returns.  `out` must be zero filled by the caller. */
int magat_gat_attention_dense(const float* att, const int32_t* nbr_out, int B, int N, int D, int P,
                              int mean_heads, float* out, void* stream);

/* ---- small-graph forward (inference): the whole layer for one instance in one CTA, ONE launch ----
 * For the simulator loop of the reference (B = 1, N <= 64 agents per step under torch.no_grad(),
 * agents/decentralplannerlocal_OnlineExpert_GAT.py:1039-1044).  Takes the dense GSO directly; writes y and, when
 * aij_or_null is given, the dense attention [B][P][N][N] (graphML.py:4650).  Nothing is kept for backward. */
int magat_gat_small_supported(int N, int G, int F, int K, int P, int concat);
int magat_gat_forward_small(const void* S, int s_dtype, const float* x, int64_t x_sb, int64_t x_sn,
                            const float* weight, const float* mixer, const float* weight_bias,
                            const float* filterWeight, const float* bias, float* y, int64_t y_sb, int64_t y_sn,
                            int64_t y_sc, float* aij_or_null, int B, int N, int G, int F, int K, int P, int mode,
                            int concat, int relu, void* stream);

/* ---- launch accounting / measurement hooks (bench.py, tests) ----
 * magat_launch_count: kernels launched by this library in this process so far.
 * magat_profile_enable(1): from now on record a CUDA event behind every launch (and at every entry
 * point) on the launching stream; magat_profile_collect synchronises those events and writes
 * "kernel name,launches,total_ms" lines, aggregated per kernel, into buf (returns bytes needed). */
long magat_launch_count(void);
void magat_profile_enable(int on);
long magat_profile_collect(char* buf, long cap);

#ifdef __cplusplus
}
#endif
#endif /* MAGAT_GAT_H_ */
