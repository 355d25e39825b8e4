/*
 * pesto_b200 -- C ABI of the B200-native (sm_100a) PeSTo forward path.
 *
 * Every entry point replaces one piece of the reference's Python hot path (citations are
 * file:line into LBM-EPFL/PeSTo).  The reference has no FFI of its own (it is pure
 * PyTorch), so this header *is* the binding surface a maintainer would add: plain
 * pointers and sizes, no torch types.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - all `const float*`, `int64_t*` ... arguments are DEVICE pointers on the current CUDA device unless
 *     the name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); every call only
 *     enqueues work on that stream and never synchronises, allocates or frees device memory
 *     (except pesto_model_finalize / pesto_model_destroy, which own the packed weights);
 *   - return value 0 = success; otherwise a negative PESTO_E* code, with a message available
 *     from pesto_last_error() (thread-local).  Nothing aborts the process: callers wrap
 *     per-structure work in try/except and continue (interfaceome/apply_model.py:81-82).
 */
#ifndef PESTO_B200_H
#define PESTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PESTO_OK            0
#define PESTO_EINVAL       -1   /* bad argument (shape, null pointer, unsupported size) */
#define PESTO_ECUDA        -2   /* a CUDA runtime call / kernel launch failed            */
#define PESTO_ESTATE       -3   /* model not finalized, missing tensor, ...              */
#define PESTO_EINPUT       -4   /* input data the device found invalid (pesto_forward_status) */
#define PESTO_STATUS_WORDS  8   /* int32 status words of a forward's workspace              */

#define PESTO_NS           32   /* state width Ns            (model/config.py: 'Ns': 32) */
#define PESTO_NH            2   /* attention heads Nh                                    */
#define PESTO_NK            3   /* key width Nk                                          */
#define PESTO_MAX_NN       64   /* neighbours kept per atom  (extract_topology(X, 64))    */
#define PESTO_STATE_STRIDE 128  /* floats per atom record: q[32] | p_x[32] | p_y[32] | p_z[32] */
#define PESTO_NUM_OUT       5   /* logits per residue of the shipped i_v4 models (model/config.py 'dm' N2); see pesto_model_num_out */

/* arithmetic modes of the per-edge MLPs (state, softmax and accumulators are always fp32) */
#define PESTO_MODE_FP32     0   /* FFMA everywhere: parity mode                                     */
#define PESTO_MODE_F16X3    1   /* tcgen05 tensor cores, 3-term split product hi*hi + lo*hi + hi*lo over fp16 planes   */
                                /* (11 + 11 mantissa bits: fp32-accurate, logits within ~1e-4 of the reference); default */
#define PESTO_MODE_F16      2   /* tcgen05 tensor cores, single fp16 pass: speed mode, ~2e-2 on the logits              */
#define PESTO_MODE_BF16X3   PESTO_MODE_F16X3   /* former names (the planes were bf16 until the fp16 planes measured 8x  */
#define PESTO_MODE_BF16     PESTO_MODE_F16     /* more accurate at the same speed); kept as aliases                      */

typedef struct pesto_model pesto_model_t;

int         pesto_abi_version(void);
const char *pesto_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Topology.  Replaces extract_topology(X, num_nn)            src/data_encoding.py:87-102
 * and, with base = 1, also the index shift / sink padding of
 * collate_batch_features                                       src/dataset.py:100-109.
 *
 * X[n_atoms,3] holds `n_seg` independent structures back to back; seg_off[n_seg+1] (int32,
 * device) are their atom offsets.  For every atom the k nearest atoms OF ITS OWN structure are
 * selected by (masked distance, index) ascending, where the distance is bit-exact with the
 * reference's fp32 recipe sqrt(fma(dz,dz,fma(dy,dy,dx*dx))) and entries closer than 1e-2 get
 * + max(D of the structure) (src/data_encoding.py:93).
 *   base = 0: ids are 0-based inside the structure (what extract_topology returns);
 *             columns >= structure size are filled with -1 (callers slice [:, :min(k, n)]).
 *   base = 1: ids are 1-based global row numbers with 0 = sink in unfilled columns
 *             (what collate_batch_features hands to Model.forward).
 * d_out[n_atoms,k] / r_out[n_atoms,k,3] (optional, may be NULL) receive D_topk / R_topk.
 * scratch: at least pesto_knn_scratch_bytes(n_atoms, n_seg) bytes, 16-byte aligned (chunk bounding boxes + row flags).
 * ------------------------------------------------------------------------------------------- */
size_t pesto_knn_scratch_bytes(int n_atoms, int n_seg);
int    pesto_knn(const float *X, int n_atoms, const int32_t *seg_off, int n_seg, int k, int base,
                 int64_t *ids_out, float *d_out, float *r_out, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Weights.  Replaces Model(config_model) + load_state_dict(torch.load(model_ckpt.pt))
 *                                                    model/model.py:7-30, apply_model.ipynb:84-93
 * Tensors are handed over one by one under their reference state-dict key
 * ("em.0.weight", "sum.7.su.evm.2.bias", "spl.zdm_vec.0.weight", ...), fp32, row-major HOST
 * memory, shapes as in the checkpoint.  pesto_model_finalize packs them into
 * the device layout the kernels read; the model is immutable afterwards.
 * The depth of the embedding / decoder heads and the number of logits per residue are read off the
 * tensors: three Linear layers when "em.2.weight" / "dm.2.weight" are present (i_v3_0, i_v4_*:
 * model/model.py:9-30), one otherwise (model/save/i_v3_1_2021-05-28_12-40/model.py:9-22); the
 * length of the last decoder bias (<= 8) is pesto_model_num_out (5 for i_v4_*, 1 for i_v3_1).
 * ------------------------------------------------------------------------------------------- */
pesto_model_t *pesto_model_create(int n_layers, const int32_t *nn_per_layer_host, int q0_dim);
int            pesto_model_set_tensor(pesto_model_t *m, const char *key, const float *data_host, int64_t numel);
int            pesto_model_finalize(pesto_model_t *m);
void           pesto_model_destroy(pesto_model_t *m);
int            pesto_model_num_layers(const pesto_model_t *m);
int            pesto_model_num_out(const pesto_model_t *m);      /* logits per residue; 0 before finalize */
int            pesto_model_layer_nn(const pesto_model_t *m, int layer);

/* ---------------------------------------------------------------------------------------------
 * Forward, stage by stage (used by the per-layer parity tests and by pesto_forward).
 *
 * Device layouts (DESIGN.md "Data layout in HBM"):
 *   state  float[n_atoms+1][128]   row 0 = sink (all zero), row i+1 = atom i: q | p_x | p_y | p_z
 *   ids32  int32[n_atoms][64]      1-based rows into `state`, 0 = sink
 *   geom   float4[n_atoms][64]     (r_x, r_y, r_z, d) of the edge
 * ------------------------------------------------------------------------------------------- */

/* q = em(q0); p = 0; D_nn, R_nn from X and ids  --  model/model.py:34-40, src/model_operations.py:6-22.
 * ids1[n_atoms, ids_cols] int64, 1-based, 0 = sink (geometry of sink slots uses X[-1], as the reference).
 * scratch8: 8 bytes of device scratch ([0] bits of the global max of D_nn, [1] set non-zero if an id is out of
 * range). */
int pesto_prologue(const pesto_model_t *m, const float *X, const int64_t *ids1, int ids_cols,
                   const float *q0, int n_atoms, float *state, int32_t *ids32, float *geom,
                   void *scratch8, void *stream);

/* one StateUpdateLayer -- src/model_operations.py:225-242 (gathers, StateUpdate.forward :87-154, sink reset).
 * node_scratch: pesto_node_scratch_bytes(n_atoms) bytes (per-atom factors + attention sums of the layer).
 * state_in and state_out must not alias.  Tensor-core modes use three launches here (per-atom head, fused edge
 * kernel, per-atom tail); pesto_forward merges the tail of layer l with the head of layer l+1. */
size_t pesto_node_scratch_bytes(int n_atoms);
int    pesto_state_update(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32,
                          const float *geom, const float *state_in, float *state_out,
                          void *node_scratch, int mode, void *stream);

/* Measurement aid for bench.py's roofline: same as pesto_state_update, but brackets the per-atom (node) kernel
 * and the fused per-edge kernel with CUDA events on `stream`, waits for them and returns both durations in ms. */
int    pesto_state_update_timed(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32,
                                const float *geom, const float *state_in, float *state_out,
                                void *node_scratch, int mode, void *stream,
                                float *ms_node_host, float *ms_edge_host);

/* Measurement aid: the fused edge kernel of `layer` alone, launched reps times back to back on `stream` between two CUDA
 * events (after one launch of the per-atom kernel that produces the layer's factors from state_in, and one untimed
 * warm-up launch); *ms_per_launch_host = elapsed / reps.  Tensor-core modes only.  The attention sums land in the Z
 * region of node_scratch (pesto_node_scratch_bytes); state_in is not modified. */
int pesto_edge_kernel_timed(const pesto_model_t *m, int layer, int n_atoms, const int32_t *ids32, const float *geom,
                            const float *state_in, void *node_scratch, int mode, int reps, void *stream,
                            float *ms_per_launch_host);

/* dense one-hot membership M[n_atoms, n_res] (fp32, the reference's 4th forward argument) -> residue
 * column per atom.  flags[0] is set non-zero on device if some row is not one-hot. */
int pesto_residue_index(const float *M, int n_atoms, int n_res, int32_t *rid, int32_t *flags, void *stream);

/* StatePoolLayer + decoder -- src/model_operations.py:197-213, model/model.py:46-50.
 * rid[n_atoms] residue column per atom (any order); z[n_res, pesto_model_num_out(m)] (5 for the i_v4 models).
 * scratch: pesto_pool_scratch_bytes(n_atoms, n_res) bytes. */
size_t pesto_pool_scratch_bytes(int n_atoms, int n_res);
int    pesto_pool_decode(const pesto_model_t *m, const float *state, const int32_t *rid, int n_atoms,
                         int n_res, float *z, void *scratch, void *stream);

/* copy a packed state into the reference's tensors q[n_atoms+1,32], p[n_atoms+1,3,32] (test taps) */
int pesto_unpack_state(const float *state, int n_atoms, float *q, float *p, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Whole forward.  Replaces Model.forward(X, ids_topk, q0, M)                model/model.py:32-52
 * Exactly one of M (dense fp32 [n_atoms, n_res]) and rid (int32 [n_atoms]) must be non-NULL.
 * workspace: pesto_forward_workspace_bytes(n_atoms, n_res) bytes of device memory.
 * ------------------------------------------------------------------------------------------- */
size_t pesto_forward_workspace_bytes(int n_atoms, int n_res);
int    pesto_forward(const pesto_model_t *m, const float *X, const int64_t *ids1, int ids_cols,
                     const float *q0, const float *M, const int32_t *rid, int n_atoms, int n_res,
                     float *z, void *workspace, size_t workspace_bytes, int mode, void *stream);

/* Errors only the device can detect never abort and never force a synchronisation: they set a status word in the
 * workspace and make every logit NaN.  pesto_forward_status copies the PESTO_STATUS_WORDS words of the forward that
 * last used `workspace` (same n_atoms, n_res) to status_host, synchronising `stream`, and returns PESTO_EINPUT
 * ([1] a neighbour id out of range, [2] a row of M not one-hot, [4] a residue index out of range) or PESTO_ECUDA
 * ([3] = id of a tensor-core stage whose completion never arrived: the kernels' bounded waits gave up; [3] = 100: the
 * state left the range the fp16 operand planes of the tensor-core modes cover, |q| or |p| > 2^14 -> PESTO_EINPUT) with
 * pesto_last_error() set; PESTO_OK if the forward was clean.  Callers that synchronise anyway (to read z) call it there. */
int pesto_forward_status(const void *workspace, int n_atoms, int n_res, int32_t *status_host, void *stream);
/* byte offset of those status words inside the workspace (callers that pipeline forwards copy the words out with the
 * logits, before the next forward on the same workspace clears them) */
size_t pesto_forward_status_offset(int n_atoms, int n_res);

/* Debug: simulate a hung tensor-core stage (on != 0: the edge kernels never signal their third-layer GEMMs and wait only
 * briefly); pesto_debug_watchdog reads and clears the device word that staged calls (pesto_state_update) report into. */
int pesto_debug_force_watchdog(int on);
int pesto_debug_watchdog(int32_t *value_host);

/* Debug / self-test: D[128,N] = A[128,K] * B[N,K]^T on the tensor cores with exactly the operand staging of the
 * fused kernel (A thread-per-row in TMEM, B K-major un-swizzled in shared memory); split != 0 selects the 3-term
 * split-bf16 product.  lbo / sbo < 0 and idesc == 0 select the library's own descriptor values. */
int pesto_debug_umma_probe(const float *A, const float *B, float *D, int K, int N, int split,
                           int lbo, int sbo, int idesc, void *stream);

/* Debug: phase timeline of the tensor-core edge kernel.  While buf != NULL, every edge-kernel launch makes CTA 0
 * write clock64() stamps at 19 phase boundaries per tile to buf[tile][half 0..1][column group 0..1][19] (device
 * int64) for the first max_tiles tiles of each half; buf == NULL switches it off. */
int pesto_debug_edge_timeline(void *buf, int max_tiles);

/* Debug: cycles of a chain of n_mma tcgen05.mma (M = 128, K = 16, fp16 operands, zeros) with the given shared-memory operand
 * layouts (descriptor layout type 0 none / 2 SWIZZLE_128B / 4 SWIZZLE_64B / 6 SWIZZLE_32B, byte strides, bytes per K step;
 * *_major_mn != 0: MN-major operand) or with the A operand in TMEM; out (device int64[2]) = issue cycles, cycles until complete. */
int pesto_debug_mma_time(int n_mma, int N, int a_major_mn, int a_layout, int a_lbo, int a_sbo, int a_kstep, int b_major_mn,
                         int b_layout, int b_lbo, int b_sbo, int b_kstep, int a_tmem, long long *out, void *stream);

/* Debug: building blocks of the edge kernel's tensor-core reduction in isolation (csrc/rmma_probe.cu).  Gathers the
 * 128 rows ids[] of p16 (device, fp16 [n_rows][192] = hi plane 96 | lo plane 96) with TMA tile::gather4 into
 * shared memory and computes D[m][n] = sum_k (hi + lo)(ids[k], m) * W[n][k] (W: device fp32 [16][128]) as a 3-term
 * split tcgen05.mma with both operands MN-major in shared memory -> D device fp32 [128][16] (rows >= 96 unspecified);
 * Prec (device fp32 [128][96]) receives hi + lo of every gathered row as read back by threads from the swizzled tile,
 * raw (optional, 48 KB) the tile's bytes.  a_/b_ lbo/sbo < 0 and idesc == 0 select the library's own descriptor values;
 * issue_lanes: 1 = one lane issues the 192 gathers, 32 = the lanes of one warp issue six each.  status (device int[3]):
 * [0] a stage id if a wait timed out, [1] clock cycles spent issuing, [2] cycles from the first issue until all rows landed. */
int pesto_debug_rmma_probe(const void *p16, int n_rows, const int32_t *ids, const float *W, float *D, float *Prec, void *raw,
                           int a_lbo, int a_sbo, int b_lbo, int b_sbo, int idesc, int issue_lanes, int *status, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host-side PDB text parser (no device work; every pointer is a HOST pointer).  Replaces the gemmi call inside
 * read_pdb                                                                  src/structure_io.py:6-55
 * ATOM / HETATM records, fixed columns, lines cut at 80 characters; models from MODEL/ENDMDL; separated parts of a
 * chain merged per model (gemmi's chain order); altloc duplicates dropped with the reference's key
 * (chain, residue number, atom name).  Outputs are fixed-width character fields (NUL padded, not terminated):
 * name4 [cap,4], element2 [cap,2] (title case), resname3 [cap,3], het [cap] ('A' | 'H'), chain [cap], icode [cap]
 * (0 = none); xyz [cap,3], bfactor [cap] fp32; resid, model int32.  *n_out = atoms found; PESTO_EINVAL if it
 * exceeds `capacity` (pesto_pdb_count_atoms_host gives an upper bound).
 * ------------------------------------------------------------------------------------------- */
int pesto_pdb_count_atoms_host(const char *text, size_t len);
int pesto_pdb_parse_host(const char *text, size_t len, int capacity, float *xyz, char *name4, char *element2,
                         char *resname3, int32_t *resid, char *het, char *chain, int32_t *model, char *icode,
                         float *bfactor, int *n_out);

/* number of kernels one pesto_forward / pesto_knn call launches (for bench.py's gpu_launches) */
int pesto_forward_launch_count(const pesto_model_t *m, int dense_m, int mode);
int pesto_knn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PESTO_B200_H */
