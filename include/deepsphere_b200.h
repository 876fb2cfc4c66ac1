/*
 * deepsphere_b200.h — C-ABI of the B200-native DeepSphere graph-convolution hot path.
 *
 * Drop-in boundary for everything the reference executes between
 * src/deepsphere/gnn_layers.py:113 and :159 (Chebyshev.call / Monomial.call), plus the
 * nested-order pool / pseudo-convolution layers either side of it
 * (src/deepsphere/healpy_layers.py:20-216).  The reference reaches this arithmetic
 * through TensorFlow ops (tf.sparse.sparse_dense_matmul via
 * utils.split_sparse_dense_matmul utils.py:49-78, tf.matmul gnn_layers.py:149, Keras
 * MaxPool1D/AveragePooling1D/Conv1D/Conv2DTranspose); a TF custom op, a ctypes stub or
 * any other FFI binds exactly the entry points below.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every entry returns 0 on success, non-zero on error; ds_last_error() gives the
 *    message of the calling thread's last failure.
 *  - tensors are caller-owned DEVICE pointers, dense row-major fp32, in the reference's
 *    own layout [B, M, F] (batch, pixel, channel); `stream` is a cudaStream_t passed as
 *    void*.  The library owns only opaque plan handles.
 *  - there is no CPU fallback: without a CUDA device every compute entry fails.
 *  - plans are immutable after creation; any number of host threads may call the
 *    compute entries concurrently with distinct workspaces/streams.
 */
#ifndef DEEPSPHERE_B200_H
#define DEEPSPHERE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_ABI_VERSION 1

/* recursion type: gnn_layers.py:135-143 (Chebyshev) / :287-290 (Monomial) */
#define DS_RECURSION_CHEBYSHEV 0 /* T_1 = L T_0 ; T_k = 2 L T_{k-1} - T_{k-2} */
#define DS_RECURSION_MONOMIAL 1  /* T_k = L T_{k-1} */

/* arithmetic of the K*Fin -> Fout contraction (gnn_layers.py:149) and its gradients */
#define DS_MODE_FP32 0   /* CUDA-core fp32 FMA                        (parity rel <= 1e-5) */
#define DS_MODE_TF32 1   /* tcgen05 tensor cores, single-pass TF32    (stated rel <= 1e-3) */
#define DS_MODE_TF32X3 2 /* tcgen05, 3xTF32 error-compensated split   (rel <= 1e-5)        */

/* fused epilogue activation (tf.keras.activations names, gnn_layers.py:55-60) */
#define DS_ACT_LINEAR 0
#define DS_ACT_RELU 1
#define DS_ACT_ELU 2
#define DS_ACT_SIGMOID 3
#define DS_ACT_TANH 4
#define DS_ACT_SOFTPLUS 5

/* pooling type: healpy_layers.py:48-65 */
#define DS_POOL_MAX 0
#define DS_POOL_AVG 1

typedef struct ds_plan ds_plan_t;
typedef struct ds_comm ds_comm_t; /* an NCCL communicator of the ranks that share one sphere / one batch */

int ds_abi_version(void);
/* message of the calling thread's last error ("" if none) */
const char* ds_last_error(void);
/* number of visible CUDA devices (0 without a GPU); never fails */
int ds_device_count(void);
/* total number of kernel launches issued by this library in this process (bench.py's
 * gpu_launches evidence) */
int64_t ds_launch_count(void);

/* ---- plan ---------------------------------------------------------------------------
 * Replaces Chebyshev.__init__'s tf.constant triple (_L_indices int64 [nnz,2],
 * _L_values floatx [nnz], _L_shape) — gnn_layers.py:68-72 — and the per-call
 * tf.sparse.reorder (gnn_layers.py:115): the rescaled Laplacian L~ is given once, as
 * HOST COO arrays in any order; the library sorts it row-major, builds a fixed-width
 * ELL slab (width = ell_width, or chosen automatically when ell_width <= 0) plus a CSR
 * tail for longer rows, does the same for L~^T (needed by the backward pass), and
 * uploads both to the current device.
 */
int ds_plan_create_coo(int64_t M, int64_t nnz, const int64_t* indices /* [nnz,2] (row,col) */,
                       const float* values /* [nnz] */, int32_t ell_width, ds_plan_t** plan_out);
int ds_plan_destroy(ds_plan_t* plan);
/* info: 0 M, 1 nnz, 2 ell_width, 3 tail_rows, 4 tail_nnz, 5 ell_width of L~^T,
 *       6 tail_rows of L~^T, 7 device bytes held by the plan, 8 is L~ symmetric (0/1),
 *       9 has a lattice attachment (0/1) */
int ds_plan_info(const ds_plan_t* plan, int32_t what, int64_t* value_out);

/* Optional: attach the HEALPix tile plan that enables the fused K-hop recursion kernel (all K-1 hops of
 * gnn_layers.py:135-143 in shared memory) for layers with K = H + 1 on a symmetric 8-neighbour graph.
 * HOST arrays: pix [n_tiles, LW*LW] row of L~ at each lattice position (-1 = hole), w [n_tiles, LW*LW, 9]
 * stencil weights (8 directions + centre), LW = T + 2H.  The rows whose H-hop neighbourhood is not a lattice
 * (the own pixels within reach of the 8 valence-3 vertices of the tessellation: the fused kernels write them too,
 * wrongly) are recomputed by the generic kernels on a compact sub-problem and overwritten: `sub_plan` is L~
 * restricted to `closure_rows` (ownership passes to `plan`), `own_sub` lists the closure rows to scatter back.
 * Without an attachment (or when it does not apply) the generic per-hop kernels run. */
int ds_plan_attach_lattice(ds_plan_t* plan, int32_t n_tiles, int32_t LW, int32_t H, int32_t T, const int32_t* pix,
                           const float* w, ds_plan_t* sub_plan, int64_t n_closure, const int32_t* closure_rows,
                           int64_t n_own, const int32_t* own_sub);

/* Optional, after ds_plan_attach_lattice: the same sub-problem grouped into its connected patches, which lets the
 * tf32-mode graph convolution serve the irregular rows in ONE launch (ds_patch.cu: hops in shared memory, fp32
 * contraction, epilogue) instead of gather + K-1 hops + GEMM + scatters.  HOST arrays: row_ptr [n_patches + 1] ranges
 * of rows; rows [n_closure] row of L~ of every patch row (a permutation of closure_rows); ell_col / ell_val
 * [n_closure, 9] L~ restricted to the patch, columns patch-local, -1 = unused slot; own_ptr [n_patches + 1] ranges of
 * own_local; own_local [n_own] patch-local rows whose result is wanted.  Replaces nothing in the reference (it
 * multiplies by one tf.SparseTensor, utils.py:76); without it the generic sub-problem path runs. */
int ds_plan_attach_patches(ds_plan_t* plan, int32_t n_patches, const int32_t* row_ptr, const int32_t* rows,
                           const int32_t* ell_col, const float* ell_val, const int32_t* own_ptr,
                           const int32_t* own_local);

/* ---- utils.split_sparse_dense_matmul (utils.py:49-78) ---------------------------------
 * out[b,m,f] = alpha * sum_j L~[m,j] in[b,j,f] + beta * prev[b,m,f] + gamma * add[b,m,f]
 * (prev/add may be NULL when their factor is 0; transpose != 0 uses L~^T).  The
 * reference's n_splits column split is a TF size workaround (utils.py:59) that this
 * kernel does not need: 64-bit addressing throughout.
 */
int ds_spmm(const ds_plan_t* plan, int32_t transpose, int64_t B, int64_t F, const float* in, float alpha,
            const float* prev, float beta, const float* add, float gamma, float* out, void* stream);

/* ---- Chebyshev.call / Monomial.call (gnn_layers.py:106-161, :255-309) ------------------
 * y[B,M,Fout] = act( sum_{f,k} T_k[b,m,f] * kernel[f*K + k, o] + bias[o] )
 * with T_0 = x and the recursion above.  kernel is [K*Fin, Fout] with the reference's row
 * order f*K + k (gnn_layers.py:145-147), bias is [Fout] (the reference's [1,1,Fout]) or
 * NULL.  BatchNorm (gnn_layers.py:152-153) sits between the contraction and the bias, so
 * a use_bn layer calls this with bias = NULL, act = LINEAR and finishes with
 * ds_bias_act_forward.
 *
 * basis: workspace of (K-1)*B*M*Fin floats; on return it holds T_1..T_{K-1}
 * ([K-1, B, M, Fin]) and may be handed to ds_graph_conv_backward to skip recomputation.
 */
int64_t ds_graph_conv_basis_elems(int64_t M, int64_t B, int64_t Fin, int32_t K);
/* 1 if ds_graph_conv_forward with these arguments writes `basis` (recursion and contraction are separate
 * kernels), 0 if it runs the fused lattice kernel, which keeps the basis on chip: `basis` may then be NULL and
 * ds_graph_conv_backward must be called with basis = NULL (it recomputes what it needs). */
int32_t ds_graph_conv_forward_writes_basis(const ds_plan_t* plan, int32_t K, int64_t B, int64_t Fin, int64_t Fout,
                                           int32_t mode);
/* only the recursion of gnn_layers.py:135-143 / :287-290: basis[k-1] = T_k(L~) x for k = 1..K-1 (L~^T when
 * transpose != 0).  On a HEALPix 8-neighbour graph with a lattice attachment all K-1 hops run fused in one
 * kernel; otherwise one streaming hop kernel per k. */
int ds_graph_conv_basis(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t F, const float* x,
                        float* basis, int32_t transpose, void* stream);
int ds_graph_conv_forward(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin,
                          int64_t Fout, const float* x, const float* kernel, const float* bias, int32_t act,
                          float* y, float* basis, int32_t mode, void* stream);

/* Gradients of the above w.r.t. x, kernel, bias for an upstream gradient dy [B,M,Fout]
 * (what TF autodiff produces for gnn_layers.py:131-159).
 *  y        : forward output, needed only when act != LINEAR (derivative from y)
 *  basis    : T_1..T_{K-1} saved by the forward, or NULL to recompute them into
 *             `workspace`
 *  dx       : [B,M,Fin] or NULL;  dkernel: [K*Fin,Fout] (overwritten);  dbias: [Fout] or NULL
 *  workspace: ds_graph_conv_backward_workspace_elems(...) floats
 */
int64_t ds_graph_conv_backward_workspace_elems(int64_t M, int64_t B, int64_t Fin, int64_t Fout, int32_t K,
                                               int32_t have_basis, int32_t act);
int ds_graph_conv_backward(const ds_plan_t* plan, int32_t recursion, int32_t K, int64_t B, int64_t Fin,
                           int64_t Fout, const float* x, const float* kernel, const float* y, const float* dy,
                           int32_t act, const float* basis, float* dx, float* dkernel, float* dbias,
                           float* workspace, int32_t mode, void* stream);

/* y = act(z + bias[o]) in place-capable elementwise form, and its backward
 * (dz = dy * act'(y); dbias = sum dz).  Used when BatchNorm sits before the bias. */
int ds_bias_act_forward(int64_t R, int64_t F, const float* z, const float* bias, int32_t act, float* y, void* stream);
int ds_bias_act_backward(int64_t R, int64_t F, const float* y, const float* dy, int32_t act, float* dz,
                         float* dbias /* nullable */, float* workspace /* 2*1024*F floats */, void* stream);

/* ---- HealpyPool (healpy_layers.py:20-84) -----------------------------------------------
 * y[b,j,f] = max | mean over c < 4^p of x[b, 4^p*j + c, f];  M % 4^p == 0 required. */
int ds_pool_forward(int64_t B, int64_t M, int64_t F, int32_t p, int32_t pool_type, const float* x, float* y,
                    void* stream);
int ds_pool_backward(int64_t B, int64_t M, int64_t F, int32_t p, int32_t pool_type, const float* x,
                     const float* dy, float* dx, void* stream);

/* ---- HealpyPseudoConv (healpy_layers.py:87-146): Conv1D(Fout, 4^p, strides 4^p) --------
 * y[b,j,o] = act( sum_{c,f} x[b,4^p*j+c,f] * w[c,f,o] + bias[o] );  w is Keras' [4^p,Fin,Fout]. */
int ds_pconv_forward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                     const float* bias, int32_t act, float* y, int32_t mode, void* stream);
int64_t ds_pconv_backward_workspace_elems(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, int32_t act);
int ds_pconv_backward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                      const float* y, const float* dy, int32_t act, float* dx, float* dw, float* dbias,
                      float* workspace, int32_t mode, void* stream);

/* ---- HealpyPseudoConv_Transpose (healpy_layers.py:149-216): Conv2DTranspose (1,4^p) ----
 * y[b,4^p*j+c,o] = act( sum_f x[b,j,f] * w[0,c,o,f] + bias[o] );  w is Keras' [1,4^p,Fout,Fin]. */
int ds_pconvT_forward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                      const float* bias, int32_t act, float* y, int32_t mode, void* stream);
int64_t ds_pconvT_backward_workspace_elems(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, int32_t act);
int ds_pconvT_backward(int64_t B, int64_t M, int64_t Fin, int64_t Fout, int32_t p, const float* x, const float* w,
                       const float* y, const float* dy, int32_t act, float* dx, float* dw, float* dbias,
                       float* workspace, int32_t mode, void* stream);

/* ---- BatchNormalization + bias + activation around the contraction (gnn_layers.py:53,152-159) --------------------
 * tf.keras.layers.BatchNormalization(axis=-1, momentum, epsilon, center=False, scale=False) over the channel axis of
 * z [B, M, F], then `+ bias`, then the activation.  Statistics over rows r0 <= m < r1 of every sample (the whole
 * tensor: r0 = 0, r1 = M; a sphere-partitioned layer passes its OWN rows: the halo rows are normalised with the same
 * statistics and receive no gradient).  `sums` = double[2F] (sum z | sum z^2, or sum g | sum g*zhat in the backward):
 * a multi-GPU caller all-reduces them (and `count` = rows contributing) between the two calls of a direction, which is
 * the whole of synchronised BatchNorm.  `workspace` = ds_bn_workspace_doubles(B, M, F) doubles.
 *   forward   ds_bn_stats -> [all-reduce] -> ds_bn_bias_act_forward: updates moving_mean / moving_var [F] in place
 *             (training), writes mean_rstd [2F] (saved for the backward), scale_shift [2F] (scratch) and
 *             y = act((z - mean) * rstd + bias);  training == 0 uses the moving statistics (sums may be NULL).
 *   backward  ds_bn_backward_stats -> [all-reduce] -> ds_bn_backward_apply: dz [B, M, F] and dbias [F] (nullable). */
int64_t ds_bn_workspace_doubles(int64_t B, int64_t M, int64_t F);
int ds_bn_stats(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z, double* sums, double* workspace,
                void* stream);
int ds_bn_bias_act_forward(int64_t B, int64_t M, int64_t F, const float* z, const double* sums, double count,
                           const double* count_dev /* nullable: device scalar overriding `count` (after an all-reduce) */,
                           float eps, float momentum, int32_t training, float* moving_mean, float* moving_var,
                           const float* bias /* nullable */, int32_t act, float* mean_rstd, float* scale_shift, float* y,
                           void* stream);
int ds_bn_backward_stats(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z, const float* y,
                         const float* dy, const float* mean_rstd, int32_t act, double* sums, double* workspace,
                         void* stream);
int ds_bn_backward_apply(int64_t B, int64_t M, int64_t F, int64_t r0, int64_t r1, const float* z, const float* y,
                         const float* dy, const float* mean_rstd, const double* sums, double count,
                         const double* count_dev /* nullable */, int32_t act, int32_t training, float* dz,
                         float* dbias /* nullable */, void* stream);

/* ---- sphere partition: halo rows around the all-to-all (the reference is single-device, these are the
 * device halves of the exchange a multi-GPU binder adds around Chebyshev.call, gnn_layers.py:130-161) -----------------
 * ds_halo_pack:     send[j, b, :] = x[b, rows[j], :], j < n.  The caller concatenates the row lists of all peers, so
 *                   each peer's block of `send` is contiguous ([rows, B, F]) and one launch packs every peer.
 * ds_halo_assemble: x_ext [B, n_ext, F]: rows own_start .. own_start+n_own-1 = x_own, row pos[j] = recv[j, b, :].
 * ds_halo_reduce:   transposed exchange (backward): g_own[b, r, :] = g_ext[b, own_start + r, :] + sum over
 *                   s in slots[row_slot_ptr[r] .. row_slot_ptr[r+1]) of recv[s, b, :]   (fixed order: reproducible).
 *                   row_slot_ptr == NULL: no contributions (copy of the own slab). */
int ds_halo_pack(int64_t B, int64_t n_src, int64_t F, int64_t n, const int32_t* rows, const float* src, float* dst,
                 void* stream);
int ds_halo_assemble(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F, int64_t n_halo,
                     const int32_t* pos, const float* x_own, const float* recv, float* x_ext, void* stream);
int ds_halo_reduce(int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F, const int32_t* row_slot_ptr,
                   const int32_t* slots, const float* g_ext, const float* recv, float* g_own, void* stream);

/* ---- ds_comm_*: thin wrappers over NCCL, so that a binder without a collective library of its own can
 * run the multi-GPU path through this C-ABI alone.  NCCL is loaded at run time (ds_comm_load_nccl(path), or
 * DEEPSPHERE_NCCL_LIB, or the loader's default "libnccl.so.2"): nothing here is needed on a single GPU.
 *   ds_comm_unique_id: rank 0 creates the 128-byte id and hands it to the others out of band; ds_comm_create: every rank.
 *   ds_comm_allreduce_sum: in place, n floats (is_double = 0) or doubles (1): weight gradients, the 2F + 1 BatchNorm sums.
 *   ds_comm_alltoallv: peer blocks in rank order, counts in floats (grouped ncclSend / ncclRecv).
 *   ds_halo_exchange / ds_halo_exchange_backward: ds_halo_pack -> all-to-all -> ds_halo_assemble (resp. -> ds_halo_reduce)
 *   in ONE call; `*_rows_per_peer` [world] split the concatenated row lists; workspace = (rows sent + rows received) * B * F
 *   floats. */
int ds_comm_load_nccl(const char* path /* nullable */);
int ds_comm_unique_id(char* id128);
int ds_comm_create(int32_t world, int32_t rank, const char* id128, ds_comm_t** out);
int ds_comm_destroy(ds_comm_t* comm);
int ds_comm_allreduce_sum(ds_comm_t* comm, void* buf, int64_t n, int32_t is_double, void* stream);
int ds_comm_alltoallv(ds_comm_t* comm, const float* send, const int64_t* send_counts, float* recv,
                      const int64_t* recv_counts, void* stream);
int ds_halo_exchange(ds_comm_t* comm, int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F,
                     const int32_t* send_rows, const int64_t* send_rows_per_peer, const int32_t* recv_pos,
                     const int64_t* recv_rows_per_peer, const float* x_own, float* x_ext, float* workspace, void* stream);
int ds_halo_exchange_backward(ds_comm_t* comm, int64_t B, int64_t n_own, int64_t n_ext, int64_t own_start, int64_t F,
                              const int64_t* send_rows_per_peer, const int32_t* recv_pos,
                              const int64_t* recv_rows_per_peer, const int32_t* row_slot_ptr, const int32_t* slots,
                              const float* g_ext, float* g_own, float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSPHERE_B200_H */
