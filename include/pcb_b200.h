/*
 * pcb_b200.h -- C ABI of the B200-native batch-evaluation engine for PyChebyshev interpolants.
 *
 * The reference (0xC000005/PyChebyshev v0.21.1) has no FFI seam: its evaluation path is the
 * Python method surface of ChebyshevApproximation / ChebyshevTT / ChebyshevSpline /
 * ChebyshevSlider.  Each entry point below replaces the *body* of one of those methods; the
 * reference lines are cited per function (paths relative to src/pychebyshev/ of the reference).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add per method.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All floating point is IEEE fp64.
 *   - "plan" = immutable device-resident copy of one interpolant on one GPU, created from HOST
 *     arrays (the Python object's NumPy arrays) and destroyed with pcb_plan_destroy.
 *   - d_* arguments are DEVICE pointers owned by the caller (PyTorch allocations), `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Evaluation calls are
 *     asynchronous on `stream`, never allocate, and are re-entrant across streams.
 *   - points are (N, D) row-major fp64 in the USER dimension order; outputs are (N, G) row-major.
 *   - return 0 on success, a negative PCB_E* code otherwise; pcb_last_error() returns a
 *     thread-local message for the most recent failure on the calling thread.
 */
#ifndef PCB_B200_H
#define PCB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCB_ABI_VERSION 1

#if defined(__GNUC__)
#define PCB_API __attribute__((visibility("default")))
#else
#define PCB_API
#endif

#define PCB_OK 0
#define PCB_EINVAL (-1)   /* bad argument (maps to ValueError)            */
#define PCB_ECUDA (-2)    /* CUDA runtime failure (maps to RuntimeError)  */
#define PCB_ENOMEM (-3)   /* allocation failure                           */
#define PCB_EUNSUPPORTED (-4) /* shape outside what the kernels cover (maps to NotImplementedError) */

#define PCB_MAX_DIMS 32

PCB_API int pcb_version(void);
PCB_API const char *pcb_last_error(void);
PCB_API int pcb_device_count(void);
/* sm count, max opt-in shared memory per block, compute capability major*10+minor */
PCB_API int pcb_device_info(int dev, int *sm_count, int *smem_optin, int *cc);
PCB_API int pcb_plan_destroy(void *plan);

/* ---------------------------------------------------------------------------------------------
 * ChebyshevTT   (tensor_train.py)
 * ------------------------------------------------------------------------------------------ */

/* Build a device plan from coefficient cores.
 *   n[k], lo[k], hi[k]  : storage-frame node counts / domain (ChebyshevTT.n_nodes, .domain)
 *   ranks[0..D]         : TT ranks, ranks[0] = ranks[D] = 1 (ChebyshevTT.tt_ranks)
 *   dim_order[k]        : user column stored at TT position k (ChebyshevTT._dim_order)
 *   cores_cat           : _coeff_cores[k] (r_{k-1}, n_k, r_k) C-order, concatenated
 */
PCB_API int pcb_tt_plan_create(int dev, int D, const int32_t *n, const int32_t *ranks, const double *lo,
                       const double *hi, const int32_t *dim_order, const double *cores_cat,
                       void **plan);

/* Replaces ChebyshevTT.eval_batch (tensor_train.py:2217-2265): d_out[i] = value at point i. */
PCB_API int pcb_tt_eval(void *plan, const double *d_points, int64_t N, double *d_out, void *stream);

/* Replaces a loop of ChebyshevTT.eval_multi over points (tensor_train.py:2267-2463): value and
 * central finite differences with h = (b-a)*1e-4 and the boundary nudge, G derivative-order
 * rows (user frame, G x D int32, HOST pointer) -> d_out (N, G).  Orders > 2 -> PCB_EINVAL
 * ("Derivative order k not supported (use 1 or 2)").  `algo`: 0 = auto, 1 = one chain per stencil
 * point (any orders, <= 3 active dims per row), 2 = shared left/right partial products (rows with
 * <= 1 active dim).  One call takes at most 16 rows, each with at most 3 differentiated dims; the
 * host layer (tt.py) splits larger row sets into several calls and unrolls deeper nested stencils
 * into value launches, so every row set the reference's eval_multi accepts is served. */
PCB_API int pcb_tt_eval_fd(void *plan, const double *d_points, int64_t N, int G, const int32_t *orders,
                   double *d_out, int algo, void *stream);

/* How the plan will be evaluated (reporting only): out8 = { values on the uniform-datapath
 * (constant bank) kernels?, shared-FD on them?, query slots per thread, threads (values), threads
 * (shared-FD), core placement of the shared-memory kernels (0 resident, 1 streamed, 2 global),
 * their slots per thread, their threads }. */
PCB_API int pcb_tt_plan_info(void *plan, int32_t *out8);

/* Which algorithm pcb_tt_eval_fd(algo = 0) runs for these rows: 1 or 2; negative PCB_E* on error. */
PCB_API int pcb_tt_fd_algo(void *plan, int G, const int32_t *orders);

/* Which kernel family pcb_tt_eval_fd(algo) runs for these rows (reporting only -- evaluation keeps
 * no state in the plan): 1 one chain per stencil point, 2 constant-bank shared-product kernel,
 * 3 the same with one launch per differentiated dim, 4 one launch per core (large trains),
 * 5 shared-memory shared-product kernel; negative PCB_E* on error. */
PCB_API int pcb_tt_fd_path(void *plan, int G, const int32_t *orders, int algo);

/* ---------------------------------------------------------------------------------------------
 * ChebyshevApproximation   (barycentric.py)
 * ------------------------------------------------------------------------------------------ */

/* Build a device plan for G pre-differentiated tensors of one interpolant.
 *   n[d]                 : ChebyshevApproximation.n_nodes
 *   nodes_cat/weights_cat: .nodes[d] / .weights[d] concatenated over d
 *   tensors_host[g]      : C-order tensor after _apply_derivative_passes(tensor_values, order_g)
 *                          (barycentric.py:951-990; made on the host with the reference's recipe)
 */
PCB_API int pcb_full_plan_create(int dev, int D, const int32_t *n, const double *nodes_cat,
                         const double *weights_cat, int G, const double *const *tensors_host,
                         void **plan);

/* The same plan from the VALUE tensor alone: the tensor is uploaded once and
 * the G derivative tensors are made ON THE DEVICE with the passes of _apply_derivative_passes
 * (barycentric.py:982-989: for d = D-1..0, orders[g][d] times  T <- T x_d D_d^T), each output
 * element one sequential FMA chain over k -- bit-identical to the reference's NumPy/OpenBLAS recipe
 * on FMA-capable hosts.
 *   diffmats_cat : .diff_matrices[d] (n_d x n_d, C-order) concatenated over d
 *   values_host  : .tensor_values (C-order)
 *   orders       : G x D derivative orders (HOST pointer) */
PCB_API int pcb_full_plan_create_from_values(int dev, int D, const int32_t *n, const double *nodes_cat,
                                     const double *weights_cat, const double *diffmats_cat,
                                     const double *values_host, int G, const int32_t *orders,
                                     void **plan);

/* Replaces G calls of ChebyshevApproximation.vectorized_eval_batch (barycentric.py:992-1047),
 * one per pre-differentiated tensor: d_out (N, G).  `algo`: 0 = auto, 1 = thread-per-query FMA
 * evaluator, 2 = DMMA mode-1 GEMM with fused tail (D >= 2), 3 = DMMA GEMM over the last TWO axes
 * jointly (D >= 3; chosen by auto when the last axis is not a multiple of 4). */
PCB_API int pcb_full_eval(void *plan, const double *d_points, int64_t N, double *d_out, int algo,
                  void *stream);

/* ---------------------------------------------------------------------------------------------
 * Mode contractions of a C-order tensor in DEVICE memory
 * ------------------------------------------------------------------------------------------ */

/* One pass of _apply_derivative_passes (barycentric.py:982-989): d_dst = d_src x_axis D^T, same
 * shape, out of place.  dmat_host: the n[axis] x n[axis] differentiation matrix (HOST, C-order). */
PCB_API int pcb_tensor_deriv(int dev, int D, const int32_t *n, int axis, const double *dmat_host,
                     const double *d_src, double *d_dst, void *stream);

/* _slice_tensor (_extrude_slice.py:79-92): d_dst = tensordot(d_src, vec, axes=([axis],[0])); the
 * axis is removed.  vec_host: n[axis] weights (HOST) -- the normalised barycentric weights of the
 * slicing value, or a one-hot row on a node hit. */
PCB_API int pcb_tensor_contract(int dev, int D, const int32_t *n, int axis, const double *vec_host,
                        const double *d_src, double *d_dst, void *stream);

/* _extrude_tensor (_extrude_slice.py:73-76): insert a new axis of length n_new at position `axis`
 * of a D-dim tensor of shape n and replicate. */
PCB_API int pcb_tensor_extrude(int dev, int D, const int32_t *n, int axis, int n_new, const double *d_src,
                       double *d_dst, void *stream);

/* ---------------------------------------------------------------------------------------------
 * ChebyshevSpline   (spline.py)
 * ------------------------------------------------------------------------------------------ */

/* Build a device plan for a spline with P = prod(num_knots[d] + 1) pieces (C-order).
 *   piece_n            : P x D node counts (nested n_nodes allowed)
 *   piece_nodes_cat    : for piece p, for dim d: nodes (piece_n[p][d] doubles), concatenated
 *   piece_weights_cat  : same layout, barycentric weights
 *   piece_tensors_host : P x G pointers, piece-major; each a C-order pre-differentiated tensor
 */
PCB_API int pcb_spline_plan_create(int dev, int D, const int32_t *num_knots, const double *knots_cat, int P,
                           const int32_t *piece_n, const double *piece_nodes_cat,
                           const double *piece_weights_cat, int G,
                           const double *const *piece_tensors_host, void **plan);

/* Replaces the routing lines of ChebyshevSpline.eval_batch (spline.py:677-690; single-point twin
 * _find_piece :414-445): d_piece[i] = C-order flat piece index, bit-exact integer work.
 * d_points needs only the natural 8-byte alignment; a 16-byte aligned base additionally enables
 * 128-bit loads for 2-D splines. */
PCB_API int pcb_spline_lookup(void *plan, const double *d_points, int64_t N, int32_t *d_piece, void *stream);

/* Replaces ChebyshevSpline.eval_batch (spline.py:633-700) for G derivative tensors: d_out (N, G).
 * d_piece may be NULL (lookup fused) or receive the piece indices. */
PCB_API int pcb_spline_eval(void *plan, const double *d_points, int64_t N, double *d_out, int32_t *d_piece,
                    void *stream);

/* ---------------------------------------------------------------------------------------------
 * ChebyshevSlider   (slider.py)
 * ------------------------------------------------------------------------------------------ */

/* Build a device plan for an additive slider.
 *   S slides; group_size[s] dims each; group_dims_cat = partition flattened
 *   slide_n_cat / slide_nodes_cat / slide_weights_cat : per slide, per group dim
 *   G outputs; out_slide[g] = -1 -> value row: pivot + sum_s (slide_s - pivot) (slider.py:310-318)
 *                           = -2 -> cross-slide mixed partial: exactly 0.0      (slider.py:297-298)
 *                           = s  -> derivative owned by slide s                  (slider.py:301-307)
 *   slide_tensors_host : G x S pointers (row g, slide s); only the entries a row uses are read
 */
PCB_API int pcb_slider_plan_create(int dev, int D, int S, const int32_t *group_size,
                           const int32_t *group_dims_cat, const int32_t *slide_n_cat,
                           const double *slide_nodes_cat, const double *slide_weights_cat,
                           double pivot_value, int G, const int32_t *out_slide,
                           const double *const *slide_tensors_host, void **plan);

/* Replaces a loop of ChebyshevSlider.eval over points (slider.py:247-318): d_out (N, G). */
PCB_API int pcb_slider_eval(void *plan, const double *d_points, int64_t N, double *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Native .pcb loader (no Python on the path): the C++ twin of the reference's stand-alone readers
 * (examples/binary_reader/reader.c, readers/rust/src/lib.rs, readers/julia/src/PCBReader.jl).
 * Parses the v1 layout of _binary.py:157-421, rebuilds nodes/weights the way the reference does on
 * load, and creates a VALUE plan (one output per point).  *kind: 1 approximation, 2 spline.
 * ------------------------------------------------------------------------------------------ */
PCB_API int pcb_plan_from_file(int dev, const char *path, void **plan, int *kind, int *D);

/* The same with G derivative-order rows (G x D int32, HOST): a plan with G outputs per point.  The
 * differentiation matrices follow barycentric.py:52-77 (row sums in NumPy's pairwise order); the
 * derivative tensors are made by the passes of barycentric.py:982-989 -- on the device for
 * approximations (pcb_full_plan_create_from_values), on the host for the small spline pieces.
 * G = 0 / orders = NULL: values only. */
PCB_API int pcb_plan_from_file_orders(int dev, const char *path, int G, const int32_t *orders, void **plan,
                              int *kind, int *D);

/* Native .pcb WRITER (reference _binary.py:208-236, 289-346): byte-identical files.
 *   pcb_file_write_approx : ChebyshevApproximation (domain lo/hi, n_nodes, C-order tensor)
 *   pcb_file_write_spline : ChebyshevSpline with flat n_nodes (knots concatenated over dims,
 *                           P = prod(num_knots + 1) piece tensors in C-order)
 *   pcb_file_rewrite      : read in_path with the native parser and write it back (round trip) */
PCB_API int pcb_file_write_approx(const char *path, int D, const double *lo, const double *hi,
                          const int32_t *n, const double *tensor);
PCB_API int pcb_file_write_spline(const char *path, int D, const double *lo, const double *hi,
                          const int32_t *n, const int32_t *num_knots, const double *knots_cat, int P,
                          const double *const *piece_tensors);
PCB_API int pcb_file_rewrite(const char *in_path, const char *out_path);

/* Nodes, barycentric weights and (optionally, dmat != NULL) the n x n differentiation matrix the
 * native loader derives for one dimension (reporting / tests). */
PCB_API int pcb_file_grid_arrays(double lo, double hi, int n, double *nodes, double *weights, double *dmat);

/* ---------------------------------------------------------------------------------------------
 * Optional result gather for query-sharded evaluation (reference: none -- the
 * reference is single-process, its batch methods return one (N,) array, barycentric.py:992-1047).
 * One process per GPU.  Each rank owns a replicated result tensor (world x rows x G doubles),
 * exported as a CUDA IPC handle; the evaluators write this rank's slice of its own replica and
 * pcb_peer_push copies that slice into every peer's replica with the copy engines over NVLink
 * (one internal stream per peer), ordered after everything already enqueued on `stream`; `stream`
 * itself does not wait, so the next chunk's kernel overlaps the copies.  pcb_peer_join makes
 * `stream` wait for every push issued so far on that device (call it before the source slice is
 * overwritten, and before the cross-process barrier that publishes the replicas).  No NCCL, no kernel.
 *   pcb_peer_alloc : cudaMalloc + cudaIpcGetMemHandle (handle: PCB_PEER_HANDLE_BYTES bytes to send
 *                    to the other processes by any host-side means)
 *   pcb_peer_open  : map a peer's allocation into this process (enables peer access lazily)
 *   pcb_peer_push  : d_src[0:bytes] -> peer_ptrs[i] + offset_bytes for i < n_peers (<= 16)
 *   pcb_peer_join  : `stream` waits for all outstanding pushes of device `dev`
 * ------------------------------------------------------------------------------------------ */
#define PCB_PEER_HANDLE_BYTES 64
PCB_API int pcb_peer_alloc(int dev, uint64_t bytes, void **d_ptr, unsigned char *handle);
PCB_API int pcb_peer_free(int dev, void *d_ptr);
PCB_API int pcb_peer_open(int dev, const unsigned char *handle, void **d_ptr);
PCB_API int pcb_peer_close(int dev, void *d_ptr);
PCB_API int pcb_peer_push(int dev, int n_peers, void *const *peer_ptrs, uint64_t offset_bytes,
                          const void *d_src, uint64_t bytes, void *stream);
PCB_API int pcb_peer_join(int dev, void *stream);

/* Values of any plan kind (the plan's own number of outputs per point). */
PCB_API int pcb_plan_eval(void *plan, const double *d_points, int64_t N, double *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Roofline probes (bench.py): measured FP64 pipe peaks of the device, in TFLOP/s.
 *   kind 0 = DFMA (register-resident FMA chains), 1 = DMMA (mma.sync m8n8k4 f64),
 *   kind 2 = both interleaved in every warp (sum of the two flop counts)
 * ------------------------------------------------------------------------------------------ */
PCB_API int pcb_probe_fp64_peak(int dev, int kind, double *tflops, double *ms);

/* Number of kernels this library has launched on the calling process (bench.py gpu_launches). */
PCB_API int64_t pcb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PCB_B200_H */
