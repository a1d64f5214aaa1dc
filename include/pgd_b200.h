/* pgd_b200.h -- C ABI of libpgdb200.so: the B200 (sm_100a) kernels behind the progressive PGD
 * enrichment hot path of BAMresearch/PGDrome (pgdrome/solver.py:306-881, pgdrome/model.py:724-860).
 *
 * The reference has no FFI of its own (pure Python on top of DOLFIN/PETSc/SciPy); each entry
 * point below names the reference call site whose arithmetic it replaces.  A reference
 * maintainer binds these with ctypes.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (e.g. torch.Tensor.data_ptr());
 *     h_* is HOST memory.  The library owns only the opaque handle and its scratch.
 *   - fp64 values, int32 indices (int64 where stated), row-major, CSR with ascending columns.
 *   - return 0 = OK, >0 = cudaError_t, <0 = argument error; pgd_last_error(h) gives the text.
 *   - asynchronous on `stream` (a cudaStream_t passed as void*) unless the name ends in _sync.
 *   - one host thread and one stream at a time per handle; one handle per GPU / rank.
 *   - every result is deterministic and bitwise reproducible: reductions are fixed-order two-stage sums, there are no
 *     floating-point atomics anywhere (the lifting of non-zero Dirichlet values is one SpMV).
 */
#ifndef PGD_B200_H
#define PGD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgd_ctx* pgd_handle_t;

int32_t pgd_abi_version(void);
int32_t pgd_create(int32_t device, pgd_handle_t* out);
int32_t pgd_destroy(pgd_handle_t h);
const char* pgd_last_error(pgd_handle_t h);
/* options: "pcg_resident" (default 1) = solve with the single-kernel SM-resident PCG whenever the
 * matrix slice of every SM fits in its shared memory; "persist" (default 1) = larger systems (>= 32 768 rows) run in the
 * persistent streaming kernel of pgd_pcg_persist_sync (0: three launches per iteration); "bsr" (default 2) = node-block
 * walk inside that kernel when a block-column list is supplied (2: direct walk, coalesced loads of the row segments, no
 * shared memory and no CTA barrier; 1: tiles of block rows through the TMA ring; 0: plain CSR); "single_reduction" (default 0) = 0 never / 1 on sharded
 * systems / 2 always use the Chronopoulos-Gear form of the iteration (one grid-wide reduction per step instead of two: same
 * Krylov method, A p carried by a recurrence; measured slower on this hardware at 1, 2 and 8 GPUs -- two more vector
 * streams per iteration cost more than the reduction saves -- hence off); "ll" (default 1) = cross-GPU
 * sums and halo entries of that kernel travel as 8-byte words carrying 32 data bits + the 32-bit sequence number (valid on
 * arrival: no system-scope fence, no flag; 0 = data, fence, release flag); "spin_ms" = budget
 * of every in-kernel wait; "prof" = per-phase timers of that kernel (pgd_get_phase_ns);
 * "p2p", "graph", "fused", "pcg3", "spmv_stream": variants of the older multi-launch paths. */
int32_t pgd_set_option(pgd_handle_t h, const char* name, int64_t value);
/* library-side counters since the last reset: h_counts[0] kernels launched, [1] PCG solves,
 * [2] PCG iterations, [3] solves done by the SM-resident kernel; *h_pcg_ms device time (CUDA events on the solve's stream) spent in the PCG
 * iteration kernels.  Used by bench.py for gpu_launches and the live roofline. */
int32_t pgd_get_stats(pgd_handle_t h, int64_t* h_counts, double* h_pcg_ms, int32_t reset);

/* ---- sparsity pattern (DOLFIN SparsityPatternBuilder behind solver.py:627-636: union of per-cell
 * dof cliques, columns ascending).  build: sorts the n_cells*ndl^2 (row,col) contributions; the
 * sorted order is also the deterministic gather list used by pgd_gather_values.
 * export: rowptr[n_dofs+1], colidx[nnz], gptr[nnz+1] (int64), gidx[n_cells*ndl*ndl] (int32). */
int32_t pgd_pattern_build_sync(pgd_handle_t h, const int32_t* d_cell_dofs, int64_t n_cells, int32_t ndl,
                               int64_t n_dofs, int64_t* h_nnz, void* stream);
int32_t pgd_pattern_export(pgd_handle_t h, int32_t* d_rowptr, int32_t* d_colidx, int64_t* d_gptr,
                           int32_t* d_gidx, void* stream);
/* dof -> contribution lists for load vectors: vptr[n_dofs+1] (int64), vidx[n_cells*ndl] (int32) */
int32_t pgd_vecmap_build_sync(pgd_handle_t h, const int32_t* d_cell_dofs, int64_t n_cells, int32_t ndl,
                              int64_t n_dofs, int64_t* d_vptr, int32_t* d_vidx, void* stream);

/* ---- separated-form assembly (dolfin.assemble / SystemAssembler + FFC tabulate_tensor behind
 * every lhs_fct/rhs_fct callback, e.g. tests/integration/test_elastic.py:71-219).
 * Generic Lagrange element kernel: element tables phi[nq,nd], dphi[nq,nd,tdim], qw[nq] are staged
 * in shared memory; T[bs,gdim+1,bs,gdim+1] is the constant form tensor (slot 0 = value, 1+m = d/dx_m;
 * first index pair = test, second = trial); wq[n_cells,nq] optional coefficient at quadrature points.
 * Output Ae[n_cells, nd*bs, nd*bs] (local dof = node*bs+comp), reduced by pgd_gather_values. */
int32_t pgd_elem_bilinear(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                          int32_t tdim, int32_t gdim, int32_t bs, int32_t nd, int32_t nq, const double* d_phi,
                          const double* d_dphi, const double* d_qw, const double* d_wq, const double* d_T,
                          double* d_Ae, void* stream);
/* be[n_cells, nd*bs] = int w * sum L[i,j] D_j v_i ; L[bs,gdim+1] */
int32_t pgd_elem_linear(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                        int32_t tdim, int32_t gdim, int32_t bs, int32_t nd, int32_t nq, const double* d_phi,
                        const double* d_dphi, const double* d_qw, const double* d_wq, const double* d_L,
                        double* d_be, void* stream);
/* out[g] = sum_{k in [gptr[g],gptr[g+1])} src[gidx[k]]  (fixed order => bitwise reproducible) */
int32_t pgd_gather_values(pgd_handle_t h, const double* d_src, const int64_t* d_gptr, const int32_t* d_gidx,
                          int64_t n_out, double* d_out, void* stream);
/* Fused P1 simplex operator  values = c_mass*M + c_stiff*K + sum_m c_adv[m]*int (d_m u) v  written
 * straight into the CSR pattern (one kernel, no element buffer): warp-segmented reduction over the
 * gather list.  gdim = tdim in {1,2,3}; c_adv may be NULL. */
int32_t pgd_assemble_p1(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                        int32_t gdim, double c_mass, double c_stiff, const double* h_c_adv,
                        const int64_t* d_gptr, const int32_t* d_gidx, int64_t nnz, double* d_values, void* stream);

/* Row-owner variant of the fused P1 operator (same arithmetic, ~20x faster on 3-D meshes): a thread
 * owns one row / mesh node and walks the cells around it in vecmap order.  The plan d_vent
 * (int32 [2 * n_cells * nv]: pairs {vidx entry, nv packed byte positions inside the CSR row}) is built
 * once per mesh from the pattern, the cell -> dof table (local dof a = local vertex a; the dof numbering
 * need not equal the vertex numbering) and the vecmap of the scalar P1 space;
 * returns -4 if a row has more than 255 entries (use pgd_assemble_p1 then).
 * d_coords_soa (optional, else NULL): the same coordinates component-major [gdim][n_verts]; with it a
 * warp-wide coordinate load touches 2 cache lines instead of 6-8 (d_coords may then be NULL). */
int32_t pgd_p1_rowplan_build_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx,
                                  const int32_t* d_cell_dofs, int64_t n_cells, int32_t nv, const int64_t* d_vptr,
                                  const int32_t* d_vidx, int64_t n_nodes, int32_t* d_vent, void* stream);
int32_t pgd_assemble_p1_rows(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                             int32_t gdim, double c_mass, double c_stiff, const double* h_c_adv,
                             const int32_t* d_rowptr, const int64_t* d_vptr, const int32_t* d_vent, int64_t n_nodes,
                             double* d_values, const double* d_coords_soa, int64_t n_verts, void* stream);

/* Row-owner assembly of ANY constant-coefficient P1 atom (form tensor h_T [bs][gdim+1][bs][gdim+1]: mass, stiffness, first
 * derivatives, Voigt elasticity, ...) on a scalar or node-blocked vector P1 space, optional per-cell coefficient d_w_cell
 * [n_cells] (degree-0 Expression), straight into the CSR pattern: one thread per dof row, no element-matrix buffer and no
 * gather pass.  The plan (d_node_vptr int64 [n_nodes+1], d_node_vent int32 pairs) is pgd_vecmap_build_sync +
 * pgd_p1_rowplan_build_sync on the NODE-level pattern (block rows / block columns; for bs = 1 the dof-level plan);
 * d_coords_soa [gdim][n_verts] component-major; n_rows = n_nodes * bs.  Deterministic. */
int32_t pgd_assemble_p1_tensor(pgd_handle_t h, int32_t gdim, int32_t bs, const double* h_T, const double* d_coords_soa,
                               int64_t n_verts, const int32_t* d_cell_verts, const int32_t* d_rowptr,
                               const int64_t* d_node_vptr, const int32_t* d_node_vent, const double* d_w_cell, int64_t n_rows,
                               double* d_values, void* stream);

/* ---- linear combinations: A_d = sum_k c_k K_{d,k} over CSR value arrays, and
 * b_d = sum c_m g_m - sum c_ik (K_k U_i) over cached vectors (the folded scalar coefficients of
 * SURVEY.md 7.1).  h_xs: HOST array of n_terms device pointers; h_coefs: HOST array. */
int32_t pgd_lincomb(pgd_handle_t h, int32_t n_terms, const double* const* h_xs, const double* h_coefs,
                    int64_t n, double* d_out, int32_t accumulate, void* stream);
/* Same with the coefficients in device memory (d_coefs[n_terms]), e.g. produced by pgd_scalar_programs. */
int32_t pgd_lincomb_dev(pgd_handle_t h, int32_t n_terms, const double* const* h_xs, const double* d_coefs, int64_t n,
                        double* d_out, int32_t accumulate, void* stream);
/* Evaluate up to 32 small arithmetic expressions over device-resident scalars: program g is the postfix code
 * h_code[h_off[g] .. h_off[g+1]) with instructions (opcode << 24) | operand; opcodes 0 = push h_consts[operand],
 * 1 = push d_pool[operand], 2 = mul, 3 = add, 4 = sub, 5 = div, 6 = negate; d_out[g] = the value left on the stack.
 * Each operation is one IEEE-754 double operation in program order (no contraction): the result is bitwise what the
 * host obtains for the same expression.  Limits: 640 instructions, 128 constants, stack depth 16 per call.  The code
 * travels as kernel parameters (no copy); used for the coefficients c_k of A_d = sum_k c_k K_{d,k}
 * (solver.py:598-621), which are products of mode integrals (pgd_bilinear / pgd_panel_dots results). */
int32_t pgd_scalar_programs(pgd_handle_t h, int32_t n_prog, const int32_t* h_off, const int32_t* h_code,
                            const double* h_consts, int32_t n_consts, const double* d_pool, double* d_out, void* stream);

/* ---- Dirichlet (DirichletBC.apply / assemble_system symmetric elimination, solver.py:186-191,
 * 364-372,704-716): zero row+col of every bc dof, diagonal 1, b lifted and set.  d_bc_vals may be
 * NULL (all reference BC values are 0).  Pattern must be structurally symmetric. d_b may be NULL.
 * Non-zero values: b -= A g is formed first as one SpMV with the un-eliminated operator (deterministic); this needs
 * n_rows (rows of the pattern) and d_work of 2 * n_rows doubles -- both ignored when d_bc_vals is NULL. */
int32_t pgd_apply_dirichlet(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, double* d_values,
                            double* d_b, const int32_t* d_bc_dofs, const double* d_bc_vals, int64_t n_bc,
                            int64_t n_rows, double* d_work, void* stream);
int32_t pgd_set_entries(pgd_handle_t h, double* d_x, const int32_t* d_idx, const double* d_vals, int64_t n_idx,
                        void* stream);

/* ---- SpMV and mode integrals (dolfin.assemble of scalar functionals, dolfin.norm:
 * solver.py:207,342,754,836-842 and the Constant(assemble(..)) factors of every callback).
 * lanes_per_row: 0 = auto, else 2|4|8|16|32. */
int32_t pgd_spmv(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                 const double* d_x, double* d_y, int64_t n_rows, int32_t lanes_per_row, void* stream);
/* y = A x and *d_dot = w^T y in the same pass (w may alias x) */
int32_t pgd_spmv_dot(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                     const double* d_x, double* d_y, const double* d_w, double* d_dot, int64_t n_rows,
                     int32_t lanes_per_row, void* stream);
/* *d_out = x^T A y  (fused SpMV-reduction, nothing written but the scalar) */
int32_t pgd_bilinear(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                     const double* d_x, const double* d_y, int64_t n_rows, double* d_out, int32_t lanes_per_row,
                     void* stream);
int32_t pgd_dot(pgd_handle_t h, const double* d_x, const double* d_y, int64_t n, double* d_out, void* stream);
/* d_out[m] = P[m,:] . x   for m < n_vecs, P row-major with leading dimension ld (cached panels K U) */
int32_t pgd_panel_dots(pgd_handle_t h, const double* d_P, int64_t ld, int32_t n_vecs, const double* d_x,
                       int64_t n, double* d_out, void* stream);

/* ---- linear solves (LinearVariationalSolver / NonlinearVariationalSolver + MUMPS LU,
 * solver.py:579-595,627-636,651-674,704-716; FD_solve spsolve solver.py:927-943).
 * Jacobi-PCG, x0 = 0, stop ||r|| <= max(rtol*||b||, atol); convergence flag and iteration counter
 * live on the device, the host only polls them every check_every iterations.
 * d_work: (5+block)*n + 8 doubles.  block = 1 (point Jacobi) or bs (bs x bs node-block Jacobi, bs <= 3). */
int32_t pgd_pcg_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                     const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                     int32_t check_every, int32_t block, int32_t lanes_per_row, double* d_work,
                     int32_t* h_iters, double* h_relres, void* stream);
/* ---- sharded PCG building blocks (spatial mesh partitioned by rows over the GPUs of one box; no
 * counterpart in the serial reference).  Local vectors are laid out [owned | ghost], the local CSR
 * (n_owned rows) uses that numbering.  One iteration on every rank:
 *   pgd_spcg_direction -> [halo exchange of p] -> pgd_spcg_matvec -> [allreduce d_sc[2]] ->
 *   pgd_spcg_update -> [allreduce d_sc[8..9]] -> pgd_spcg_rotate
 * with the collectives issued by the host (NCCL) on the same stream.  No call synchronises or
 * allocates, so an iteration can be captured in a CUDA graph.  Caller-owned device state:
 *   d_sc (>= 16 doubles): [0] r.z old, [1] r.z, [2] p.q, [3] r.r, [4] b.b, [5] tol^2, [8..9] partial sums
 *   d_fl (>= 4 int32):    [0] done, [1] iterations, [2] NaN seen
 *   d_work: n_owned*(3+block) + n_local + 8 doubles, laid out r, z, q (stride ns = n_owned rounded up to even),
 *           M^-1, then p with its ghost tail at d_work + 3*ns + even(n_owned*block); d_x [n_owned] is the solution.
 * pgd_spcg_init leaves the local (r.z, b.b, r.r) in d_sc[8..10]: all-reduce them, then pgd_spcg_init_fin. */
int32_t pgd_spcg_init(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                      const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block, double* d_work,
                      double* d_sc, int32_t* d_fl, void* stream);
int32_t pgd_spcg_init_fin(pgd_handle_t h, double* d_sc, int32_t* d_fl, double rtol, double atol, void* stream);
int32_t pgd_spcg_direction(pgd_handle_t h, double* d_work, int64_t n_owned, int32_t block, const double* d_sc,
                           const int32_t* d_fl, void* stream);
int32_t pgd_spcg_matvec(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                        double* d_work, int64_t n_owned, int32_t block, double* d_sc, void* stream);
int32_t pgd_spcg_update(pgd_handle_t h, double* d_x, double* d_work, int64_t n_owned, int32_t block, double* d_sc,
                        const int32_t* d_fl, void* stream);
int32_t pgd_spcg_rotate(pgd_handle_t h, double* d_sc, int32_t* d_fl, void* stream);

/* The whole sharded solve in one call: the loop above with NCCL send/recv for the halo of p (ghosts land in
 * p's tail, grouped by source rank) and ncclAllReduce for the three dot products, all enqueued on
 * `stream`; the host polls the convergence flag every check_every iterations.  Communicator: rank 0 calls
 * pgd_comm_unique_id (128 bytes, NCCL is dlopen'ed at run time), the host distributes it (e.g.
 * torch.distributed.broadcast) and every rank calls pgd_comm_init on its handle.  Without a communicator
 * the call is the single-rank solve.  d_send_idx: int64 local owned indices grouped by destination
 * rank; h_send_counts / h_recv_counts: [world] host arrays.
 * d_work: n_owned*(3+block) + n_local + sum(send_counts) + 8 doubles. */
int32_t pgd_comm_unique_id(void* h_id128);
int32_t pgd_comm_init(pgd_handle_t h, const void* h_id128, int32_t rank, int32_t world);
int32_t pgd_comm_destroy(pgd_handle_t h);
int32_t pgd_spcg_solve_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                            const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block,
                            const int64_t* d_send_idx, const int64_t* h_send_counts, const int64_t* h_recv_counts,
                            double rtol, double atol, int32_t maxit, int32_t check_every, double* d_work,
                            int32_t* h_iters, double* h_relres, void* stream, const int64_t* h_peer_ghost_base);
/* NVLink peer window (optional, replaces NCCL inside the iteration): every rank creates a window
 * (cudaMalloc + CUDA IPC handle, p_capacity >= the largest n_local of all ranks, same value everywhere),
 * the host all-gathers the 64-byte handles and every rank opens them.  With a window and
 * h_peer_ghost_base[r] = offset (in doubles) inside rank r's p where this rank's block of ghosts starts
 * (= n_owned_r + sum of r's recv_counts from lower ranks), pgd_spcg_solve_sync keeps p inside the window:
 * neighbours store their boundary values of p straight into the ghost tail over NVLink and publish a
 * sequence flag (st.release.sys); the dot products are reduced by a one-shot mailbox all-reduce summed in
 * rank order (bitwise identical on all ranks).  pgd_set_option("p2p", 0) falls back to NCCL.  Returns -6 if
 * a peer does not arrive within the "spin_ms" budget (default 20 s) instead of hanging the GPU; -6 is fatal for the
 * WINDOW (the ranks' sequence numbers may have diverged): it is disabled and later solves run over NCCL until a new
 * window is created and opened collectively.  pgd_peer_window_create on a handle that already has a window closes
 * the old peer mappings and frees it first (put a barrier across the ranks in front). */
int32_t pgd_peer_window_create(pgd_handle_t h, int64_t p_capacity, void* h_ipc64);
int32_t pgd_peer_window_open(pgd_handle_t h, int32_t rank, int32_t world, const void* h_all_ipc);
int32_t pgd_peer_window_destroy(pgd_handle_t h);

/* Same solve with a warm start: on entry d_x holds the initial guess x0 (r0 = b - A x0); the stopping
 * rule stays relative to ||b||.  The fixed-point sweeps of solver.py:531-757 solve a sequence of nearby
 * systems, so the previous sweep's mode is passed as x0 (the reference's direct LU has no such notion). */
int32_t pgd_pcg_x0_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                        const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                        int32_t check_every, int32_t block, int32_t lanes_per_row, double* d_work,
                        int32_t* h_iters, double* h_relres, void* stream);

/* Persistent PCG for systems that do not fit on chip (the HBM-bound regime of configs[2]-[3]): ONE cooperative kernel
 * per solve -- direction update, TMA-pipelined SpMV, vector update and all reductions inside, CG scalars in registers,
 * stop test on the device -- on one GPU or on every GPU of a row-sharded system (same settings-forwarding call site as
 * pgd_pcg_sync: solver.py:634-635).  Local layout as for the sharded blocks: d_rowptr has n_owned rows, columns and
 * d_x [n_local] in [owned | ghost] numbering (single GPU: n_local = n_owned, the four halo arguments NULL).
 *   warm != 0: d_x holds the initial guess INCLUDING valid ghost entries; on return d_x holds the solution and, sharded,
 *   its ghost entries are up to date as well (exchanged through the peer window inside the kernel).
 *   Sharded: needs an opened peer window (pgd_peer_window_create with p_capacity >= 6 * max n_local + 16 over the ranks:
 *   the vector z at [0, p_capacity / 3) and the LL halo region behind it; with pgd_set_option "ll" = 0,
 *   2 * max n_local + 8 suffices) and the halo description of pgd_spcg_solve_sync.  Neighbour values of the
 *   preconditioned residual travel straight into the peers' windows over NVLink, the dot products are summed through
 *   per-rank mailboxes in rank order (bitwise identical on all ranks).  A window that is too small for the LL halo is an
 *   argument error (-1), not a silent change of protocol.  Every in-kernel wait has a wall-clock budget (pgd_set_option "spin_ms", default 20 000): a missing peer
 *   ends the call with -6 instead of hanging, and the peer window is disabled afterwards (the ranks' sequence numbers
 *   may have diverged; pgd_spcg_solve_sync over NCCL keeps working).
 *   d_bcol / max_blocks_per_row (optional, block > 1): block-column list of the node-block walk -- d_bcol[j] = node of the
 *   j-th block x block block, block rows in order (nnz / block^2 entries); the kernel then streams 8 B per nonzero plus
 *   4 B per block instead of 12 B per nonzero.  Requires a block-structured pattern (every node pair coupled by a full
 *   block, as produced by pgd_pattern_build_sync on a node-blocked vector space).
 *   d_work: 3 * even(n_owned) + even(n_owned * block) + 2 * even(n_local) + 8 doubles, 16-byte aligned. */
int32_t pgd_pcg_persist_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                             const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block, double rtol,
                             double atol, int32_t maxit, int32_t warm, double* d_work, const int64_t* d_send_idx,
                             const int64_t* h_send_counts, const int64_t* h_recv_counts, const int64_t* h_peer_ghost_base,
                             const int32_t* d_bcol, int32_t max_blocks_per_row, int32_t* h_iters, double* h_relres,
                             void* stream);

/* Phase profile of that kernel (pgd_set_option "prof" = 1): accumulated nanoseconds as seen by CTA 0 in
 * [0] direction update + halo push, [1] grid barrier + halo wait, [2] SpMV, [3] reduce-broadcast of p.q,
 * [4] vector update, [5] reduce-broadcast of r.z / r.r.  Diagnostic (tools/pcg_bench.py); synchronises. */
int32_t pgd_get_phase_ns(pgd_handle_t h, int64_t* h_ns6, int32_t reset);

/* General banded LU with partial pivoting, one CTA, for the 1-D parameter / time dimensions
 * (tiny, possibly non-symmetric).  d_perm[new] = old dof (band ordering), kl/ku bandwidths in the
 * permuted numbering; d_work: (2*kl+ku+1)*n + n doubles; *d_info != 0 on a zero pivot. */
int32_t pgd_banded_solve(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                         const double* d_b, double* d_x, int32_t n, const int32_t* d_perm, int32_t kl, int32_t ku,
                         double* d_work, int32_t* d_info, void* stream);

/* ---- evaluate (PGD.evaluate, model.py:724-860; interp1d / Function.__call__ at model.py:798,838).
 * weights: W[k,c] = prod_i phi_{i,k}(p[c,i]) for 1-D Lagrange free dims.  Per free dim i:
 *   xs[i]   ascending cell boundaries [nc_i+1];  cd[i] int32 [nc_i, deg_i+1] dofs of (left, right[, mid]);
 *   Phi[i]  modes [R, ld_i] row-major.  *d_flag set to 1+i if a point lies outside dim i's range. */
int32_t pgd_eval_weights(pgd_handle_t h, int32_t n_free, const double* const* h_xs, const int32_t* const* h_cd,
                         const double* const* h_Phi, const int32_t* h_nc, const int32_t* h_deg, const int64_t* h_ld,
                         int32_t R, const double* d_points, int64_t C, double* d_W, int32_t* d_flag, void* stream);
/* u[n] = sum_k X[k,n] w[k]    (single-point reconstruction, HBM-bound) */
int32_t pgd_eval_gemv(pgd_handle_t h, const double* d_X, int64_t ldx, int32_t R, const double* d_w, int64_t N,
                      double* d_u, void* stream);
/* U[c,n] = sum_k W[k,c] X[k,n]  FP64 tensor-core (DMMA) tile GEMM; W [R,ldw], X [R,ldx], U [C,ldu] */
int32_t pgd_eval_gemm_f64(pgd_handle_t h, const double* d_W, int64_t ldw, const double* d_X, int64_t ldx,
                          int32_t R, int64_t C, int64_t N, double* d_U, int64_t ldu, void* stream);

/* Row-wise reductions of a sweep U [C, ldu] (one row per parameter point) in one launch: d_stats[c*7 + ...] =
 * {min, max, min|u|, max|u|, sum u^2, sum (u-f)^2, sum f^2}, the last two against the reference rows d_F [C, ldf] (NULL: 0).
 * Replaces the per-sample host loops of PGD.evaluate_min/_max/... (model.py:955-1086) and of
 * PGDErrorComputation.evaluate_error (model.py:1785-1825) when a batch of points is evaluated.  Deterministic. */
int32_t pgd_row_stats(pgd_handle_t h, const double* d_U, int64_t ldu, const double* d_F, int64_t ldf, int64_t C, int64_t N,
                      double* d_stats, void* stream);

/* Started / finished solve: pgd_pcg_start enqueues the SM-resident PCG (one cooperative kernel, the result lands in
 * d_x in stream order) and returns without waiting, so the host can record the next sub-problem's forms while the
 * GPU iterates; pgd_pcg_finish waits for it and reports iterations / relative residual (*h_iters = -1 when nothing
 * was started).  Returns 1 (nothing enqueued) unless an earlier pgd_pcg[_x0]_sync call on the same d_rowptr / n /
 * block ran SM-resident, i.e. the system is known to fit; the caller then uses the _sync entry points.  One solve
 * in flight per handle; every other pcg entry point finishes a pending one first.  d_work as for pgd_pcg_sync;
 * warm != 0: d_x holds the initial guess. */
int32_t pgd_pcg_start(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                      const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit, int32_t block,
                      double* d_work, int32_t warm, void* stream);
int32_t pgd_pcg_finish(pgd_handle_t h, int32_t* h_iters, double* h_relres);

/* The same hand-over for the persistent streaming kernel on ONE GPU (systems that do not fit on chip): enqueue
 * pgd_pcg_persist_sync's kernel and return at once; pgd_pcg_finish collects iterations / residual (-3 NaN, -6 a CTA did
 * not arrive).  The fixed-point sweep of solver.py:531-757 solves the spatial dimension first: the host records the
 * forms of the parameter dimensions while the GPU iterates.  Arguments as pgd_pcg_persist_sync with n_local = n_owned
 * = n and no halo description. */
int32_t pgd_pcg_persist_start(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                              const double* d_b, double* d_x, int64_t n, int32_t block, double rtol, double atol,
                              int32_t maxit, int32_t warm, double* d_work, const int32_t* d_bcol,
                              int32_t max_blocks_per_row, void* stream);

/* ---- sensor evaluation (PGD.evaluate_sensor_response, model.py:862-953; eval_fixed_modes with
 * fenicstools.Probes, model.py:107-130).
 * pgd_locate_points: for each of n_points points [n_points, gdim] find the LOWEST-numbered simplex
 * (d_cells int32 [n_cells, gdim+1] into d_coords [n_verts, gdim]) whose barycentric coordinates are all
 * >= -tol; d_cell[i] = that cell or -1 (outside the mesh), d_bary[i, 0..gdim] = barycentric coordinates
 * w.r.t. the cell's vertices in d_cells order (zeros when outside).  Deterministic (atomicMin on the id).
 * d_cells must be aligned to one row (8 B for gdim 1, 16 B for gdim 3: rows are read as one vector load). */
int32_t pgd_locate_points(pgd_handle_t h, const double* d_coords, const int32_t* d_cells, int64_t n_cells, int32_t gdim,
                          const double* d_points, int32_t n_points, double tol, int32_t* d_cell, double* d_bary,
                          void* stream);
/* E[k, r] = sum_j w[r, j] X[k, dofs[r, j]]   (k < R modes, r < n_rows probe rows, nd basis functions per
 * row): every mode of the fixed dimension evaluated at the located points; E [R, lde] is laid out as the X
 * operand of pgd_eval_gemv / pgd_eval_gemm_f64. */
int32_t pgd_probe_modes(pgd_handle_t h, const double* d_X, int64_t ldx, int32_t R, const int32_t* d_dofs,
                        const double* d_w, int64_t n_rows, int32_t nd, double* d_E, int64_t lde, void* stream);

#ifdef __cplusplus
}
#endif
#endif
