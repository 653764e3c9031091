/* c3b200.h -- C ABI of the B200-native PWC propagator engine (libc3b200.so).
 *
 * Drop-in boundary for ONE hot path of q-optimize/c3: the piecewise-constant propagator
 * computation behind Experiment.compute_propagators / c3.libraries.propagation.  The
 * reference is pure Python on TensorFlow and has no FFI of its own; each entry point below
 * names the reference function (file:line, relative to the reference checkout @ 48b7917e)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the library never
 *     allocates or frees: the caller owns inputs, outputs and the workspace;
 *   - complex128 is interleaved (re, im) doubles, matrices are row-major, exactly the
 *     memory image of a contiguous numpy/torch complex128 tensor;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing
 *     synchronises; the library keeps no process-wide mutable state except an atomic launch
 *     counter (error message, tuning knobs and profile events are per calling thread), so calls
 *     are re-entrant: one thread per GPU is the intended use;
 *   - return value 0 = success, otherwise a negative C3B_E* code; c3b_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - ordered products put LATER slices on the LEFT: U = dU_{N-1} ... dU_1 dU_0
 *     (c3/utils/tf_utils.py:120-129,144-193).
 */
#ifndef C3B200_H
#define C3B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C3B_OK 0
#define C3B_EINVAL (-1)   /* bad argument (null pointer, non-positive size, ...) */
#define C3B_EWORKSPACE (-2) /* workspace too small */
#define C3B_ECUDA (-3)    /* a CUDA runtime call failed */
#define C3B_EUNSUPPORTED (-4)

/* Library version (major*10000 + minor*100 + patch). */
int c3b_version(void);

/* Message of the last error on this thread ("" if none). */
const char* c3b_last_error(void);

/* Bytes of device workspace needed by c3b_pwc_closed / _hlist / _lindblad for the given
 * problem.  `lindblad` != 0 means d is the Hilbert dimension and the matrices are d^2 x d^2.
 * `batched_model` != 0 means h0/hks (and col_ops) carry a leading batch axis of length B. */
size_t c3b_pwc_workspace_bytes(int B, int K, int N, int d, int lindblad, int batched_model);

/* Closed-system PWC propagators for a whole batch of control signals.
 *   replaces  tf_batch_propagate + tf_propagation_vectorized + tf.linalg.expm
 *             (c3/libraries/propagation.py:460-515, 426-440) and tf_matmul_n / tf_matmul_left
 *             (c3/utils/tf_utils.py:120-193), called B times by the reference's serial loops.
 *   h0       [d,d]     or [B,d,d]   if batched_model       drift Hamiltonian
 *   hks      [K,d,d]   or [B,K,d,d] if batched_model       control Hamiltonians (may be NULL if K==0)
 *   signals  [B,K,N]   float64, contiguous in N            control fields c_k[b,n]
 *   dt                                                      slice length
 *   U_out    [B,d,d]   product of all N slice propagators
 *   dUs_out  [B,N,d,d] every slice propagator expm(-i H_n dt), or NULL to skip the store
 */
int c3b_pwc_closed(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                   int batched_model, void* U_out, void* dUs_out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Gated variant of c3b_pwc_closed for HOST-resident control fields (shared model, no dUs, d = 9: the headline
 * kernel).  The caller enqueues the host->device copies of `signals` in batch order on a copy stream, each chunk
 * followed by a 4-byte copy that raises gate[0] to the number of batch rows that have landed, and launches this ONE call
 * on another stream as soon as the first chunk is in.  Warps take batch rows in order and wait on gate[0] only if they
 * overtake the copy engine, so PCIe time hides behind the whole-batch kernel instead of cutting it into per-chunk
 * launches.  In-flight rows are never read through the read-only path: past L1 (ld.global.cg) in general, through L1
 * (ld.global.ca) when every row starts on its own 128-byte line (K*N*8 a multiple of 128 and `signals` 128-byte aligned), where
 * a line can only be fetched after the acquire has seen its row.
 *   gate [2] uint32, device memory, both words zero before the first chunk:
 *        gate[0]  rows landed (written by the caller's copy stream, read with acquire loads by the kernel)
 *        gate[1]  set to 1 by the kernel if some row did not arrive within ~4 s WITHOUT copy progress: the warp that
 *                 waited stops taking work (no GPU hang) and the rows it never computed stay as the caller pre-filled
 *                 them (fill U_out with NaN).  The caller reads gate[1] at its next synchronisation point and turns a
 *                 non-zero value into an error -- the call itself returns before the kernel has run.
 * c3b_pwc_gated_supported(d) != 0 says whether the dimension takes this path (with the calling thread's tuning). */
int c3b_pwc_closed_gated(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                         void* U_out, uint32_t* gate, void* workspace, size_t workspace_bytes, void* stream);
int c3b_pwc_gated_supported(int d);

/* Prepared models: everything of a launch that depends on the MODEL only -- the generators G_0 = -i dt h0,
 * G_k = -i dt h_k (closed system) or the Lindblad superoperator generators (c3/libraries/propagation.py:563-582),
 * trace-shifted, with their row sums -- built once per model update instead of once per call.  The reference
 * rebuilds these inside every tf_propagation_* call (propagation.py:426-440, 551-585); an optimiser that only changes
 * pulse parameters calls c3b_model_prepare once and c3b_pwc_prepared per evaluation (one fused launch, plus one fold
 * launch when the time axis is segmented).
 *   h0 [n_models,d,d], hks [n_models,K,d,d], col_ops [n_models,C,d,d] (n_models = 1: shared model; = B: one per row)
 *   model_out: caller-owned device buffer of c3b_model_bytes(K, d, lindblad, n_models) bytes, 16-byte aligned.
 * c3b_pwc_prepared: same outputs as c3b_pwc_closed / c3b_pwc_lindblad (U_out [B,D,D], dUs_out [B,N,D,D] or NULL,
 * D = d or d*d); workspace of c3b_pwc_prepared_workspace_bytes(B, N, d, lindblad, n_models) bytes. */
size_t c3b_model_bytes(int K, int d, int lindblad, int n_models);
int c3b_model_prepare(const void* h0, const void* hks, const void* col_ops, int C, double dt, int K, int d, int lindblad,
                      int n_models, void* model_out, size_t model_bytes, void* stream);
size_t c3b_pwc_prepared_workspace_bytes(int B, int N, int d, int lindblad, int n_models);
int c3b_pwc_prepared(const void* model, const double* signals, int B, int K, int N, int d, int lindblad, int n_models,
                     void* U_out, void* dUs_out, void* workspace, size_t workspace_bytes, void* stream);

/* Same with explicit per-slice Hamiltonians (the reference's `signals is None` branch,
 * c3/libraries/propagation.py:294-308, 491-499, 437-438):  Hs [B,N,d,d]. */
int c3b_pwc_closed_hlist(const void* Hs, double dt, int B, int N, int d, void* U_out, void* dUs_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Lindblad (open-system) superoperator propagators, D = d*d:
 *   replaces tf_propagation_lind (c3/libraries/propagation.py:551-585) incl. the Kronecker
 *   helpers tf_kron/tf_spre/tf_spost (c3/utils/tf_utils.py:257-280), which are never
 *   materialised per slice.
 *   col_ops [C,d,d] or [B,C,d,d];  U_out [B,D,D];  dUs_out [B,N,D,D] or NULL. */
int c3b_pwc_lindblad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                     int B, int K, int N, int d, int batched_model, void* U_out, void* dUs_out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Gradient of a real scalar loss L through the closed-system propagators with respect to the control
 * fields: what tf.GradientTape gives the reference's gradient optimisers
 * (c3/optimizers/optimizer.py:210-215, 277-313; c3/libraries/algorithms.py:391-420).
 *   Ubar     [B,d,d]  cotangent of U with dL = Re tr(Ubar^dag dU)  (torch's grad_output for U)
 *   grad_out [B,K,N]  float64, dL/d signals[b,k,n]
 *   U_out    [B,d,d]  or NULL: the forward result is produced on the way
 *   chunk    batch rows processed per pass (bounds the workspace: ~2 N d^2 16 bytes per row; ~10 N d^2 16 with the
 *            augmented-exponential cross-check, tuning "grad_variant" 0, d <= 32); <= 0: all
 * Shared model only (h0 [d,d], hks [K,d,d]).  d = 7..9 and 16 < d <= 32 with Hermitian h0 / hks: ONE fused kernel,
 * no stored slice propagators -- Y_n = F_n Ubar^dag U F_n^-1 is carried forward by unitarity and the Frechet derivative of
 * the Taylor scheme runs in lockstep with the scheme (workspace ~ N/8 d^2 16 bytes per row).  Whether the Hamiltonians are
 * Hermitian is checked on the device (one 4-byte read-back per call, i.e. a stream synchronisation; tuning "grad_unitary"
 * 1 / 0 asserts the answer and skips it); if they are not, or tuning "grad_variant" is 2, the stored-propagator kernels
 * below serve the call.  Other d <= 16: one warp per slice (Frechet derivative of the Taylor scheme in shared memory);
 * 16 < d <= 128: CTA kernels on the fp64 tensor-core product (sweeps + Frechet derivative on (X, dX) pairs). */
size_t c3b_pwc_grad_workspace_bytes(int B, int K, int N, int d, int chunk);
int c3b_pwc_closed_grad(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d,
                        const void* Ubar, double* grad_out, void* U_out, int chunk, void* workspace,
                        size_t workspace_bytes, void* stream);

/* The same gradient as two calls, for an autograd node (forward now, backward when the cotangent arrives) -- available where
 * the fused kernels are (closed d = 7..9 and 16 < d <= 32, Hermitian h0 / hks; c3b_pwc_closed_saved_bytes returns 0 elsewhere
 * and c3b_pwc_closed_fwd_saved returns C3B_EUNSUPPORTED for non-Hermitian Hamiltonians: use c3b_pwc_closed_grad then):
 *   c3b_pwc_closed_fwd_saved   U_out [B,d,d]; leaves the model, the chunk products and their prefix products in `state`
 *                              (caller-owned device buffer of c3b_pwc_closed_saved_bytes bytes, untouched until the backward call)
 *   c3b_pwc_closed_saved_chunks  the chunking (Q chunks of CL slices) the forward call used for this shape and tuning
 *   c3b_pwc_closed_bwd_saved   grad_out [B,K,N] from Ubar [B,d,d], the same signals and the state: no forward pass is repeated
 * forward + backward = 3.0 + 1.0 forward passes' worth of work instead of 1.0 + 4.0 through c3b_pwc_closed + c3b_pwc_closed_grad. */
size_t c3b_pwc_closed_saved_bytes(int B, int K, int N, int d);
int c3b_pwc_closed_saved_chunks(int B, int K, int N, int d, int* Q_out, int* CL_out);
int c3b_pwc_closed_fwd_saved(const void* h0, const void* hks, const double* signals, double dt, int B, int K, int N, int d, void* U_out,
                             void* state, size_t state_bytes, void* stream);
int c3b_pwc_closed_bwd_saved(const double* signals, int B, int K, int N, int d, int Q, int CL, const void* Ubar, double* grad_out,
                             void* state, size_t state_bytes, void* stream);

/* Gate infidelities straight from a batch of propagators: every goal function
 * of c3/libraries/fidelities.py on a single gate is a function of the gathered overlap
 *     t[b] = sum_{I,J} U[b, sel[I], sel[J]] * conj(ideal[I,J])
 * (tf_project_to_comp + trace, c3/utils/tf_utils.py:326-364, 380-413, 430-438).
 *   U      [B,D,D]   propagators (D = d, or d^2 for Lindblad superoperators)
 *   ideal  [C,C]     ideal gate G on the computational subspace (modes 2/3: G (x) G^*, C = c^2)
 *   sel    [C] int32 rows/columns of U kept by the projector (qt_utils.projector, c3/utils/qt_utils.py:178-193)
 *   mode   0 unitary_infid (fidelities.py:152-183)            1 - |t|^2 / C^2
 *          1 average_infid (:288-311)                          1 - (|t|^2 / C + 1) / (C + 1)
 *          2 lindbladian_unitary_infid (:221-249)              1 - |t| / C
 *          3 lindbladian_average_infid (:377-399)              1 - |conj(t)/sqrt(C) + 1| / (sqrt(C) + 1)
 *   infid_out [B] float64 (or NULL), overlap_out [B] complex128 (or NULL; needed by the gradient). */
int c3b_gate_infid(const void* U, int B, int D, const void* ideal, const int32_t* sel, int C, int mode,
                   double* infid_out, void* overlap_out, void* stream);

/* Cotangent of U for the loss sum_b gbar[b] * infid[b] (modes 0 and 1), in the convention c3b_pwc_closed_grad
 * takes (dL = Re tr(Ubar^dag dU)): what tf.GradientTape propagates from the goal function back to the
 * propagator (c3/optimizers/optimalcontrol.py:200-228).  gbar [B] float64 or NULL (= ones); Ubar_out [B,D,D]. */
int c3b_gate_infid_grad(const void* overlap, const void* ideal, const int32_t* sel, const double* gbar, int B, int D,
                        int C, int mode, void* Ubar_out, void* stream);

/* Final states and populations of gate sequences applied to psi0 (NULL: basis state 0):
 *   psi_s = gates[idx[s,len_s-1]] ... gates[idx[s,0]] psi0  as a chain of matrix-vector products.
 *   replaces the gate loop of Experiment.evaluate_legacy + populations (c3/experiment.py:273-302, 603-624) and
 *   evaluate_sequences + matmul + pop0 in orbit_infid (c3/libraries/fidelities.py:753-790).
 *   lindblad_d == 0: pops_out [S,D] = |psi|^2;  lindblad_d = d (D = d^2, psi = density vector):
 *   pops_out [S,d] = Re diag(rho).  psi_out [S,D] or NULL; pops_out may be NULL if psi_out is given. */
int c3b_seq_populations(const void* gates, int Gn, const int32_t* seq_idx, const int32_t* seq_len, int S, int Lmax,
                        int D, const void* psi0, int lindblad_d, double* pops_out, void* psi_out, void* stream);

/* Control fields from pulse parameters for a whole batch of parameter samples: the
 * standard device chain of c3/generator/generator.py:172-229,
 *   LO + AWG(envelopes) -> DigitalToAnalog -> Response|ResponseFFT -> Mixer -> VoltsToHertz|FluxTuning
 * (c3/generator/devices.py:1063-1130, 1159-1197, 296-351, 585-701, 906-939, 187-221, 480-529; envelopes
 * c3/signal/gates.py:341-370, c3/signal/pulse.py:93-180, c3/libraries/envelopes.py), one kernel, output written
 * directly in the [B,K,N] layout c3b_pwc_closed reads.  All values are what Quantity.get_value() returns
 * (angular frequencies).
 *   env_params [B,K,E,9]  amp, t_final, sigma, xy_angle, freq_offset, delta, t_up, t_down, risefall
 *   env_shape  [K,E] int32  0 no_drive, 1 rect, 2 gaussian_nonorm, 3 gaussian_sigma, 4 cosine, 5 flattop; < 0: unused
 *   env_flags  [K,E] int32  bit 0: EnvelopeDrag quadrature (-delta * d env/dt * dt), bit 1: use_t_before
 *   lo_freq    [B,K]        carrier frequency of the line
 *   chain      [K,11] (or [B,K,11] if chain_batched): sim_res, awg_res, rise_time, response kind (0 none,
 *              1 Response, 2 ResponseFFT), output kind (0 VoltsToHertz, 1 FluxTuning), V_to_Hz, phi, phi_0,
 *              omega_0, anhar, d (NaN: no junction asymmetry)
 *   N = c3b_signal_slice_num(t_start, t_end, sim_res) (Device.calc_slice_num); the AWG grid must not be finer
 *   than the simulation grid.  signals_out [B,K,N] float64. */
int c3b_signal_slice_num(double t_start, double t_end, double resolution);
int c3b_generate_signals(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                         const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                         int B, int K, int E, int N, double* signals_out, void* stream);

/* Same chain with the reference's NOISE devices in it (c3/generator/devices.py:943-1035; exercised by test/test_noise.py:93-138),
 * one independent realisation per batch row -- the Monte-Carlo trajectory axis of the batch:
 *   noise [K,7] (or [B,K,7] if noise_batched), per drive line:
 *     0 awg_amp   Additive_Noise behind the AWG: amp * N(0,1) on every in-phase and quadrature AWG sample
 *     1 lo_perc   LONoise: perc * N(0,1) on the local oscillator's cos and sin at every simulation sample
 *     2 add_amp   Additive_Noise behind the mixer: amp * N(0,1) per simulation sample
 *     3 dc_amp    DC_Noise: one amp * N(0,1) offset per realisation
 *     4 pink_amp  Pink_Noise: amp * (sum of bfl_num two-level fluctuators), fluctuator i flipping at a step with probability
 *     5 bfl_num       1 / rate_i, rate = logspace(0, ln N, bfl_num + 1, base 10)[1:]  (<= 32 fluctuators)
 *     6 dc_offset DC_Offset: deterministic offset behind the mixer
 *   amplitudes below 1e-17 switch a device off exactly (as the reference does).
 *   seed: the random numbers are Philox4x32-10 words of (seed, batch row * K + line, device stream, sample index): no state,
 *   reproducible, independent across rows, lines, devices and samples; advance the seed for a new realisation.
 *   noise_out [B,K,7,N] or NULL: the realised traces (AWG in-phase / quadrature on the AWG grid in the first n_awg entries,
 *   LO cos / sin, additive, dc, pink) -- what Device.signal["noise"] holds in the reference. */
int c3b_generate_signals_noisy(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                               const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                               int B, int K, int E, int N, const double* noise, int noise_batched, unsigned long long seed,
                               double* signals_out, double* noise_out, void* stream);

/* Same with the array-parametrised and grid-defined envelope shapes of c3/libraries/envelopes.py (shape ids 12..23: flattop_cut,
 * flattop_cut_center, flattop_variant, cosine_flattop, delta_pulse, pwc, pwc_shape, pwc_symmetric, pwc_shape_plateau, fourier_sin,
 * fourier_cos, slepian_fourier):
 *   env_table [K,E,T] float64 or NULL (T = 0): per envelope, the array parameters in the order signal_chain.cuh documents next
 *             to each shape id (e.g. pwc: M, inphase[M], quadrature[M]); shared by the batch rows.
 * Scalar parameters without a column of their own ride in the sigma column (width, ramp, t_rise).  These shapes are forward
 * only: c3b_generate_signals_grad writes NaN into their grad_env rows. */
int c3b_generate_signals_table(const double* env_params, const int32_t* env_shape, const int32_t* env_flags, const double* env_table,
                               int T, const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                               int B, int K, int E, int N, const double* noise, int noise_batched, unsigned long long seed,
                               double* signals_out, double* noise_out, void* stream);

/* Reverse mode of c3b_generate_signals: from dL/d signals [B,K,N] (e.g. the output of c3b_pwc_closed_grad) to the
 * gradient with respect to the pulse parameters -- what tf.GradientTape propagates through
 * Generator.generate_signals in the reference's gradient-based optimal control (c3/optimizers/optimalcontrol.py:200-228).
 *   grad_env  [B,K,E,9]  dL / d(amp, t_final, sigma, xy_angle, freq_offset, delta, t_up, t_down, risefall)
 *   grad_lo   [B,K]      dL / d(carrier frequency)
 *   grad_v2hz [B,K]      dL / d(V_to_Hz) (VoltsToHertz lines; 0 for FluxTuning lines), or NULL
 *   n_awg_max            largest AWG sample count of any line (bounds shared memory); <= 0: N + 1 */
int c3b_generate_signals_grad(const double* env_params, const int32_t* env_shape, const int32_t* env_flags,
                              const double* lo_freq, const double* chain, int chain_batched, double t_start, double t_end,
                              int B, int K, int E, int N, int n_awg_max, const double* gsignals, double* grad_env,
                              double* grad_lo, double* grad_v2hz, void* stream);

/* Same for the Lindblad superoperator propagators (Ubar, U_out [B,D,D], D = d*d): the generators dA/dc_k are the
 * commutator superoperators -i dt (h_k (x) I - I (x) h_k^T).  D <= 128: the BASELINE Lindblad shape D = 81 runs on the CTA
 * kernels with a per-CTA global workspace. */
size_t c3b_pwc_lindblad_grad_workspace_bytes(int B, int K, int N, int d, int chunk);
int c3b_pwc_lindblad_grad(const void* h0, const void* hks, const void* col_ops, int C, const double* signals, double dt,
                          int B, int K, int N, int d, const void* Ubar, double* grad_out, void* U_out, int chunk,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Dressing of a batch of model samples: eigendecomposition of every drift Hamiltonian,
 * reordering of the eigenvectors by overlap with the bare states, and T^dag X T of the sample's operators --
 * Model.update_drift_eigen / reorder_frame / update_dressed (c3/model.py:453-534), once per sample instead of once
 * per host-side model update.
 *   drift [B,d,d] Hermitian;  ops [M,d,d] (or [B,M,d,d] if ops_batched) control Hamiltonians / collapse operators, or NULL
 *   ordered != 0: reorder_frame(ordered=True); 0: ascending eigenvalues, T = eigenvectors
 *   eigenframe [B,d] float64;  transform [B,d,d];  dressed_drift [B,d,d] or NULL;  dressed_ops [B,M,d,d] or NULL
 *   info [B] int32 (Jacobi sweeps used, 100 = not converged) or NULL.   d <= 32. */
int c3b_dress_models(const void* drift, const void* ops, int ops_batched, int B, int M, int d, int ordered,
                     double* eigenframe, void* transform, void* dressed_drift, void* dressed_ops, int32_t* info,
                     void* stream);

/* Frame rotation and dephasing channel of Experiment.compute_propagators (c3/experiment.py:482-522; Model.get_Frame_Rotation
 * c3/model.py:536-578, Model.get_dephasing_channel :597-639) applied IN PLACE to a batch of propagators.  Both operators are
 * functions of bare number operators, i.e. diagonal in the product basis, so the left-multiplication is a row scaling:
 *   closed (lindblad = 0, D = d):   U[b][r, :]     *= exp(1j sum_l occ[l,r] phases[b,l])
 *   Lindblad (D = d^2, r = i d + j): S[b][r, :]     *= exp(1j sum_l (occ[l,i] - occ[l,j]) phases[b,l])
 *                                                     * prod_l ((1 - probs[b,l]) + probs[b,l] (-1)^(occ[l,i] - occ[l,j]))
 *   occ    [L,d] int32   occupation number of the qubit driven by line l in product state s (diag of a_q^dag a_q)
 *   phases [B,L]         freq_l * t_final + framechange_l per batch row (gate or parameter sample), or NULL (no frame rotation)
 *   probs  [B,L]         t_final * amp_l * dephasing_strength in [0,1], or NULL (no dephasing); Lindblad only. */
int c3b_frame_dephase(void* U, int B, int D, int d, const int32_t* occ, int L, const double* phases, const double* probs,
                      int lindblad, void* stream);

/* Crosstalk device (c3/generator/devices.py:225-293, applied by Generator.generate_signals to the finished lines,
 * c3/generator/generator.py:229-234), IN PLACE on signals [B,K,N]:
 *   out[b, chan[i], n] = sum_j matrix[i,j] * in[b, chan[j], n],   chan [C] int32 (distinct line indices), matrix [C,C], C <= 16. */
int c3b_crosstalk(double* signals, int B, int K, int N, const int32_t* chan, int C, const double* matrix, void* stream);

/* Ordered product of M matrices per batch row: out[b] = mats[b,M-1] ... mats[b,0].
 *   replaces tf_matmul_left (c3/utils/tf_utils.py:120-129) and tf_matmul_n (:144-193).
 *   mats [B,M,D,D], out [B,D,D]. */
size_t c3b_product_workspace_bytes(int B, int M, int D);
int c3b_ordered_product(const void* mats, int B, int M, int D, void* out, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Gate-sequence propagators: out[s] = gates[idx[s,len_s-1]] ... gates[idx[s,0]], identity
 * for an empty sequence.
 *   replaces evaluate_sequences (c3/libraries/propagation.py:588-627).
 *   gates [Gn,D,D];  seq_idx [S,Lmax] int32 (entries >= seq_len[s] ignored);  seq_len [S] int32;
 *   out [S,D,D]. */
int c3b_seq_product(const void* gates, int Gn, const int32_t* seq_idx, const int32_t* seq_len, int S, int Lmax,
                    int D, void* out, void* workspace, size_t workspace_bytes, void* stream);

/* Batched Kronecker product out[b] = A[b] (x) Bm[b]  (tf_kron, c3/utils/tf_utils.py:257-267).
 *   A [batch,ra,ca] (or [ra,ca] with a_batched==0), Bm likewise, out [batch, ra*rb, ca*cb]. */
int c3b_kron(const void* A, const void* Bm, void* out, int batch, int ra, int ca, int rb, int cb, int a_batched,
             int b_batched, void* stream);

/* Tuning knobs of the CALLING THREAD (thread-local: one thread per GPU is the intended use, and a thread that flips a
 * knob for an experiment cannot disturb another thread's launches).  Keys: "target_units" (0 automatic), "min_chunk" (segmentation of
 * the time axis), "d9_variant" (0 generic 3x3-block kernel, 1 own-block shared-memory kernel, 2 shuffle-exchange kernel),
 * "force_cta", "cta_variant" (0 literal Higham cross-check, 1 four-product Taylor scheme on DMMA tiles), "cta_threads", "gemm_big",
 * "norm_bound", "seq_variant", "grad_variant" (1 best available, 0 augmented exponential, 2 stored-propagator kernels),
 * "grad_unitary" (-1 check the Hamiltonians on the device, 1 Hermitian, 0 not), "grad_chunk" (slices per chunk of the fused
 * gradient kernels, 0 automatic), "profile".  Returns C3B_EINVAL for an unknown key. */
int c3b_set_tuning(const char* key, long long value);

/* Which kernel c3b_pwc_* would pick for this shape: 1 = lane-group kernel (d <= 12, shared model),
 * 2 = DMMA CTA kernel with shared-memory matrices, 3 = DMMA CTA kernel with a global workspace. */
int c3b_pwc_path(int K, int D, int batched_model);

/* Measured fp64 throughput of this GPU in TFLOP/s (2 flops per FMA) -- the roofline
 * denominators bench.py reports against.  kind 0: DFMA (fp64 vector pipe),
 * kind 1: DMMA (mma.sync m8n8k4 f64).  Synchronises the device; not for hot paths.
 * Returns a negative error code on failure. */
double c3b_measure_fp64_peak(int kind, int device, double seconds);

/* Number of CUDA kernels this library has launched in this process so far. */
long long c3b_launch_count(void);

/* With tuning key "profile" = 1, every c3b_pwc_* call of this thread brackets its main (fused) kernel with CUDA
 * events on the caller's stream (events live on that stream's device); this returns the duration in ms of the
 * thread's last one (synchronises on that event).  Negative on error. */
double c3b_last_kernel_ms(void);


#ifdef __cplusplus
}
#endif
#endif /* C3B200_H */
