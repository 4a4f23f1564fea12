/*
 * hlm_b200.h — C ABI of libhlm_b200.so: batched Dormand–Prince RK45 integration of per-link
 * runoff ODEs on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of PrincetonUniversity/Tiger_HLM_GPU: what its host
 * code does between "per-link parameters, forcings and y0 are in host memory" and "final and
 * dense states are back in host memory".  Each entry point names the reference interface it
 * replaces (paths relative to the reference's src/).  Plain pointers and sizes only; no C++ types,
 * no exceptions, no templates cross this boundary.  The header-only C++ shims in
 * include/hlm_b200/rk45_api.hpp rebuild the reference's own operator surface
 * (rk45_api::setModelParameters<T>(), rk45_api::run_rk45<T>(), the Model trait) on top of it.
 *
 * Conventions
 *   - every function returns 0 on success or a negative hlm_status; hlm_last_error() then holds
 *     a message for the calling thread (the reference throws std::runtime_error instead,
 *     solver/rk45_api.hpp:87-108 — the shims re-throw);
 *   - times are in MINUTES, query times ascending (solver/rk45_kernel.cu:139);
 *   - a context is bound to one CUDA device and is not thread-safe; create one per device
 *     (the reference: one MPI rank per GPU and process-global __constant__ state);
 *   - there is no CPU fallback: every entry point that computes fails with HLM_ERR_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef HLM_B200_H
#define HLM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hlm_ctx hlm_ctx;

typedef enum hlm_status {
    HLM_OK = 0,
    HLM_ERR_INVALID = -1, /* bad argument (null pointer, size, unknown model uid, ...) */
    HLM_ERR_CUDA = -2,    /* CUDA runtime failure, message has the CUDA error string */
    HLM_ERR_STATE = -3,   /* call out of order (e.g. run before parameters were uploaded) */
    HLM_ERR_NOMEM = -4    /* host or device allocation failed */
} hlm_status;

/* Model UIDs compiled into the library.  204 is the reference's Model204::UID
 * (models/model_204.hpp:18); 0 is the DummyModel of model_dummy_python.ipynb:65-89, to which the
 * reference never assigned a UID. */
#define HLM_MODEL_DUMMY 0
#define HLM_MODEL_204 204
/* 200 = hillslope-link runoff (channel discharge + Model204's hillslope stores without snow).  The
 * reference names "model 200" (README.md:95) but ships no definition: this one is project-defined
 * (csrc/models.cuh), pinned by its own CPU restatement and SciPy only.  It is the routed model: its
 * channel takes the discharge of upstream links (hlm_route_*). */
#define HLM_MODEL_200 200

/* per-link result codes written to out_stiff: 0 = integrated to tf; 1 = flagged stiff and
 * abandoned, exactly the reference's d_stiff[sys] = 1 (solver/rk45_kernel.cu:160-170);
 * 2 = attempt budget of hlm_set_max_attempts() exhausted (no reference equivalent: the reference
 * would spin forever). */
#define HLM_LINK_OK 0
#define HLM_LINK_STIFF 1
#define HLM_LINK_STALLED 2
/* 3 = flagged stiff by the RK45 path, then carried to tf by the implicit fallback of
 * hlm_set_stiff_fallback(): final state and dense records are valid. */
#define HLM_LINK_STIFF_SOLVED 3

/* ---- context ------------------------------------------------------------------------------- */

/* Bind a context to CUDA device `device`.  Replaces the per-rank cudaSetDevice of main.cpp:314-319. */
int hlm_create(int device, hlm_ctx** out);
void hlm_destroy(hlm_ctx* ctx);
/* Message of the last failure on this thread ("" if none). */
const char* hlm_last_error(void);
/* Library/ABI version, for the loader check in the Python mirror. */
int hlm_abi_version(void);
/* Run all work of this context on `cuda_stream` (a cudaStream_t); NULL = the context's own stream.
 * The reference uses the default stream and cudaDeviceSynchronize (solver/rk45_api.hpp:150). */
int hlm_set_stream(hlm_ctx* ctx, void* cuda_stream);
/* Block until all work queued by this context has finished. */
int hlm_synchronize(hlm_ctx* ctx);

/* ---- model registry -------------------------------------------------------------------------- */

/* N_EQ, number of SoA parameter columns and number of forcings read by model `uid`.
 * Replaces the compile-time Model::N_EQ of the trait (models/model_204.hpp:19). */
int hlm_model_info(int uid, int* n_eq, int* n_sp, int* n_forc);

/* Replaces rk45_api::setModelParameters<T>(const T::Parameters&) (model_registry.hpp:9-13,
 * model_registry.cpp:18-60), which copies 48 bytes into __constant__ devParams.
 * p = {initialStep, rtol, atol, safety, minScale, maxScale} (models/model_204.hpp:22-30). */
int hlm_set_model_parameters(hlm_ctx* ctx, int uid, const double p[6]);
int hlm_get_model_parameters(hlm_ctx* ctx, int uid, double p[6]);

/* ---- per-link inputs ------------------------------------------------------------------------- */

/* Replaces the cudaMalloc/cudaMemcpy of the AoS SpatialParams array (main.cpp:392-404) and, across
 * GPUs, the MPI rank-0 scatter of its row chunks (main.cpp:269-309,357-366): the caller hands each
 * device context its own contiguous slice.  `aos` points at `n` records `stride_bytes` apart, each
 * laid out as the reference's 136-byte SpatialParams (I_O/parameters_loader.hpp:19-37), in HOST memory
 * or in DEVICE memory (the d_sp the reference's run_rk45 receives, solver/rk45_api.hpp:279; the copy
 * direction is taken from the pointer).  The records are transposed on the device into the
 * structure-of-arrays columns each model needs. */
int hlm_upload_spatial_params(hlm_ctx* ctx, const void* aos, long long n, long long stride_bytes);

/* Replaces the forcing side channel: d_forc_data upload plus cudaMemcpyToSymbol of c_forc_dt /
 * c_forc_nT (main.cpp:552-574, I_O/forcing_data.h:5-13).  Forcing j (0 = precipitation, 1 = 2 m
 * temperature for Model204) is a float array [nT][ncols], sampled with step-hold at
 * index size_t(t / (dt_hours*60)) clamped to [0, nT-1] (solver/rk45_kernel.cu:90-98).
 * With hlm_set_forcing_columns(NULL) column c serves link c — the reference's per-link-expanded
 * [forcing][time][system] layout (main.cpp:543-548).  With a column map the array is the forcing
 * GRID (ncols = lat*lon cells) and link s reads column col[s] = lat_index*lon_size + lon_index
 * (main.cpp:501-505); the per-link expansion is never materialised. */
int hlm_upload_forcing(hlm_ctx* ctx, int j, double dt_hours, long long nT, long long ncols, const float* data);
/* The same with only samples [i0, i0 + nT_chunk) of an nT_total-sample record resident: the
 * time-chunked loading of I_O/forcing_loader.cpp:165-196 (NetCDFLoader::loadTimeChunk) carried to the
 * device.  Indexing and clamping stay those of the whole record; hlm_solve_window fails with
 * HLM_ERR_STATE when the resident chunk does not cover the interval being integrated.  The upload is
 * ordered after the windows already queued, so the next interval's chunk can be sent while the host
 * writes the previous one's output. */
int hlm_upload_forcing_chunk(hlm_ctx* ctx, int j, double dt_hours, long long nT_total, long long i0,
                             long long nT_chunk, long long ncols, const float* data);
int hlm_set_forcing_columns(hlm_ctx* ctx, const int* col, long long n);
int hlm_clear_forcings(hlm_ctx* ctx);

/* ---- solver options (no reference equivalent) ------------------------------------------------- */

/* Attempt budget per link per window launch; <= 0 = unbounded like the reference.  Default 0. */
int hlm_set_max_attempts(hlm_ctx* ctx, long long per_link);
/* How many consecutive rejected attempts flag a link stiff: more than `n`.  Default 5, the reference's rule
 * (`reject_count > 5`, solver/rk45_kernel.cu:160).  The rule doubles as a kink detector: across a switch of one of the
 * model's min/max terms the error estimate falls like h, not h^5, and the controller needs a run of 6-8 rejections
 * to get past — not stiffness.  Routed runs, where one abandoned link starves everything downstream and the
 * implicit fallback costs a serial chain of Newton solves per interval, raise it (20) and leave true stiffness to
 * the other test, h < (tf - t0) * 1e-6; so do the drivers for unrouted Model 200 runs (hlm_run, bench.py), whose first
 * day otherwise hands 1.5 % of the links to the fallback for seconds. */
int hlm_set_reject_limit(hlm_ctx* ctx, int n);
/* Bytes of device memory one dense-output window buffer may take (two are allocated when the run
 * needs more than one window).  Default 8 GiB. */
int hlm_set_dense_window_bytes(hlm_ctx* ctx, long long bytes);
/* What happens to a link the RK45 path flags stiff (solver/rk45_kernel.cu:160-170).  0 (default): it is
 * abandoned, the state of the path in the reference whenever its Radau kernel is left out.  1: after every
 * window the flagged links are carried on by a 3-stage Radau IIA integrator — the role of
 * radau_kernel_multi in run_rk45 (solver/rk45_api.hpp:198-247, solver/radau_kernel.cu:20-140), with the
 * numerics done properly (csrc/radau_fallback.cuh says what was kept and what was replaced).  FP64 only. */
int hlm_set_stiff_fallback(hlm_ctx* ctx, int enable);
/* How links are dealt to lanes.  TILES: a warp takes 32 consecutive links and stays with them until the slowest
 * is done — right when neighbouring links step alike (links sorted by forcing cell).  LANES: a lane takes the
 * next unclaimed link the moment it is done with its own, so a warp never idles behind its slowest link —
 * right when links take unlike numbers of attempts per launch (routed runs with short coupling intervals).
 * AUTO (default): SORTED_TILES (below) for Model 200 and for routed runs, falling back to LANES for a launch that has no
 * attempt counts to sort by yet (the first one, or one over part of the links); TILES otherwise.  Results are
 * bit-identical under every schedule. */
#define HLM_SCHEDULE_AUTO 0
#define HLM_SCHEDULE_TILES 1
#define HLM_SCHEDULE_LANES 2
/* SORTED_TILES: tiles of 32 links that took the SAME number of attempts in the previous launch, whatever their
 * indices — lockstep tiles where neighbouring links differ (FP64; a launch that has no counts to sort by yet, or covers
 * part of the links, takes LANES).  For a Model204 data set whose neighbouring links do not step alike. */
#define HLM_SCHEDULE_SORTED_TILES 3
int hlm_set_schedule(hlm_ctx* ctx, int mode);
/* 64 (default, the reference's arithmetic) or 32 (FP32 state/stages; no reference counterpart). */
int hlm_set_precision(hlm_ctx* ctx, int bits);
/* Which states a dense record carries: bit i of `mask` = state i of the model, records hold the selected states
 * in ascending order; 0 (default) = all N_EQ states, the reference's layout.  This is config.yaml's
 * `output.states` (data/config.yaml, main.cpp:788-793) applied where the records are produced: the window kernel
 * interpolates and stores only the selected states, so a discharge-only run of Model 200 moves a fifth of the
 * bytes over PCIe.  Every dense buffer of the ABI (hlm_run_rk45's out_dense, hlm_solve_fetch_window*, the device
 * window buffer) then is [ns][nq][n_selected].  Final states are always all N_EQ.  Takes effect for windows
 * queued after the call. */
int hlm_set_output_states(hlm_ctx* ctx, unsigned int mask);
/* Dense records as double (64, default: the reference's type, I_O/output_series.cpp:88-123) or float (32): the
 * states are integrated in the solver's precision and rounded once when the record is stored.  Halves the output
 * bytes; an archive type, like NetCDF's own NC_FLOAT packing.  Dense buffers then hold float. */
int hlm_set_output_precision(hlm_ctx* ctx, int bits);
/* Columns per dense record and bytes per value for model `uid` under the current two settings. */
int hlm_output_layout(hlm_ctx* ctx, int uid, int* n_columns, int* bytes_per_value);

/* ---- the operator ---------------------------------------------------------------------------- */

/* Replaces rk45_api::run_rk45<T>(h_y0, t0, tf, h_query_times, d_sp) (solver/rk45_api.hpp:273-313)
 * = setup_gpu_buffers + launch_rk45_kernel + retrieve_and_free (solver/rk45_api.hpp:63-270), the
 * Radau re-integration of flagged links included when hlm_set_stiff_fallback is on.  A large output
 * leaves in chunks of links whose copies overlap the integration of the next chunk.
 *   y0          host [ns][N_EQ]
 *   tq          host [nq] ascending query times (may be NULL when nq == 0)
 *   out_final   host [ns][N_EQ]; rows of links that did not reach tf are zero (the reference
 *               leaves them unwritten, solver/rk45_kernel.cu:167-175)
 *   out_dense   host [ns][nq][N_EQ] doubles — the order retrieve_and_free returns
 *               (solver/rk45_api.hpp:255-267); slots never reached (tq <= t0, or after a stiff
 *               bail-out) are zero.  May be NULL.  With hlm_set_output_states / hlm_set_output_precision:
 *               [ns][nq][n_selected] of double or float (hlm_output_layout).
 *   out_stiff   host [ns] HLM_LINK_* codes.  May be NULL.
 *   out_n_*     host [ns] accepted / rejected / slope-jump attempt counts.  May be NULL.
 * ns must equal the number of uploaded SpatialParams records for models that use them. */
int hlm_run_rk45(hlm_ctx* ctx, int uid, const double* y0, long long ns, double t0, double tf,
                 const double* tq, long long nq, double* out_final, void* out_dense, int* out_stiff,
                 long long* out_n_accept, long long* out_n_reject, long long* out_n_jump);

/* ---- resident session: the same operator cut into output windows, state left in HBM ---------- */

/* Upload y0 and the query times, reset per-link state (t = t0, h = initialStep, counters 0). */
int hlm_solve_begin(hlm_ctx* ctx, int uid, const double* y0, long long ns, double t0, double tf,
                    const double* tq, long long nq);
/* Start a new interval [t0, tf] from the resident final states: exactly what a second
 * run_rk45(h_y0 = previous final, t0, tf, tq) does (t = t0, h = initialStep, query cursor 0; links
 * flagged stiff/stalled stay flagged; counters keep accumulating) without moving states through
 * the host.  This is how a long run is driven in forcing-sized chunks: the reference's stiffness
 * threshold h < (tf - t0) * 1e-6 (solver/rk45_kernel.cu:160) scales with the interval, so a
 * one-year interval would flag nearly every link, SURVEY §7.3. */
int hlm_solve_restart(hlm_ctx* ctx, double t0, double tf, const double* tq, long long nq);
/* Continue the session from its current end to a later tf with new query times: unlike hlm_solve_restart
 * every link keeps its time and its step size (no ramp-up from initialStep), as if the previous interval
 * had simply gone on.  The stiffness floor h < (tf - t0)*1e-6 (solver/rk45_kernel.cu:160) is taken over
 * the new interval.  Used by routed runs, whose coupling intervals are short. */
int hlm_solve_advance(hlm_ctx* ctx, double tf, const double* tq, long long nq);
/* Advance every unfinished link until it has emitted all queries with index < q_hi (to tf when
 * q_hi >= nq).  Dense records of queries [previous q_hi, q_hi) go to a device buffer
 * [ns][q_hi - q_lo][N_EQ] owned by the context (skipped when want_dense == 0).  Asynchronous. */
int hlm_solve_window(hlm_ctx* ctx, long long q_hi, int want_dense);
/* Device pointer, query range and row pitch of the last window's dense buffer. */
int hlm_solve_window_buffer(hlm_ctx* ctx, void** dev_ptr, long long* q_lo, long long* q_hi);
/* Copy the last window's dense records into the full host array [ns][nq][N_EQ]. Asynchronous if
 * `host_dense` is pinned. */
int hlm_solve_fetch_window(hlm_ctx* ctx, void* host_dense);
/* Copy the last window's dense records packed as [ns][q_hi - q_lo][N_EQ] (the feed of a windowed file
 * writer).  Asynchronous on the copy stream if `host_win` is pinned; `ticket` (may be NULL) names the
 * copy for hlm_solve_wait_copy(), which blocks until that copy has landed (ticket < 0: all queued
 * copies) and may be called from another host thread.  The reference copies the whole dense array with
 * one blocking cudaMemcpy after the run (solver/rk45_api.hpp:173-196). */
int hlm_solve_fetch_window_packed(hlm_ctx* ctx, void* host_win, int* ticket);
int hlm_solve_wait_copy(hlm_ctx* ctx, int ticket);
/* Page-locked host memory for those copies (cudaHostAlloc / cudaFreeHost without linking the CUDA
 * runtime into the host program). */
int hlm_host_alloc(void** out, long long bytes);
int hlm_host_free(void* p);
/* Sums over links, computed on the device: {accepted, rejected, slope-jump, unfinished-active,
 * done, stiff, stalled}.  Synchronises. */
int hlm_solve_totals(hlm_ctx* ctx, long long totals[7]);
/* Download final states and per-link codes/counters (any pointer may be NULL).  Synchronises. */
int hlm_solve_end(hlm_ctx* ctx, double* out_final, int* out_stiff, long long* out_n_accept,
                  long long* out_n_reject, long long* out_n_jump);
/* Accepted implicit (Radau) steps per link so far; they are not part of out_n_accept.  Synchronises. */
int hlm_solve_radau_steps(hlm_ctx* ctx, long long* out_steps);
/* Download the raw resident state (t, h per link) for inspection/tests.  Either may be NULL. */
int hlm_solve_peek(hlm_ctx* ctx, double* out_t, double* out_h, double* out_y);
/* Number of kernel launches issued by this context since creation (bench's gpu_launches). */
long long hlm_launch_count(hlm_ctx* ctx);
/* CUDA-event time, in ms, of the window kernels launched since the last call (sum), and how many.
 * Synchronises the context's stream. */
int hlm_kernel_time_ms(hlm_ctx* ctx, double* sum_ms, long long* n_launches);

/* ---- routed runs: links coupled through upstream channel discharge ----------------------------
 *
 * The reference carries `next_stream` in every parameter record (I_O/parameters_loader.cpp:74,
 * stream.hpp:31,47) and sketches per-step MPI buffers (data/config.yaml:66-70) but couples nothing.
 * Here a routed run advances in coupling intervals (hlm_solve_restart per interval): over an interval every
 * link is integrated on its own with the discharge entering from upstream held at its value at the
 * interval's start, so the RK45 path stays one independent system per thread; between intervals the
 * inflow is re-gathered.  Across ranks only the discharge of boundary links (links whose downstream
 * link another rank owns) is exchanged: the window kernel's epilogue writes it into the send buffer,
 * the caller runs its collective (NCCL all-gather; MPI in the reference's world) on the same stream and
 * hands the gathered halo vector to hlm_route_gather.
 *
 * Topology is CSR over the links this context owns, in the order of hlm_solve_begin's y0: the upstream
 * links of link i are up_idx[up_ptr[i] .. up_ptr[i+1]).  An entry e >= 0 is a local link; e < 0 is
 * element -(e+1) of the halo vector.  Entries are added in the order given (give them by ascending
 * global link id and the sums are the same under every partition).  send_idx[n_send] lists the local
 * links other ranks need, in the order of this rank's segment of the halo vector.  Copies its inputs. */
int hlm_route_set_topology(hlm_ctx* ctx, const long long* up_ptr, const int* up_idx, long long ns,
                           const int* send_idx, long long n_send);
int hlm_route_clear(hlm_ctx* ctx);
/* Where boundary discharge is written: NULL = a buffer of the context's own, else a device buffer of the
 * caller's (n_send doubles, e.g. the input of its all-gather). */
int hlm_route_set_send_buffer(hlm_ctx* ctx, double* dev_buf);
int hlm_route_send_buffer(hlm_ctx* ctx, void** dev_ptr, long long* n_send);
/* Fill the send buffer from the resident state.  Needed before the first interval only: afterwards the
 * window kernel has already done it. */
int hlm_route_pack(hlm_ctx* ctx);
/* inflow[i] = sum of upstream discharge for the next interval; dev_halo (device pointer, at least as long as
 * the largest halo slot up_idx refers to) may be NULL only when no entry of up_idx is negative.  Queued on
 * the context's stream. */
int hlm_route_gather(hlm_ctx* ctx, const double* dev_halo);
/* Peer-memory exchange, for ranks that are processes on ONE node (NVLink/NVSwitch): instead of a send buffer that
 * a collective moves, every rank's halo vector is mapped into every other rank's address space (CUDA IPC) and the
 * integration kernel's epilogue stores each boundary link's discharge straight into all of them — the transfer is
 * part of the compute kernel, and what remains between intervals is a barrier (any collective on the stream) so
 * that every rank's stores have landed before hlm_route_gather reads.  Two parities alternate, so a fast rank
 * never overwrites what a slow one still reads.  Sequence: every rank hlm_route_peer_alloc (gets a 64-byte IPC
 * handle), the handles are all-gathered by the host ([world][64] bytes, rank order), every rank
 * hlm_route_peer_open; from then on hlm_route_pack and the kernels publish to the peers and hlm_route_gather
 * ignores its argument and reads the rank's own halo vector.  hlm_route_peer_close (after a barrier: peers may
 * still be storing) returns to the send-buffer scheme.  max_send = the plan's segment length. */
int hlm_route_peer_alloc(hlm_ctx* ctx, int world, int rank, long long max_send, void* ipc_handle_out);
int hlm_route_peer_open(hlm_ctx* ctx, const void* handles);
int hlm_route_peer_close(hlm_ctx* ctx);
/* Download the current inflow [ns] and send buffer [n_send] (either may be NULL).  Synchronises. */
int hlm_route_peek(hlm_ctx* ctx, double* out_qin, double* out_send);

/* ---- measurement helpers --------------------------------------------------------------------- */

/* Register-resident DFMA / FFMA microbenchmark: achieved FMA throughput of this device in
 * TFLOP/s (2 flops per FMA).  Used as the roofline denominator, which MEASURED_PEAKS.json lacks
 * for FP64/FP32. bits = 64 or 32. */
int hlm_measure_fma_peak(hlm_ctx* ctx, int bits, double* tflops);

/* Element-wise probe of the device's arithmetic on host arrays x, y -> out (n elements):
 * op 0 = libdevice pow(x, y); 1 = rcp.approx.ftz.f64(x) (the MUFU.RCP64H seed pow starts from);
 * 2 = x / y (div.rn.f64); 3 = sqrt.rn.f64(x); 4 = the kernels' inlined pow (fp_exact.cuh pow_pos);
 * 5 = Model 200's x^(1/5), 6 = its x^(2/3) (fp_exact.cuh root5 / cbrt2; y unused).
 * Lets tests compare device and host arithmetic. */
int hlm_debug_eval(hlm_ctx* ctx, int op, const double* x, const double* y, double* out, long long n);

#ifdef __cplusplus
}
#endif
#endif /* HLM_B200_H */
