/*
 * svb200.h -- C ABI of libsvb200.so: B200-native (sm_100a) kernels for the batched Villain /
 * worldline Metropolis hot path of evanberkowitz/supervillain.
 *
 * The reference has no FFI of its own for this path: the seam is the duck-typed Python protocol
 * `generator.step(cfg)` / `generator.inline_observables(steps)` / `generator.report()` driven by
 * supervillain/ensemble.py:74-98, plus the operator functions `d`, `delta`, `Form.face_sum`,
 * `Form.coface_sum` of supervillain/lattice/compact.py.  The entry points below are what a
 * ctypes/cffi binding on the reference side would call to replace the bodies of those Python
 * functions; each one names the reference function it stands in for.  INTEGRATION.md shows the
 * binding.
 *
 * Conventions
 *  - every function returns int: 0 = ok, <0 = SVB_E_* (bad argument), >0 = a cudaError_t;
 *    nothing throws; svb_last_error() returns a thread-local message for the last failure.
 *  - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *    buffer; the library allocates nothing and keeps no global state besides kernel attributes.
 *  - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL is the
 *    legacy default stream); no implicit synchronisation.
 *  - field layout is the reference's compact layout with a leading chain axis, C-contiguous:
 *        phi (chains, 1, N, N)   n, m (chains, 2, N, N)   v (chains, 1, N, N)
 *    (supervillain/lattice/compact.py:665-716, supervillain/action/villain.py:76-91,
 *    supervillain/action/worldline.py:96-114).  Axis -2 is lattice direction 0.
 *  - D = 2 only (every configuration named by BASELINE.json).
 */
#ifndef SVB200_H
#define SVB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_VERSION_MAJOR 0
#define SVB_VERSION_MINOR 1

/* error codes (negative); positive return values are cudaError_t */
#define SVB_OK              0
#define SVB_E_NULL         -1   /* a required pointer is NULL */
#define SVB_E_SHAPE        -2   /* chains/N out of range */
#define SVB_E_DTYPE        -3   /* unsupported dtype code */
#define SVB_E_PARAM        -4   /* bad scalar parameter (kappa<=0, W<1, interval...) */
#define SVB_E_UNSUPPORTED  -5   /* valid request this build cannot serve (e.g. odd N in a tiled path) */
#define SVB_E_ALIGN        -6   /* pointer alignment */

/* dtype codes */
#define SVB_F64 0
#define SVB_F32 1
#define SVB_I32 2
#define SVB_I64 3

/* rng modes */
#define SVB_RNG_PHILOX   0   /* in-kernel Philox4x32-10, counter = (site, chain, sweep) keyed by seed */
#define SVB_RNG_INJECTED 1   /* proposals and uniforms read from caller-supplied dense arrays */

/* arithmetic modes */
#define SVB_ARITH_STRICT 0   /* the reference's operation order, no FMA contraction (bit-faithful dS) */
#define SVB_ARITH_FAST   1   /* FMA-contracted dS (agrees with STRICT to ~1e-15 relative) */

/* code paths for the sweeps (SVB_PATH_AUTO picks by N) */
#define SVB_PATH_AUTO    0
#define SVB_PATH_SMEM    1   /* whole lattice of a chain resident in shared memory, all colours and
                                n_sweeps sweeps per launch: one HBM read + one write per launch */
#define SVB_PATH_GLOBAL  2   /* one launch per colour pass straight out of HBM/L2 (any N) */

/* per-chain observable record written by the sweep / observable entry points */
#define SVB_VOBS_ACTION        0   /* S = kappa/2 sum (dphi - 2 pi n)^2       (action/villain.py:51-66) */
#define SVB_VOBS_SUM_DN2       1   /* sum_p (dn)_p^2   -> WindingSquared = /N^2 (observable/winding.py:30-37) */
#define SVB_VOBS_WRAP0         2   /* sum_x n_0[x]      TorusWrapping[0]        (observable/wrapping.py:17-25) */
#define SVB_VOBS_WRAP1         3   /* sum_x n_1[x]      TorusWrapping[1] */
#define SVB_VOBS_ACCEPTED      4   /* accepted proposals over the sweeps of this call (neighborhood.py:117,133) */
#define SVB_VOBS_ACCEPTANCE    5   /* sum over proposals of min(1, e^-dS)         (neighborhood.py:118,132).
                                      A MONITOR, not an observable: the FAST kernels (Philox draws, fp32-filtered decisions)
                                      accumulate it from the fp32 ex2.approx estimate of e^-dS, so it agrees with the fp64
                                      sum to ~1e-5 relative (the tests state rtol = 1e-5) -- outside the 1e-12 that holds for
                                      ACTION .. WRAP1 and for every accept/reject decision (those are exact).  STRICT
                                      arithmetic and injected draws evaluate it in fp64. */
#define SVB_VOBS_COUNT         6

#define SVB_WOBS_SUM_F2        0   /* sum_l (m - delta v / W)_l^2  (action/worldline.py:94; observable/action.py:35-47) */
#define SVB_WOBS_SUM_DF2       1   /* sum_p (d(m - delta v/W))_p^2 (observable/winding.py:40-52) */
#define SVB_WOBS_WRAP0         2   /* sum_x m_0[x]   (TorusWrapping = /N, observable/wrapping.py:28-39) */
#define SVB_WOBS_WRAP1         3
#define SVB_WOBS_ACCEPTED      4
#define SVB_WOBS_ACCEPTANCE    5   /* as SVB_VOBS_ACCEPTANCE: an fp32-accurate monitor on the W = 1 Philox kernels */
#define SVB_WOBS_DELTA_M_ABS   6   /* sum_x |(delta m)[x]|: 0 iff the constraint holds (action/worldline.py:54-70);
                                      evaluated by svb_worldline_observables; a sweep may report -1 = not evaluated
                                      (the move preserves delta m identically) */
#define SVB_WOBS_COUNT         7

/* worldline sweep modes */
#define SVB_WL_JOINT    0   /* (dm, dv) jointly per plaquette: PlaquetteUpdate's move (worldline/plaquette.py:79-101) */
#define SVB_WL_VORTEX   1   /* v only:  VortexUpdate  (worldline/vortex.py:51-136) */
#define SVB_WL_COEXACT  2   /* m only (m += delta t): CoexactUpdate (worldline/coexact.py:53-128) */

/* form operators */
#define SVB_OP_D          0   /* supervillain.lattice.d            compact.py:973-1001  */
#define SVB_OP_DELTA      1   /* supervillain.lattice.delta        compact.py:1008-1037 */
#define SVB_OP_FACE_SUM   2   /* Form.face_sum                     compact.py:848-867   */
#define SVB_OP_COFACE_SUM 3   /* Form.coface_sum                   compact.py:869-890   */

int         svb_version(void);
const char* svb_last_error(void);

/*
 * Replaces the body of NeighborhoodUpdate.step (supervillain/generator/villain/neighborhood.py:59-137)
 * for `chains` independent chains at once, `n_sweeps` sweeps per call, IN PLACE.
 *
 *  phi            (chains,1,N,N) SVB_F64 or SVB_F32
 *  n              (chains,2,N,N) SVB_I32
 *  kappa          coupling used when kappa_chain == NULL
 *  kappa_chain    optional (chains,) f64: one coupling per chain (kappa scans)
 *  W              constraint integer (finite, >= 1): dn proposals are W * [-interval_n, interval_n]
 *  seed, sweep0, chain0   Philox key / counter offsets: chain c of this call, sweep s of this call
 *                 draws from counter (site, chain0 + c, sweep0 + s); results do not depend on
 *                 how chains are split over calls or GPUs.  Every entry point (generator kind) draws from its own
 *                 Philox stream pair, so the same seed may be given to all of them.
 *  interval_n     in [0, 31].  0 and 1 run on the fp32-filtered kernels; >= 2 on the generic fp64 kernels (there the
 *                 Metropolis uniform's leading bits come from a second Philox block, independent of the dn digits)
 *  rng_mode       SVB_RNG_PHILOX, or SVB_RNG_INJECTED with
 *                   inj_u, inj_dphi  (n_sweeps, chains, N, N) f64
 *                   inj_dn_fwd/bwd   (n_sweeps, chains, 2, N, N) i32, indexed by the PROPOSING site
 *                                    and already multiplied by W  
 *  obs            optional (chains, SVB_VOBS_COUNT) f64, state after the last sweep + counters
 *  accept_mask    optional (chains, N, N) u8: accept decisions of the LAST sweep (parity tests)
 *  dS_out         optional (chains, N, N) f64: dS of the LAST sweep's proposals (parity tests)
 */
int svb_villain_sweep(void* phi, int phi_dtype, int32_t* n,
                      int64_t chains, int N,
                      double kappa, const double* kappa_chain, int W,
                      double interval_phi, int interval_n,
                      int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                      int rng_mode, int arith_mode, int path,
                      const double* inj_u, const double* inj_dphi,
                      const int32_t* inj_dn_fwd, const int32_t* inj_dn_bwd,
                      double* obs, uint8_t* accept_mask, double* dS_out,
                      void* stream);

/*
 * The same sweeps for lattices too large for shared memory (N a multiple of 32, e.g. configs 4 and 5): one CTA per
 * 32 x 32 tile with a ghost zone, one pass per sweep, ping-pong between (phi, n) and a caller-provided workspace
 * (phi_ws, n_ws) of the same shapes.  The result always ends in (phi, n).  fp64 phi, Philox mode.
 */
#define SVB_PATH_TILED 3
int svb_villain_sweep_tiled(void* phi, int32_t* n, void* phi_ws, int32_t* n_ws,
                            int64_t chains, int N, double kappa, const double* kappa_chain, int W,
                            double interval_phi, int interval_n, int n_sweeps,
                            uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode,
                            double* obs, uint8_t* accept_mask, double* dS_out, void* stream);

/*
 * svb_villain_sweep_tiled for a caller that owns both buffer pairs and exchanges their roles (a resident ensemble stepping
 * one sweep at a time; the reference's NeighborhoodUpdate.step, neighborhood.py:59-137, returns fresh arrays anyway): after
 * an odd number of FAST sweeps the state is left in (phi_ws, n_ws) and *state_in_workspace = 1 (host int, written before
 * the call returns), sparing the copy back; otherwise the state is in (phi, n) and *state_in_workspace = 0.
 */
int svb_villain_sweep_tiled_swap(void* phi, int32_t* n, void* phi_ws, int32_t* n_ws,
                                 int64_t chains, int N, double kappa, const double* kappa_chain, int W,
                                 double interval_phi, int interval_n, int n_sweeps,
                                 uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode,
                                 double* obs, int* state_in_workspace, void* stream);

/*
 * The same sweeps IN PLACE for lattices beyond a CTA (config 5: one L = 4096 lattice; any N that is a multiple of 16): one
 * launch per colour pass, the fields stay where they are, only accepted proposals are written (svb_villain_stream.cuh; the tile
 * of a CTA is staged in shared memory by TMA tensor loads when N is a multiple of 128).  No workspace.  Philox draws, fp64 phi,
 * FAST arithmetic, interval_n <= 1.
 *  obs            optional (chains, SVB_VOBS_COUNT).  With obs_in == NULL: the full record of the state AFTER the sweeps
 *                 (one more read of the state) and this call's ACCEPTED / ACCEPTANCE.
 *  obs_in         optional (needs obs): as in svb_villain_sweep_overlapped -- the state columns ACTION .. WRAP1 of the lattice AS IT
 *                 ARRIVES are written to obs_in (action and wrapping ride along with the first colour pass, sum (dn)^2 is one pass
 *                 over n), its counter columns are left alone, and `obs` receives only this call's ACCEPTED and ACCEPTANCE.
 */
int svb_villain_sweep_inplace(void* phi, int32_t* n, int64_t chains, int N,
                              double kappa, const double* kappa_chain, int W,
                              double interval_phi, int interval_n,
                              int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                              double* obs, double* obs_in, void* stream);

/*
 * svb_villain_sweep_inplace as ONE launch per step (N a multiple of 128; configs 4 and 5): the phases of a step -- sum (dn)^2 of
 * the arriving state, colour 0, colour 1 [, the colours of further sweeps] -- follow one another down the lattice a few tile
 * rows apart, so each phase reads from L2 what the one before left there, and DRAM sees one read of the state per launch plus
 * the sectors that changed (the three launches of svb_villain_sweep_inplace read it 2.5 times).  Same proposals, same
 * decisions, same fields bit for bit (NeighborhoodUpdate.step, generator/villain/neighborhood.py:59-137); obs / obs_in as above.
 * On a B200 it takes the same time as svb_villain_sweep_inplace without a record and is slower with one (the colour passes are
 * bound by instruction issue, not by DRAM: DESIGN 3.3), so the host layer uses it only on request.
 *
 *  workspace      int32 device scratch owned by the caller, at least svb_villain_wavefront_workspace(chains, N, 1, obs_in != NULL)
 *                 ints (with room for k sweeps, k sweeps share a launch); it must be ALL ZERO before its first use and every
 *                 call leaves it all zero.  Calls that share a workspace must be ordered (one stream).
 */
long long svb_villain_wavefront_workspace(int64_t chains, int N, int n_sweeps, int with_obs_in);
int svb_villain_sweep_wavefront(void* phi, int32_t* n, int64_t chains, int N,
                                double kappa, const double* kappa_chain, int W,
                                double interval_phi, int interval_n,
                                int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                                double* obs, double* obs_in, int32_t* workspace, int64_t workspace_ints, void* stream);

/*
 * The same sweeps as OVERLAPPED launches: a launch may begin while the previous launch in the stream is still running
 * (programmatic dependent launch), so the ramp-up of one sweep hides under the tail of the one before -- at config 2 a
 * quarter of a non-overlapped step.  Data dependencies are tracked per chain instead of per kernel:
 *
 *  epochs         (chains,) u32 device scratch owned by the caller, one word per chain of THIS chain set.  A launch loads
 *                 chain c only once epochs[c] == wait_epoch, and sets epochs[c] = signal_epoch when its store of chain c
 *                 and the chain's observable record are complete.  A sequence of steps on one chain set therefore passes
 *                 (wait, signal) = (e, e + 1), (e + 1, e + 2), ...
 *  flags          0: the launch first waits for ALL earlier work in the stream (griddepcontrol.wait) and ignores
 *                 wait_epoch -- use it for the first launch of a sequence, whose inputs were produced by anything else.
 *                 SVB_OVERLAP_PREDECESSOR: the caller vouches that everything this launch reads was written either
 *                 before the last launch that passed flags = 0 on this stream, or by overlapped launches (whose chains
 *                 the epochs order).  Only then does the launch skip the grid-wide wait and overlap its predecessor.
 *                 Launches on different chain sets may be interleaved freely; each set has its own epochs.
 *  obs            optional (chains, SVB_VOBS_COUNT); give consecutive launches DIFFERENT records (e.g. rows of a
 *                 (steps, chains, SVB_VOBS_COUNT) column) if they may overlap.
 *  obs_in         optional (chains, SVB_VOBS_COUNT).  NULL: `obs` is the full record of the state AFTER the sweeps (one
 *                 more fp64 pass over the chain).  Non-NULL: the state columns ACTION .. WRAP1 of the chain AS IT ARRIVES
 *                 are written to obs_in (they ride along with the pass that builds the residuals, for a quarter of the
 *                 cost), its two counter columns are left alone, and `obs` receives only this launch's ACCEPTED and
 *                 ACCEPTANCE.  In a sequence of steps pass obs_in = the previous step's record: every record is then
 *                 complete one launch later, with exactly the bits the separate pass would have produced; the last
 *                 state's columns come from svb_villain_observables.
 * Philox draws, fp64 phi, FAST arithmetic, N in {16, 32, 64, 128}, interval_n <= 1; anything else returns
 * SVB_E_UNSUPPORTED (use svb_villain_sweep).  A kernel launched normally after these waits for all of them, as usual.
 */
#define SVB_OVERLAP_PREDECESSOR 1
int svb_villain_sweep_overlapped(void* phi, int32_t* n, int64_t chains, int N,
                                 double kappa, const double* kappa_chain, int W,
                                 double interval_phi, int interval_n,
                                 int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                                 double* obs, double* obs_in, uint32_t* epochs, uint32_t wait_epoch, uint32_t signal_epoch,
                                 int flags, void* stream);

/*
 * The decoupled Villain updates: SiteUpdate.step (generator/villain/site.py:43-120: phi alone, checkerboard),
 * LinkUpdate.step (link.py:53-101: every n independently against the frozen phi) and ExactUpdate.step
 * (exact.py:50-129: n += d z, z on one colour at a time), IN PLACE, fp64 phi, STRICT arithmetic.
 *  interval_phi   SITE: dphi ~ uniform(-interval_phi, +interval_phi)
 *  interval       LINK: dn in W * ([-interval, interval] \ {0});  EXACT: z in [-interval, interval] \ {0};  1 <= interval <= 128
 *  injected       inj_u (n_sweeps, chains, N, N) f64 (LINK: (n_sweeps, chains, 2, N, N));
 *                 SITE: inj_dphi (n_sweeps, chains, N, N) f64;  LINK: inj_a (n_sweeps, chains, 2, N, N) i32 = the
 *                 proposed change, already times W;  EXACT: inj_a (n_sweeps, chains, N, N) i32 = z
 *  obs            optional (chains, SVB_VOBS_COUNT): state after the last sweep + ACCEPTED / ACCEPTANCE of the call
 *  accept_mask / dS_out   optional, last sweep, shaped like inj_u (no accept_mask for LINK)
 */
#define SVB_VU_SITE  0
#define SVB_VU_LINK  1
#define SVB_VU_EXACT 2
int svb_villain_decoupled(int kind, void* phi, int32_t* n, int64_t chains, int N,
                          double kappa, const double* kappa_chain, int W,
                          double interval_phi, int interval,
                          int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                          int rng_mode, int path,
                          const double* inj_u, const double* inj_dphi, const int32_t* inj_a,
                          double* obs, uint8_t* accept_mask, double* dS_out, void* stream);

/*
 * CohomologyUpdate.step (generator/villain/cohomology.py:64-117): per chain and direction mu one proposal h in
 * [-interval, interval] \ {0} added to n_mu on the whole slice x_mu = 0 (changes the winding sector), IN PLACE on n.
 *  injected       inj_u (chains, 2) f64 and inj_h (chains, 2) i32, [mu] in the reference's draw order (h, u per direction)
 *  counters       optional (chains, 2) f64, ACCUMULATED: accepted, sum of min(1, e^-dS);  dS_out optional (chains, 2)
 */
int svb_villain_cohomology(const void* phi, int32_t* n, int64_t chains, int N,
                           double kappa, const double* kappa_chain, int interval,
                           uint64_t seed, uint64_t sweep, uint64_t chain0, int rng_mode,
                           const double* inj_u, const int32_t* inj_h,
                           double* counters, double* dS_out, void* stream);

/*
 * The same sweep with HOST buffers: the reference's `step(cfg)` contract (host arrays in, host arrays out,
 * neighborhood.py:59-137) for a whole batch.  phi_host / n_host / obs_host are pinned HOST buffers updated in place;
 * phi_dev / n_dev / obs_dev are caller-owned DEVICE staging buffers of the same shapes.  The chains are processed in
 * `n_chunks` chunks round-robin over the `n_streams` streams (cudaStream_t[]), each chunk H2D -> sweep -> D2H, so copies
 * in both directions overlap the kernels.  Philox mode only.  Asynchronous: the caller synchronises the streams.
 */
int svb_villain_sweep_host(void* phi_host, int phi_dtype, int32_t* n_host, double* obs_host,
                           void* phi_dev, int32_t* n_dev, double* obs_dev,
                           int64_t chains, int N, double kappa, int W, double interval_phi, int interval_n,
                           int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0, int arith_mode,
                           int n_chunks, void* const* streams, int n_streams);

/*
 * Villain.__call__ (action/villain.py:51-66) and the scalar observables of
 * observable/{action,energy,winding,wrapping}.py for every chain: fills obs[:, ACTION..WRAP1]
 * and zeroes the two counters.
 */
int svb_villain_observables(const void* phi, int phi_dtype, const int32_t* n,
                            int64_t chains, int N, double kappa, const double* kappa_chain,
                            double* obs, void* stream);

/*
 * The worldline checkerboard sweep: the PlaquetteUpdate move (worldline/plaquette.py:79-101) in
 * the red/black order of VortexUpdate/CoexactUpdate (worldline/vortex.py:86-128,
 * worldline/coexact.py:91-120), IN PLACE on m (chains,2,N,N) i32 and v (chains,1,N,N) i32.
 *  mode           SVB_WL_JOINT / SVB_WL_VORTEX / SVB_WL_COEXACT
 *  interval       proposals are drawn from [-interval, interval] \ {0} for dv (VORTEX) and t
 *                 (COEXACT), 1 <= interval <= 128; JOINT is the reference's dm in {-1,+1}, dv in {-1,0,+1}
 *  injected       inj_u (n_sweeps, chains, N, N) f64; inj_a (n_sweeps, chains, N, N) i32 = dm (JOINT),
 *                 dv (VORTEX) or t (COEXACT); inj_b same shape = dv (JOINT only)
 */
int svb_worldline_sweep(int32_t* m, int32_t* v,
                        int64_t chains, int N,
                        double kappa, const double* kappa_chain, int W,
                        int mode, int interval,
                        int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                        int rng_mode, int path,
                        const double* inj_u, const int32_t* inj_a, const int32_t* inj_b,
                        double* obs, uint8_t* accept_mask, double* dS_out,
                        void* stream);

/*
 * The worldline sweeps (W = 1, Philox draws, N in {16, 32, 64, 128}; interval 1 or 2 for VORTEX / COEXACT) as OVERLAPPED
 * launches: the same protocol as svb_villain_sweep_overlapped (epochs, wait_epoch / signal_epoch,
 * SVB_OVERLAP_PREDECESSOR); `obs` is the full record of the state after the sweeps.
 */
int svb_worldline_sweep_overlapped(int32_t* m, int32_t* v, int64_t chains, int N,
                                   double kappa, const double* kappa_chain, int mode, int interval,
                                   int n_sweeps, uint64_t seed, uint64_t sweep0, uint64_t chain0,
                                   double* obs, uint32_t* epochs, uint32_t wait_epoch, uint32_t signal_epoch,
                                   int flags, void* stream);

/*
 * WrappingUpdate.step (worldline/wrapping.py:43-90): one proposal per torus cycle (2N per chain), IN PLACE on m.
 *  interval       proposals are drawn from [-interval, interval] \ {0}
 *  sweep          Philox counter of this step (stream 3, counter word 0 = mu * N + k)
 *  injected       inj_u (chains, 2, N) f64 and inj_c (chains, 2, N) i32: [mu, k] is the mu-direction cycle at
 *                 perpendicular coordinate k (the reference's draw order: choices mu=0, mu=1, uniforms mu=0, mu=1)
 *  counters       optional (chains, 2) f64: accepted, sum of min(1, e^-dS);  dS_out optional (chains, 2, N)
 */
int svb_worldline_wrapping(int32_t* m, const int32_t* v, int64_t chains, int N,
                           double kappa, const double* kappa_chain, int W, int interval,
                           uint64_t seed, uint64_t sweep, uint64_t chain0, int rng_mode,
                           const double* inj_u, const int32_t* inj_c,
                           double* counters, double* dS_out, void* stream);

/* Worldline.__call__ ingredients and observables (action/worldline.py:72-94): obs (chains, SVB_WOBS_COUNT) */
int svb_worldline_observables(const int32_t* m, const int32_t* v,
                              int64_t chains, int N, int W,
                              double* obs, void* stream);

/*
 * d / delta / face_sum / coface_sum on `chains` forms of input degree `degree`, dtype-preserving
 * (compact.py:954-966 -> lattice/_kernels.py:19-46).  in (chains, C(2,degree), N, N) ->
 * out (chains, C(2,degree +/- 1), N, N).  Returns SVB_E_PARAM at the ends of the complex, where
 * the reference returns the scalar 0.
 */
int svb_form_op(int op, int degree, int dtype, const void* in, void* out,
                int64_t chains, int N, void* stream);

/*
 * Spin_Spin.Villain (observable/spin.py:28-42): C[dx] = N^-2 sum_x e^{-i phi_x} e^{+i phi_{x-dx}},
 * out (chains, N, N, 2) f64 (re, im), 16-byte aligned.  FFT for power-of-two N from 16 to 4096 (`out` is the workspace for
 * N >= 128), direct O(N^4) evaluation out of shared memory for the other N that fit it.
 */
int svb_villain_spin_spin(const void* phi, int phi_dtype, int64_t chains, int N,
                          double* out, void* stream);

/*
 * Lattice.correlation (compact.py:465-536) of a per-site field derived from `field`:
 *   SVB_CORR_SPIN     field = phi (chains,1,N,N) f64|f32, s = exp(i phi)          Spin_Spin.Villain
 *   SVB_CORR_WINDING  field = n   (chains,2,N,N) i32,     s = dn                  Winding_Winding.Villain (winding.py:77-86)
 *   SVB_CORR_VORTEX   field = v   (chains,1,N,N) i32,     s = exp(2 pi i v / W)   Vortex_Vortex.Worldline (vortex.py:22-37)
 * out (chains, N, N, 2) f64 (re, im).  As svb_villain_spin_spin: FFT for power-of-two N in [16, 4096], else the direct sum.
 */
#define SVB_CORR_SPIN    0
#define SVB_CORR_WINDING 1
#define SVB_CORR_VORTEX  2
int svb_correlation(int kind, const void* field, int dtype, int64_t chains, int N, int W, double* out, void* stream);

/*
 * supervillain.analysis.autocorrelation (analysis/autocorrelation.py:7-66) for `series` scalar columns of length T at
 * once (e.g. one observable of every chain): data (series, T) f64 -> C (series, T) f64, the circular autocorrelation
 * function normalised to C(0) = 1, and tau (series,) i32, the ceiling of the integrated autocorrelation time up to the
 * first zero of C.  mean: optional (series,) imposed means (NULL: computed from the data).  A series without
 * fluctuations (|C(0)| < 1e-16 before normalisation, where the reference raises) gets tau = -1.  T <= 12800.
 */
int svb_autocorrelation(const double* data, int64_t series, int T, const double* mean, double* C, int32_t* tau, void* stream);

/*
 * The taxicab reweighting observables (D = 2): Spin_Spin.Worldline (observable/spin.py:50-224) from links = m - delta(v)/W
 * and Vortex_Vortex.Villain (observable/vortex.py:63-189) from links = d(phi) - 2 pi n.  links (chains, 2, N, N) f64 ->
 * out (chains, N, N) f64 indexed by the displacement in FFT coordinates, out[., 0, 0] = 1, averaged over all starting sites.
 * N (N + 1) 16 bytes of shared memory: N <= 118.
 */
#define SVB_TAXI_SPIN   0
#define SVB_TAXI_VORTEX 1
int svb_taxicab_correlator(int kind, const double* links, int64_t chains, int N, double kappa, const double* kappa_chain, double* out,
                           void* stream);

/*
 * Blocking and Bootstrap of scalar columns (analysis/blocking.py:54-66 `_block`, analysis/bootstrap.py:57-67 `_resample`),
 * one column of T samples per series (e.g. one observable of every chain): data (series, T) f64, weight optional (T,) f64
 * (NULL: unit weights, what every generator on this path produces).
 *   svb_block_mean      out (series, (T - drop) / width):  mean over `width` consecutive samples of weight * data, the first
 *                       `drop` samples left out; (T - drop) must be a multiple of width.
 *   svb_bootstrap_mean  out (series, draws):  mean_c(weight[idx[c,d]] data[s, idx[c,d]]) / mean_c(weight[idx[c,d]]) with
 *                       idx (T, draws) int64 -- the resampling indices, drawn by the caller (the reference draws them with
 *                       numpy: np.random.randint(0, T, (T, draws)), bootstrap.py:52).
 */
int svb_block_mean(const double* data, const double* weight, int64_t series, int64_t T, int width, int64_t drop, double* out, void* stream);
int svb_bootstrap_mean(const double* data, const double* weight, int64_t series, int64_t T, const int64_t* idx, int draws, double* out,
                       void* stream);

/*
 * Test hook: the decision u < A of the lazily refined Metropolis uniform (leading 32 bits f known; trailing bits from word
 * `word` of the Philox block (c0, chain, sweep) in refinement stream `stream_id`: 4 NeighborhoodUpdate, 5 PlaquetteUpdate,
 * 7 LinkUpdate, 10 SiteUpdate, 12 ExactUpdate, 14 VortexUpdate, 16 CoexactUpdate) and
 * the refined uniform itself, for n caller-chosen inputs.  The sweeps reach the refinement with probability 2^-32 per
 * proposal; this makes it testable against the oracle.
 */
int svb_debug_decide_lazy(const double* A, const uint32_t* f, const uint32_t* c0, const uint32_t* word, int64_t n,
                          uint32_t stream_id, uint64_t seed, uint64_t chain, uint64_t sweep,
                          uint8_t* decision, double* u_out, void* stream);

/* Philox4x32-10 block, exposed for known-answer tests: out[4] = philox(ctr[4], key[2]) (host). */
void svb_philox4x32_10_host(const uint32_t* ctr_host, const uint32_t* key_host, uint32_t* out_host);

/* The draw mapping used in SVB_RNG_PHILOX mode, evaluated on the device for
 * (chain0 + c, sweep, site) so tests can compare against the oracle's restatement:
 * u, dphi (chains, N, N) f64;  dn (chains, 4, N, N) i32 ordered (fwd0, bwd0, fwd1, bwd1). */
int svb_villain_draws(int64_t chains, int N, int W, double interval_phi, int interval_n,
                      uint64_t seed, uint64_t sweep, uint64_t chain0,
                      double* u, double* dphi, int32_t* dn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVB200_H */
