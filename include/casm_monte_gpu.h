/* casm_monte_gpu.h -- C ABI of the B200-native Ising SGC Metropolis path.
 *
 * This is the drop-in boundary for ONE hot path of libcasm-monte 2.2.0: the
 * semi-grand-canonical Ising Metropolis loop and the sampling / statistics it
 * feeds.  The reference has no FFI of its own for this path (it is a set of
 * header templates bound with pybind11), so each entry point cites the
 * reference interface it stands behind (paths relative to the reference root).
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns CMG_OK (0) or a negative CMG_E* code, never throws;
 *     cmg_last_error() gives the message for the last failure on that context,
 *     cmg_last_global_error() for failures without a context (create).
 *   - one host thread per context.  Work is enqueued on the context's CUDA
 *     stream (cmg_set_stream); functions that return host data synchronise that
 *     stream, the others are asynchronous unless stated.
 *   - occupation crosses the boundary in the reference's representation:
 *     int32, values +1/-1, column-major  l = i + n0*(j + n1*k)
 *     (include/casm/monte/ising_cpp/model.hh:48, :82-99).  On the device it is
 *     held as two int8 checkerboard colour planes.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     CMG_ENODEVICE.
 */
#ifndef CASM_MONTE_GPU_H
#define CASM_MONTE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMG_ABI_VERSION 1

/* status codes */
#define CMG_OK 0
#define CMG_EINVAL -1     /* bad argument (also: odd extent in checkerboard mode) */
#define CMG_ENODEVICE -2  /* no CUDA device / driver */
#define CMG_ECUDA -3      /* a CUDA runtime call failed */
#define CMG_ENOMEM -4
#define CMG_ESTATE -5     /* call sequence error (e.g. conditions not set) */
#define CMG_EUNSUPPORTED -6

/* update modes for cmg_run_passes */
#define CMG_MODE_CHECKERBOARD 0     /* production: two coloured half-sweeps per pass, Philox4x32-10 */
#define CMG_MODE_SERIAL_REFERENCE 1 /* validation: the reference's serial random-site loop on the
                                       reference's std::mt19937_64 stream, trajectory-exact */

/* sampled quantities (the reference's default sampling functions,
 * include/casm/monte/ising_cpp/basic_semigrand_canonical.hh:486-590) */
#define CMG_Q_PARAM_COMPOSITION 0
#define CMG_Q_FORMATION_ENERGY 1
#define CMG_Q_POTENTIAL_ENERGY 2

/* Boltzmann constant used for beta = 1/(KB*T)
 * (include/casm/monte/methods/basic_occupation_metropolis.hh:361; the value
 * lives in CASMcode_global's casm/global/definitions.hh). */
#define CMG_KB 8.6173303E-05

typedef struct cmg_context cmg_context;

/* ---- library ------------------------------------------------------------ */
int cmg_abi_version(void);
const char *cmg_last_global_error(void);
const char *cmg_last_error(const cmg_context *ctx);
int cmg_device_count(int *count);

/* ---- lifecycle ------------------------------------------------------------
 * A context holds n_chains independent lattices of the same shape (one Markov
 * chain each, e.g. the points of a (T, mu) grid) on one device.
 * Replaces: IsingConfiguration(shape) + IsingState
 * (include/casm/monte/ising_cpp/model.hh:19-34, :141-157); dim is 2 (the
 * reference) or 3 (simple cubic, extension).  All lattices start filled +1.
 * Extents must be >= 2; checkerboard mode additionally needs even extents. */
int cmg_create(int dim, const int64_t *shape, int n_chains, int device,
               cmg_context **out);
int cmg_destroy(cmg_context *ctx);
/* cuda_stream is a cudaStream_t (e.g. torch.cuda.Stream.cuda_stream); 0 = legacy default */
int cmg_set_stream(cmg_context *ctx, void *cuda_stream);
int cmg_sync(cmg_context *ctx);
int cmg_n_sites(const cmg_context *ctx, int64_t *n_sites);

/* ---- slab decomposition (one context per GPU, column slabs along the slowest
 * axis).  Call instead of cmg_create: the context owns columns
 * [col_begin, col_begin + n_cols_local) of a global lattice and two halo
 * columns per colour plane.  Philox counters are keyed on GLOBAL site indices,
 * so the decomposed trajectory is bit-identical to the single-GPU one. */
int cmg_create_slab(int dim, const int64_t *global_shape, int64_t col_begin,
                    int64_t n_cols_local, int device, cmg_context **out);
/* Halo plumbing.  The halo columns live in device memory owned by the context;
 * side 0 = low neighbour, 1 = high neighbour.  get_boundary/ set_halo give the
 * host/driver (NCCL send/recv through torch.distributed, or a gloo test) a
 * device pointer + byte count for the freshly updated boundary column of
 * `colour` and for the halo it must fill on the neighbour. */
int cmg_slab_boundary_ptr(cmg_context *ctx, int colour, int side, void **dev_ptr,
                          int64_t *n_bytes);
int cmg_slab_halo_ptr(cmg_context *ctx, int colour, int side, void **dev_ptr,
                      int64_t *n_bytes);
/* CUDA-IPC export/import of the halo buffers for the fused peer-store path:
 * once peers are attached, the half-sweep kernel itself writes its boundary
 * results into the neighbour's halo over NVLink and raises a flag there. */
int cmg_slab_ipc_export(cmg_context *ctx, void *handle_out, int64_t handle_bytes);
int cmg_slab_ipc_attach(cmg_context *ctx, int side, const void *handle,
                        int64_t handle_bytes, int same_process,
                        cmg_context *peer_if_same_process);
/* one half-sweep of one pass (checkerboard), for drivers that interleave their
 * own halo exchange.  pass_index is the global pass number (Philox counter). */
int cmg_slab_half_sweep(cmg_context *ctx, int colour, uint64_t pass_index,
                        int sample);
/* The pass loop of a slab with both neighbours attached (the body of
 * methods::basic_occupation_metropolis, include/casm/monte/methods/
 * basic_occupation_metropolis.hh:381-422, on this rank's columns): n_passes x
 * (colour-0 half-sweep, colour-1 half-sweep) with the halo push and the
 * neighbour flags fused into the kernels -- no host work, collective or second
 * call between half-sweeps.  Continues from the context's pass counter;
 * samples the slab's partial (S, B) whenever the pass count is a multiple of
 * sample_period (> 0).  Every rank of the ring must make the same calls.
 * Asynchronous; a neighbour that never arrives is reported by the next
 * synchronising call (bounded waits), not by a hung GPU. */
int cmg_slab_run_passes(cmg_context *ctx, int64_t n_passes, int64_t sample_period);
/* measurement aid: enabled = 0 makes the half-sweeps neither push their
 * boundary columns nor wait for the neighbours (halos go stale; the timing is
 * that of the sweep alone, for the halo share of a decomposed run) */
int cmg_slab_set_halo_exchange(cmg_context *ctx, int enabled);

/* ---- model and conditions --------------------------------------------------
 * Replaces IsingFormationEnergy(J, lattice_type) (model.hh:168-181) and
 * SemiGrandCanonicalConditions{temperature, exchange_potential[0]}
 * (basic_semigrand_canonical.hh:36-80).  The host builds, with the reference's
 * exact expression order, the table of dE for every (spin, neighbour-sum) case
 * (model.hh:312-314, :430-433; basic_semigrand_canonical.hh:185-191) and of
 * exp(-dE*beta) (methods/metropolis.hh:33); kernels only look them up.
 * chain = -1 applies to all chains. */
int cmg_set_model(cmg_context *ctx, double J, int lattice_type);
int cmg_set_conditions(cmg_context *ctx, int chain, double temperature, double mu);
/* tables as built: index = 2*n_up + b, n_up = number of +1 neighbours
 * (0..2*dim), b = 1 if the site holds +1.  Arrays of 2*(2*dim+1) entries. */
int cmg_get_tables(cmg_context *ctx, int chain, double *dE, double *prob,
                   uint32_t *thr_m1);

/* ---- occupation ------------------------------------------------------------
 * Replaces IsingConfiguration::set_occupation / occupation()
 * (model.hh:52-60).  n must equal n_sites.  Values must be +1 / -1. */
int cmg_upload_occupation_i32(cmg_context *ctx, int chain, const int32_t *occ,
                              int64_t n);
int cmg_download_occupation_i32(cmg_context *ctx, int chain, int32_t *occ,
                                int64_t n);
/* same, with DEVICE pointers (no host copy; for callers that already hold the
 * lattice in HBM, e.g. torch tensors) */
int cmg_upload_occupation_i32_dev(cmg_context *ctx, int chain,
                                  const int32_t *occ_dev, int64_t n);
int cmg_download_occupation_i32_dev(cmg_context *ctx, int chain, int32_t *occ_dev,
                                    int64_t n);
/* The same occupation in compact host formats, for callers that move many
 * lattices across PCIe: int8 +1/-1 per site, or one bit per site (bit l & 7 of
 * byte l >> 3 set iff site l holds +1; (n_sites + 7) / 8 bytes).  Site order l
 * as above.  The upload of bits is asynchronous (nothing to validate). */
int cmg_upload_occupation_i8(cmg_context *ctx, int chain, const int8_t *occ, int64_t n);
int cmg_download_occupation_i8(cmg_context *ctx, int chain, int8_t *occ, int64_t n);
int cmg_upload_occupation_bits(cmg_context *ctx, int chain, const uint8_t *bits,
                               int64_t n_sites);
int cmg_download_occupation_bits(cmg_context *ctx, int chain, uint8_t *bits,
                                 int64_t n_sites);
int cmg_fill_occupation(cmg_context *ctx, int chain, int value);
/* single-site access: IsingConfiguration::occ / set_occ (model.hh:62-70) */
int cmg_get_occ(cmg_context *ctx, int chain, int64_t linear_site_index, int32_t *value);
int cmg_set_occ(cmg_context *ctx, int chain, int64_t linear_site_index, int32_t value);
/* Change of formation energy and of N*x for an event on n_event sites, with the
 * reference's semantics for multi-site events (each flip evaluated after the
 * previous ones are applied, then all un-applied; model.hh:354-379, :425-435):
 *   dE_f = sum_i ((-J) * (new_i - old_i)) * (sum of the 2*dim neighbours)
 *   dNx  = sum_i (new_i - old_i) / 2.0
 * Evaluated on the device; the lattice is left unchanged. */
int cmg_event_delta(cmg_context *ctx, int chain, int n_event,
                    const int64_t *linear_site_index, const int32_t *new_occ,
                    double *dE_formation, double *dNx);
/* i.i.d. +1/-1 from Philox (synthetic benchmark input), probability of +1 = p_up */
int cmg_randomize_occupation(cmg_context *ctx, int chain, uint64_t seed, double p_up);

/* ---- random numbers --------------------------------------------------------
 * Checkerboard mode: counter-based Philox4x32-10; key = seed, counter = (site
 * group, chain, pass index, colour [, refinement]).  The acceptance uniform of a
 * site is the 32-bit integer R = rotl16(r16, 1) << 16 | r16' built from the
 * site's 16-bit lane of the leading and of the refinement call; the site flips
 * iff R <= ceil(p * 2^32) - 1, p = exp(-dE * beta) (DESIGN.md section 4 gives
 * the definition in full).  Serial mode: the reference's
 * RandomNumberGenerator<std::mt19937_64> (include/casm/monte/
 * RandomNumberGenerator.hh:15-42, definitions.hh:17) restated on the device:
 * libstdc++-13 uniform_int_distribution (Lemire) and generate_canonical. */
int cmg_seed_philox(cmg_context *ctx, uint64_t seed);
/* Rounds of the checkerboard mode's Philox4x32 stream: 10 (the published default, what
 * every number in BASELINE / bench.py is measured with) or 7, the fewest rounds that pass
 * BigCrush (Salmon et al., SC'11, table 2) -- an opt-in worth ~+17 % on the resident
 * kernel.  The trajectory is a different one; kernels ring2d,
 * bulk2d and generic.  No counterpart in the reference (its stream is mt19937_64). */
int cmg_set_philox_rounds(cmg_context *ctx, int rounds);
int cmg_set_pass_counter(cmg_context *ctx, uint64_t pass_index);
/* chains sharded over several contexts / GPUs keep the Philox stream of their
 * GLOBAL chain index: global index = this offset + local chain */
int cmg_set_chain_offset(cmg_context *ctx, int64_t global_index_of_chain_0);
int cmg_seed_mt19937_64(cmg_context *ctx, int chain, uint64_t seed);
/* raw engine state: 312 words + position, i.e. what operator<< of the engine
 * prints (python/src/monte.cpp:495-513 dump()/load()) */
int cmg_set_mt19937_64_state(cmg_context *ctx, int chain, const uint64_t *state312,
                             int position);
int cmg_get_mt19937_64_state(cmg_context *ctx, int chain, uint64_t *state312,
                             int *position);
/* draw through the device engine exactly as RandomNumberGenerator does
 * (random_int(max) in [0,max], random_real(max) in [0,max)) -- parity probe */
int cmg_rng_draw(cmg_context *ctx, int chain, int n, const int64_t *int_max,
                 const double *real_max, const uint8_t *is_real,
                 int64_t *int_out, double *real_out);

/* ---- the hot loop ----------------------------------------------------------
 * Replaces the body of methods::basic_occupation_metropolis
 * (include/casm/monte/methods/basic_occupation_metropolis.hh:381-422):
 * propose (basic_semigrand_canonical.hh:308-315), dE (:178-192), acceptance
 * (methods/metropolis.hh:26-35), apply (:318-320), pass counting (:399-403) and
 * sampling of the three default observables when n_pass hits the sample
 * schedule (:406-411).  Runs n_passes passes (n_sites attempts each) on every
 * chain; if sample_period > 0 a sample is appended to the on-device series
 * whenever the pass count is a multiple of sample_period.  Asynchronous. */
int cmg_run_passes(cmg_context *ctx, int64_t n_passes, int mode,
                   int64_t sample_period);
int cmg_counters(cmg_context *ctx, int chain, int64_t *n_pass, int64_t *n_accept,
                 int64_t *n_reject);
int cmg_reset_counters(cmg_context *ctx);

/* ---- sampling --------------------------------------------------------------
 * Integer sums of the current state: S = sum_l s_l and
 * B = sum_l s_l*(s_right + s_down [+ s_back]); the host (or cmg_read_samples)
 * applies the reference formulae (model.hh:266-270, :412-422;
 * basic_semigrand_canonical.hh:165-174) to get bit-identical doubles. */
int cmg_sample_now(cmg_context *ctx, int chain, int64_t *S, int64_t *B);
/* use_nlist=false form of the energy (model.hh:273-285): per-line integer dot
 * products, n0 "row" values then n1 "column" values (2-d only) */
int cmg_line_dots(cmg_context *ctx, int chain, int64_t *row_dots, int64_t *col_dots);
int cmg_n_samples(cmg_context *ctx, int64_t *n_samples);
int cmg_clear_samples(cmg_context *ctx);
int cmg_read_samples_sb(cmg_context *ctx, int chain, int64_t first, int64_t count,
                        int64_t *S, int64_t *B);
int cmg_read_samples(cmg_context *ctx, int chain, int quantity, int64_t first,
                     int64_t count, double *out);

/* ---- parity probes ---------------------------------------------------------
 * dE of flipping each site, in site order l (SemiGrandCanonicalPotential::
 * occ_delta_per_supercell for every single-site event); and the acceptance
 * decision of metropolis_acceptance for each site given one uniform per site.
 * Bit-exact against the reference expressions. */
int cmg_delta_e_probe(cmg_context *ctx, int chain, double *dE_per_site);
int cmg_accept_probe(cmg_context *ctx, int chain, const double *uniforms,
                     uint8_t *accept);

/* ---- statistics on the device-resident sample series -----------------------
 * Replaces BasicStatisticsCalculator::operator() (src/casm/monte/
 * BasicStatistics.cc:114-131: mean, variance, lag-k autocovariance search,
 * precision of the mean) and default_equilibration_check (src/casm/monte/
 * checks/EquilibrationCheck.cc:50-162) over samples [first, first+count).
 * k_star: lag found (0 = no-variation early-out, -1 = none found). */
int cmg_series_stats(cmg_context *ctx, int chain, int quantity, int64_t first,
                     int64_t count, double confidence, double *mean,
                     double *calculated_precision, double *variance,
                     int64_t *k_star);
int cmg_series_equilibration(cmg_context *ctx, int chain, int quantity,
                             int64_t count, double abs_precision,
                             int *is_equilibrated, int64_t *n_equil);
/* the same two, for every chain at once (results arrays of n_chains) */
int cmg_series_stats_all(cmg_context *ctx, int quantity, const int64_t *first,
                         int64_t count_total, double confidence, double *mean,
                         double *calculated_precision, double *variance,
                         int64_t *k_star);
int cmg_series_equilibration_all(cmg_context *ctx, int quantity, int64_t count,
                                 double abs_precision, int *is_equilibrated,
                                 int64_t *n_equil);
/* One completion check on the device-resident series of `chain` in a single
 * round trip: CompletionCheck::_check_convergence (include/casm/monte/checks/
 * CompletionCheck.hh:353-376) for n_components (<= 3) requested components with
 * absolute precisions: default_equilibration_check of samples [0, count) for each
 * (is_equilibrated / n_equil per component); if all equilibrated, n_stats =
 * count - max(n_equil) and BasicStatisticsCalculator over the last n_stats
 * samples of each (mean / calculated_precision); otherwise n_stats = 0 and the
 * statistics are not evaluated.  The callers' early exit at the first component
 * that has not equilibrated is applied by the caller to the returned arrays. */
int cmg_series_check(cmg_context *ctx, int chain, int n_components, const int *quantity,
                     const double *abs_precision, int64_t count, double confidence,
                     int *is_equilibrated, int64_t *n_equil, int64_t *n_stats, double *mean,
                     double *calculated_precision);
/* The same check ENQUEUED on the context's stream and not waited for.  A driver that knows
 * a check is due at the current sample count enqueues it, then its next (speculative) block
 * of passes behind a cmg_mark, and collects the verdict with cmg_series_check (identical
 * arguments) while that block runs: the device never idles between a block and the host's
 * decision (checks/CompletionCheck.hh:273-376 is evaluated exactly when and on what the
 * reference evaluates it; a "complete" verdict is followed by cmg_rollback). */
int cmg_series_check_prefetch(cmg_context *ctx, int chain, int n_components, const int *quantity,
                              const double *abs_precision, int64_t count, double confidence);
/* Restore point for drivers that sweep on while a completion check is evaluated
 * (methods::basic_occupation_metropolis decides after every sample, include/casm/monte/
 * methods/basic_occupation_metropolis.hh:381-422; a device loop that waited for every
 * decision would idle during the checks).  cmg_mark records "now" on the stream: a copy of
 * the state (planes, acceptance counters, pass and sample counters) follows the work
 * enqueued so far; a cmg_series_check of samples taken up to the mark then runs on a second
 * stream, concurrently with the passes enqueued after the mark.  cmg_rollback returns the
 * context to the mark (the samples taken since are dropped), so a "complete" verdict leaves
 * exactly the state the reference's loop would have stopped in.  Checkerboard contexts. */
int cmg_mark(cmg_context *ctx);
int cmg_rollback(cmg_context *ctx);

/* statistics of an arbitrary host series (Sampler columns that did not come
 * from the device path): uploads, computes on the device, returns */
int cmg_host_series_stats(int device, const double *x, int64_t n, double confidence,
                          double *mean, double *calculated_precision,
                          double *variance, int64_t *k_star);
int cmg_host_series_equilibration(int device, const double *x, int64_t n,
                                  double abs_precision, int *is_equilibrated,
                                  int64_t *n_equil);
/* weighted observations (time series of unequal intervals):
 * BasicStatisticsCalculator::operator()(observations, sample_weight)
 * (src/casm/monte/BasicStatistics.cc:144-188; method 1 = weighted mean and
 * weighted_variance (misc/math.hh:46-55) with the autocorrelation factor of the
 * resampled series at rho = 2^(-1/(k*increment)); method 2 = all statistics
 * from the resampled series), resample (BasicStatistics.cc:50-73) and the
 * weighted branch of default_equilibration_check (src/casm/monte/checks/
 * EquilibrationCheck.cc:137-161: x(i) * ((N/W) * w(i)) then the same scan).
 * W and the resampling walk follow the reference's sequential order. */
int cmg_host_series_stats_weighted(int device, const double *x, const double *w,
                                   int64_t n, double confidence, int method,
                                   int64_t n_resamples, double *mean,
                                   double *calculated_precision, double *variance,
                                   double *weight_sum, int64_t *k_star);
int cmg_host_series_resample(int device, const double *x, const double *w,
                             int64_t n, double weight_sum,
                             int64_t n_equally_spaced, double *out);
int cmg_host_series_equilibration_weighted(int device, const double *x,
                                           const double *w, int64_t n,
                                           double abs_precision,
                                           int *is_equilibrated, int64_t *n_equil);

/* ---- supercell index conversions --------------------------------------------
 * Replaces Conversions::l_to_b / l_to_ijk / bijk_to_l (include/casm/monte/
 * Conversions.hh:43-135, src/casm/monte/Conversions.cc:181-229) for a diagonal
 * transformation matrix diag(n0,n1,n2) and n_basis sublattices, batched on the
 * device: l = b*n_unitcells + i + n0*(j + n1*k), ijk wrapped periodically. */
int cmg_conv_l_to_bijk(int device, const int64_t *n3, int64_t n_basis,
                       const int64_t *l, int64_t count, int64_t *bijk_out);
int cmg_conv_bijk_to_l(int device, const int64_t *n3, int64_t n_basis,
                       const int64_t *bijk, int64_t count, int64_t *l_out);

/* The same two conversions for ANY integer transformation matrix T (row-major
 * 3 x 3, S = P * T), batched on the device.  xtal::UnitCellCoordIndexConverter
 * (CASMcode_crystallography, absent here) is restated in
 * include/casm_monte_b200/snf.hh: unit cells are numbered through the Smith
 * normal form of T; for diag(n0,n1,n2) with n0 | n1 | n2 this is the rule above,
 * for other T the order of unit cells is this library's own (the reference pins
 * no value).  The first-index-fastest forms above are the ones IsingConfiguration
 * uses (include/casm/monte/ising_cpp/model.hh:82-99). */
int cmg_conv_general_l_to_bijk(int device, const int64_t *T9, int64_t n_basis,
                               const int64_t *l, int64_t count, int64_t *bijk_out);
int cmg_conv_general_bijk_to_l(int device, const int64_t *T9, int64_t n_basis,
                               const int64_t *bijk, int64_t count, int64_t *l_out);

/* Energy form of the sampled formation / potential energies.  use_nlist != 0
 * (default): -J * (integer bond sum), IsingFormationEnergy::per_supercell with
 * the neighbour list (include/casm/monte/ising_cpp/model.hh:261-270).
 * use_nlist == 0: the row/column form of model.hh:273-285, sum over rows then
 * columns of (-J * dot), each term rounded and accumulated in double (it differs
 * in the last bits when J is not dyadic): the integer dots of every sampled pass
 * are taken on the device in checkerboard mode.  2-d, even extents, n0 % 32 == 0;
 * otherwise CMG_EUNSUPPORTED (callers then evaluate on the downloaded state). */
int cmg_set_energy_form(cmg_context *ctx, int use_nlist);

/* ---- k-state lattice model behind the general multi-species proposal tables ----
 * The step on the INPUT side of the path: sites hold one of
 * n_species <= 4 occupants (occupation index 0..K-1, one asymmetric-unit orbit,
 * occ_index == species_index); nearest-neighbour pair energy V[a][b] (symmetric,
 * row-major K x K) and one exchange potential per species:
 *     potential = sum_<ij> V[o_i][o_j] - sum_i mu[o_i]
 * (the Ising SGC potential, include/casm/monte/ising_cpp/basic_semigrand_canonical.hh:
 * 165-191, is K = 2, V = [[-J, J], [J, -J]], mu = (0, mu)).  Setting the model switches
 * the context to this path; seeds, counters, pass counter and streams are shared with
 * the Ising entry points.
 *   CMG_MODE_SERIAL_REFERENCE: the reference's general proposal machinery --
 *     OccCandidateList (src/casm/monte/events/OccCandidate.cc:32-60),
 *     make_semigrand_canonical_swaps (:136-157), OccLocation::initialize / choose_mol /
 *     apply (src/casm/monte/events/OccLocation.cc:39-116, :253-283),
 *     choose_semigrand_canonical_swap + propose_semigrand_canonical_event
 *     (include/casm/monte/events/OccEventProposal.hh:260-348) -- and
 *     metropolis_acceptance, on the mt19937_64 stream, n_sites steps per pass;
 *     the OccLocation lists persist between calls until the occupation is changed
 *     otherwise.
 *   CMG_MODE_CHECKERBOARD: coloured half-sweeps; every site proposes one of its
 *     K - 1 other species (word 0 of the site's Philox call) and accepts with the
 *     32-bit uniform of word 1 against the table threshold.
 * Tables hold, for entry (from * K + to) * n_cfg + cfg with cfg = sum_{s>=1} n_s *
 * (z + 1)^(s - 1) (n_s neighbours of species s, z = 2 * dim): dPhi, exp(-dPhi*beta),
 * the threshold and the never-accept flag.  Samples are integer sums: the number of
 * sites of every species and the histogram of bond types bonds[a * K + b], a <= b. */
int cmg_kstate_set_model(cmg_context *ctx, int n_species, const double *V);
int cmg_kstate_set_conditions(cmg_context *ctx, int chain, double temperature, const double *mu);
int cmg_kstate_get_tables(cmg_context *ctx, int chain, double *dPhi, double *prob, uint32_t *thr_m1,
                          uint8_t *never, int64_t n_entries);
int cmg_kstate_upload_occupation_i32(cmg_context *ctx, int chain, const int32_t *occ_index, int64_t n);
int cmg_kstate_download_occupation_i32(cmg_context *ctx, int chain, int32_t *occ_index, int64_t n);
int cmg_kstate_run_passes(cmg_context *ctx, int64_t n_passes, int mode, int64_t sample_period);
int cmg_kstate_n_samples(cmg_context *ctx, int64_t *n_samples);
int cmg_kstate_clear_samples(cmg_context *ctx);
int cmg_kstate_read_samples(cmg_context *ctx, int chain, int64_t first, int64_t count,
                            int64_t *counts, int64_t *bonds);

/* ---- N-fold way (rejection-free) driver ------------------------------------------
 * Replaces the loop of methods::nfold (include/casm/monte/methods/nfold.hh:80-147) for
 * the Ising SGC model: per step the total rate, the selection of (event,
 * time_increment), a sample if one is due by count -- taken before the event is applied
 * and weighted with the time increment (:104-123) -- and the event (:129-137).  The event
 * selector is a template parameter supplied from outside the reference tree; this
 * library's is the Bortz-Kalos-Lebowitz selector over the rate classes of the acceptance
 * table (rate = 1 if dE < 0 else exp(-dE*beta)): class by random_real(total_rate), member
 * by random_int(n - 1) in the class list, time_increment = -log(1 - random_real(1)) /
 * total_rate, all on the chain's mt19937_64 stream.  Every step is an accepted event.
 * Samples (S, B) are appended to the regular series (cmg_read_samples*); their weights
 * and the expected Metropolis acceptance rate total_rate / n_sites (NfoldData::
 * expected_acceptance_rate, :38-41) are read with cmg_nfold_read_weights and feed the
 * weighted statistics (cmg_host_series_stats_weighted). */
int cmg_nfold_run(cmg_context *ctx, int64_t n_steps, int64_t sample_period_steps);
int cmg_nfold_read_weights(cmg_context *ctx, int chain, int64_t first, int64_t count,
                           double *weight, double *expected_acceptance_rate);
int cmg_nfold_time(cmg_context *ctx, int chain, double *time, int64_t *n_steps);

/* ---- introspection for bench / tests ---------------------------------------- */
/* number of kernels this context has launched since creation */
int cmg_launch_count(const cmg_context *ctx, int64_t *n_launches);
/* name of the half-sweep kernel variant the context selected ("bulk2d", ...) */
const char *cmg_kernel_variant(const cmg_context *ctx);
/* force a variant, for tests and measurements: "auto", "generic", "bulk2d"
 * (HBM-streaming strips), "bulk3d", "tile2d" (shared-memory tiles with
 * temporal blocking), "ring2d" (the whole lattice resident in the shared memory
 * of the GPU, one cooperative launch of many passes; n0 = 256, 512 or a multiple
 * of 1024 up to 8192); options are appended as
 * ":js=56" (strip length), ":p=3" / ":nt=512" (tile passes / threads) and
 * ":rp=128" (passes per ring launch), ":ns=74" (balanced strips per lattice /
 * layer), ":pdl=0" (bulk2d / bulk3d: plain launches instead of programmatic
 * dependent ones) and ":chain=0" (dependent launches that wait for the whole
 * previous half-sweep instead of for their neighbour CTAs).  Every variant and
 * every option produces the same trajectory. */
int cmg_set_kernel_variant(cmg_context *ctx, const char *name);

#ifdef __cplusplus
}
#endif
#endif /* CASM_MONTE_GPU_H */
