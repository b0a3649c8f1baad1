// kstate_oracle.hh -- CPU restatement of the general multi-species proposal
// machinery of libcasm-monte driving a k-state lattice model.
//
// TEST INFRASTRUCTURE ONLY (see monte_oracle.hh).
//
// Restated from the reference, each block citing its source:
//   OccCandidate / OccSwap / OccCandidateList, make_semigrand_canonical_swaps
//       include/casm/monte/events/OccCandidate.hh:20-175,
//       src/casm/monte/events/OccCandidate.cc:12-157
//   Mol / OccTransform / OccEvent        include/casm/monte/events/OccEvent.hh:20-73
//   OccLocation (initialize / apply / choose_mol / cand_size, no atom tracking)
//       src/casm/monte/events/OccLocation.cc:39-116, :216-283,
//       include/casm/monte/events/OccLocation.hh:255-275
//   choose_semigrand_canonical_swap, propose_semigrand_canonical_event
//       include/casm/monte/events/OccEventProposal.hh:260-348
// The reference ships no k-state model for this machinery to drive (its Ising
// generator bypasses it), so the MODEL is this repo's: K <= 4 species on a
// square / simple-cubic lattice with a symmetric nearest-neighbour pair energy
// V[a][b] and an exchange potential mu[s] per species,
//     potential = sum_<ij> V[o_i][o_j] - sum_i mu[o_i],
// which is the Ising SGC potential for K = 2, V = [[-J, J], [J, -J]], mu = (0, mu)
// (include/casm/monte/ising_cpp/basic_semigrand_canonical.hh:165-191).
// PARITY: the proposal machinery is pinned by no value-holding test of the
// reference (tests/unit/monte/OccLocation_test.cpp checks self-consistency only):
// "parity unpinned"; the restatement is line-by-line.
#ifndef CASM_MONTE_B200_KSTATE_ORACLE_HH
#define CASM_MONTE_B200_KSTATE_ORACLE_HH

#include "monte_oracle.hh"

namespace monte_oracle {
namespace kstate {

constexpr int kMaxSpecies = 4;

// ---- OccCandidate.hh:20-80 -------------------------------------------------
struct OccCandidate {
  Index asym, species_index;
  bool operator<(OccCandidate const &B) const {
    if (asym != B.asym) return asym < B.asym;
    return species_index < B.species_index;
  }
};
struct OccSwap {
  OccCandidate cand_a, cand_b;
};

/// Minimal Conversions: one sublattice whose sites all share one asymmetric-unit
/// orbit and allow species 0..K-1 in that order (occ_index == species_index)
struct SimpleConversions {
  Index n_sites, K;
  Index asym_size() const { return 1; }
  Index species_size() const { return K; }
  Index occ_size(Index) const { return K; }
  Index l_to_asym(Index) const { return 0; }
  Index species_index(Index, Index occ_index) const { return occ_index; }
  Index occ_index(Index, Index species_index) const { return species_index; }
  bool species_allowed(Index, Index species_index) const { return species_index < K; }
};

// ---- OccCandidate.cc:32-60 (all possible candidates) -------------------------
struct OccCandidateList {
  std::vector<OccCandidate> m_candidate;
  std::vector<std::vector<Index>> m_species_to_cand_index;
  Index m_end = 0;
  OccCandidateList() {}
  explicit OccCandidateList(SimpleConversions const &convert) {
    for (Index asym = 0; asym < convert.asym_size(); ++asym) {
      if (convert.occ_size(asym) < 2) continue;
      for (Index i = 0; i < convert.occ_size(asym); ++i)
        m_candidate.push_back(OccCandidate{asym, convert.species_index(asym, i)});
    }
    m_end = static_cast<Index>(m_candidate.size());
    m_species_to_cand_index.assign(convert.asym_size(), std::vector<Index>(convert.species_size(), m_end));
    Index index = 0;
    for (auto const &cand : m_candidate) m_species_to_cand_index[cand.asym][cand.species_index] = index++;
  }
  Index index(OccCandidate const &cand) const { return m_species_to_cand_index[cand.asym][cand.species_index]; }
  Index index(Index asym, Index species_index) const { return m_species_to_cand_index[asym][species_index]; }
  Index size() const { return m_end; }
};

// ---- OccCandidate.cc:118-157 --------------------------------------------------
inline bool allowed_semigrand_canonical_swap(SimpleConversions const &convert, OccCandidate a, OccCandidate b) {
  return a.asym == b.asym && a.species_index != b.species_index && convert.species_allowed(a.asym, b.species_index);
}
inline std::vector<OccSwap> make_semigrand_canonical_swaps(SimpleConversions const &convert,
                                                           OccCandidateList const &list) {
  std::vector<OccSwap> swaps;
  for (auto const &a : list.m_candidate)
    for (auto const &b : list.m_candidate)
      if (allowed_semigrand_canonical_swap(convert, a, b)) swaps.push_back(OccSwap{a, b});
  return swaps;
}

// ---- OccEvent.hh:20-73 ----------------------------------------------------------
struct Mol {
  Index id, l, asym, species_index, loc;
};
struct OccTransform {
  Index l, mol_id, asym, from_species, to_species;
};
struct KOccEvent {
  std::vector<Index> linear_site_index;
  std::vector<int> new_occ;
  std::vector<OccTransform> occ_transform;
};

// ---- OccLocation (no atom tracking) ------------------------------------------------
struct OccLocation {
  SimpleConversions convert;
  OccCandidateList candidate_list;
  std::vector<std::vector<Index>> m_loc;  // [cand][i] -> mol id
  std::vector<Mol> m_mol;
  std::vector<Index> m_l_to_mol;
  OccLocation(SimpleConversions const &c, OccCandidateList const &list)
      : convert(c), candidate_list(list), m_loc(list.size()) {}
  // OccLocation.cc:39-116
  void initialize(std::vector<int> const &occupation) {
    m_mol.clear();
    m_l_to_mol.clear();
    for (auto &v : m_loc) v.clear();
    Index Nmut = 0;
    for (Index l = 0; l < static_cast<Index>(occupation.size()); ++l)
      if (convert.occ_size(convert.l_to_asym(l)) > 1) Nmut++;
    m_mol.resize(Nmut);
    Index mol_id = 0;
    for (Index l = 0; l < static_cast<Index>(occupation.size()); ++l) {
      Index asym = convert.l_to_asym(l);
      if (convert.occ_size(asym) > 1) {
        Index species_index = convert.species_index(asym, occupation[l]);
        Index cand_index = candidate_list.index(asym, species_index);
        Mol &mol = m_mol[mol_id];
        mol.id = mol_id;
        mol.l = l;
        mol.asym = asym;
        mol.species_index = species_index;
        mol.loc = static_cast<Index>(m_loc[cand_index].size());
        m_loc[cand_index].push_back(mol_id);
        m_l_to_mol.push_back(mol_id);
        mol_id++;
      } else {
        m_l_to_mol.push_back(Nmut);
      }
    }
  }
  // OccLocation.cc:253-283
  void apply(KOccEvent const &e, std::vector<int> &occupation) {
    for (auto const &occ : e.occ_transform) {
      Mol &mol = m_mol[occ.mol_id];
      if (mol.species_index != occ.from_species)
        throw std::runtime_error("Error in OccLocation::apply: species mismatch");
      occupation[mol.l] = static_cast<int>(convert.occ_index(mol.asym, occ.to_species));
      Index cand_index = candidate_list.index(mol.asym, mol.species_index);
      Index back = m_loc[cand_index].back();
      m_loc[cand_index][mol.loc] = back;
      m_mol[back].loc = mol.loc;
      m_loc[cand_index].pop_back();
      mol.species_index = occ.to_species;
      cand_index = candidate_list.index(mol.asym, mol.species_index);
      mol.loc = static_cast<Index>(m_loc[cand_index].size());
      m_loc[cand_index].push_back(mol.id);
    }
  }
  // OccLocation.hh:255-262
  template <typename GeneratorType>
  Mol const &choose_mol(Index cand_index, GeneratorType &rng) const {
    return m_mol[m_loc[cand_index][rng.random_int(static_cast<Index>(m_loc[cand_index].size()) - 1)]];
  }
  Index cand_size(OccCandidate const &cand) const {
    return static_cast<Index>(m_loc[candidate_list.index(cand)].size());
  }
};

// ---- OccEventProposal.hh:260-348 -------------------------------------------------
template <typename GeneratorType>
OccSwap const &choose_semigrand_canonical_swap(OccLocation const &occ_location,
                                               std::vector<OccSwap> const &swaps, GeneratorType &rng) {
  Index tsize = static_cast<Index>(swaps.size());
  std::vector<double> tsum(tsize + 1);
  tsum[0] = 0.;
  for (Index i = 0; i < tsize; ++i) tsum[i + 1] = tsum[i] + ((double)occ_location.cand_size(swaps[i].cand_a));
  if (tsum.back() == 0.0) throw std::runtime_error("Error in choose_semigrand_canonical_swap: No events possible.");
  double rand = rng.random_real(tsum.back());
  for (Index i = 0; i < tsize; ++i)
    if (rand < tsum[i + 1]) return swaps[i];
  throw std::runtime_error("Error in choose_semigrand_canonical_swap");
}
template <typename GeneratorType>
KOccEvent &propose_semigrand_canonical_event(KOccEvent &e, OccLocation const &occ_location,
                                             std::vector<OccSwap> const &swaps, GeneratorType &rng) {
  OccSwap const &swap = choose_semigrand_canonical_swap(occ_location, swaps, rng);
  e.occ_transform.resize(1);
  e.linear_site_index.resize(1);
  e.new_occ.resize(1);
  OccTransform &transform = e.occ_transform[0];
  Mol const &mol = occ_location.choose_mol(occ_location.candidate_list.index(swap.cand_a), rng);
  transform.mol_id = mol.id;
  transform.l = mol.l;
  transform.asym = swap.cand_a.asym;
  transform.from_species = swap.cand_a.species_index;
  transform.to_species = swap.cand_b.species_index;
  e.linear_site_index[0] = transform.l;
  e.new_occ[0] = static_cast<int>(occ_location.convert.occ_index(transform.asym, transform.to_species));
  return e;
}

// ---- the k-state model (this repo's; see the header) ---------------------------------
struct KStateModel {
  int dim = 2, K = 3;
  double V[kMaxSpecies][kMaxSpecies] = {};
};
/// neighbour configuration index: sum_{s >= 1} n_s * (z + 1)^(s - 1), n_s = number of
/// neighbours holding species s (n_0 is implied), z = 2 * dim
inline int config_index(int const *n_of_species, int K, int z) {
  int idx = 0, w = 1;
  for (int s = 1; s < K; ++s) {
    idx += n_of_species[s] * w;
    w *= (z + 1);
  }
  return idx;
}
inline int n_configs(int K, int z) {
  int w = 1;
  for (int s = 1; s < K; ++s) w *= (z + 1);
  return w;
}
/// change of the potential when a site with the given neighbours goes from -> to:
/// dE = sum over species s of n_s * (V[to][s] - V[from][s]) accumulated in species
/// order, then dPhi = dE - (mu[to] - mu[from])
inline double delta_potential(KStateModel const &m, double const *mu, int from, int to, int const *n_of_species) {
  double dE = 0.0;
  for (int s = 0; s < m.K; ++s) dE += n_of_species[s] * (m.V[to][s] - m.V[from][s]);
  return dE - (mu[to] - mu[from]);
}
struct KStateTable {
  int K, z, n_cfg;
  double beta;
  // [from][to][cfg]; entries with impossible neighbour counts stay 0
  std::vector<double> dPhi, prob;
  std::vector<uint32_t> thr_m1;
  std::vector<uint8_t> never;
  size_t at(int from, int to, int cfg) const { return (static_cast<size_t>(from) * K + to) * n_cfg + cfg; }
};
inline KStateTable make_kstate_table(KStateModel const &m, double T, double const *mu) {
  KStateTable t;
  t.K = m.K;
  t.z = 2 * m.dim;
  t.n_cfg = n_configs(m.K, t.z);
  t.beta = 1.0 / (KB * T);
  const size_t n = static_cast<size_t>(m.K) * m.K * t.n_cfg;
  t.dPhi.assign(n, 0.0);
  t.prob.assign(n, 0.0);
  t.thr_m1.assign(n, 0u);
  t.never.assign(n, 0);
  int cnt[kMaxSpecies];
  for (int cfg = 0; cfg < t.n_cfg; ++cfg) {
    int rest = cfg, total = 0;
    for (int s = 1; s < m.K; ++s) {
      cnt[s] = rest % (t.z + 1);
      rest /= (t.z + 1);
      total += cnt[s];
    }
    if (total > t.z) continue;
    cnt[0] = t.z - total;
    for (int from = 0; from < m.K; ++from)
      for (int to = 0; to < m.K; ++to) {
        if (from == to) continue;
        const double d = delta_potential(m, mu, from, to, cnt);
        const double p = std::exp(-d * t.beta);
        const size_t i = t.at(from, to, cfg);
        t.dPhi[i] = d;
        t.prob[i] = p;
        if (d < 0.0 || p >= 1.0) {
          t.thr_m1[i] = 0xFFFFFFFFu;
        } else if (!(p > 0.0)) {
          t.thr_m1[i] = 0u;
          t.never[i] = 1;
        } else {
          double scaled = std::ceil(p * 4294967296.0);
          if (scaled < 1.0) scaled = 1.0;
          if (scaled > 4294967296.0) scaled = 4294967296.0;
          t.thr_m1[i] = static_cast<uint32_t>(static_cast<uint64_t>(scaled) - 1u);
        }
      }
  }
  return t;
}

struct Lattice {
  std::vector<int> shape;
  long n0, n1, n2;
  int dim;
  explicit Lattice(std::vector<int> const &s) : shape(s), n0(s[0]), n1(s[1]), n2(s.size() == 3 ? s[2] : 1), dim(static_cast<int>(s.size())) {}
  long n_sites() const { return n0 * n1 * n2; }
  /// neighbours in the order +i, +j, -i, -j [, +k, -k]
  void neighbours(long l, long *nb) const {
    const long i = l % n0, j = (l / n0) % n1, k = l / (n0 * n1);
    nb[0] = (i + 1) % n0 + n0 * (j + n1 * k);
    nb[1] = i + n0 * ((j + 1) % n1 + n1 * k);
    nb[2] = (i + n0 - 1) % n0 + n0 * (j + n1 * k);
    nb[3] = i + n0 * ((j + n1 - 1) % n1 + n1 * k);
    if (dim == 3) {
      nb[4] = i + n0 * (j + n1 * ((k + 1) % n2));
      nb[5] = i + n0 * (j + n1 * ((k + n2 - 1) % n2));
    }
  }
};

/// integer observables: count of every species, and the histogram of bond types
/// B[a][b], a <= b, over the bonds (+i, +j [, +k]) of every site
struct KStateSample {
  long long count[kMaxSpecies] = {0, 0, 0, 0};
  long long bonds[kMaxSpecies][kMaxSpecies] = {};
};
inline KStateSample kstate_observables(Lattice const &L, std::vector<int> const &occ, int K) {
  KStateSample s;
  long nb[6];
  for (long l = 0; l < L.n_sites(); ++l) {
    s.count[occ[l]]++;
    L.neighbours(l, nb);
    const int fwd[3] = {0, 1, 4};
    for (int d = 0; d < L.dim; ++d) {
      int a = occ[l], b = occ[nb[fwd[d]]];
      if (a > b) std::swap(a, b);
      s.bonds[a][b]++;
    }
  }
  return s;
}
/// potential per supercell from the integer sums: sum_{a<=b} V[a][b] * B[a][b]
/// (a outer, b inner) - sum_s mu[s] * count[s]
inline double kstate_potential(KStateModel const &m, double const *mu, KStateSample const &s) {
  double e = 0.0;
  for (int a = 0; a < m.K; ++a)
    for (int b = a; b < m.K; ++b) e += m.V[a][b] * static_cast<double>(s.bonds[a][b]);
  double x = 0.0;
  for (int a = 0; a < m.K; ++a) x += mu[a] * static_cast<double>(s.count[a]);
  return e - x;
}

struct KStateRunResult {
  std::vector<int> occupation;
  long long n_accept = 0, n_reject = 0;
  std::vector<KStateSample> samples;
};

/// The reference's loop (methods/basic_occupation_metropolis.hh:381-411) with the general
/// proposal machinery: propose_semigrand_canonical_event -> dPhi -> metropolis_acceptance
/// -> OccLocation::apply, n_sites steps per pass.
template <typename EngineType>
KStateRunResult kstate_serial_run(std::vector<int> const &shape, std::vector<int> occ, KStateModel const &m, double T,
                                  double const *mu, std::shared_ptr<EngineType> engine, long n_passes,
                                  long sample_period) {
  Lattice L(shape);
  SimpleConversions convert{L.n_sites(), m.K};
  OccCandidateList list(convert);
  std::vector<OccSwap> swaps = make_semigrand_canonical_swaps(convert, list);
  OccLocation loc(convert, list);
  loc.initialize(occ);
  RandomNumberGenerator<EngineType> rng(engine);
  const double beta = 1.0 / (KB * T);
  KStateRunResult r;
  KOccEvent e;
  long nb[6];
  int cnt[kMaxSpecies];
  for (long pass = 0; pass < n_passes; ++pass) {
    for (long step = 0; step < L.n_sites(); ++step) {
      propose_semigrand_canonical_event(e, loc, swaps, rng);
      const long l = e.linear_site_index[0];
      L.neighbours(l, nb);
      for (int s = 0; s < m.K; ++s) cnt[s] = 0;
      for (int d = 0; d < 2 * L.dim; ++d) cnt[occ[nb[d]]]++;
      const double dPhi = delta_potential(m, mu, occ[l], e.new_occ[0], cnt);
      if (metropolis_acceptance(dPhi, beta, rng)) {
        r.n_accept++;
        loc.apply(e, occ);
      } else {
        r.n_reject++;
      }
    }
    if (sample_period > 0 && ((pass + 1) % sample_period) == 0) r.samples.push_back(kstate_observables(L, occ, m.K));
  }
  r.occupation = occ;
  return r;
}

/// Checkerboard order for the k-state model, scalar statement of the device kernel.
/// The two sites p = 2g, 2g + 1 of column (j, k) of colour c (plane index p as in
/// checkerboard_pass; h = n0 / 2 plane indices per column) share one call: group
/// G = g + ceil(h / 2) * (j + n1 * k), pass t, chain ch,
///   w = Philox4x32-10(counter = {lo32(G), (hi32(G)&0xff) | ch<<8, lo32(t), (hi32(t)<<2) | c}, key = seed)
/// The even site uses (w[0], w[1]), the odd one (w[2], w[3]): the first word chooses the
/// proposed species, j = (w * (K-1)) >> 32, to = j + (j >= from); the second is the
/// acceptance uniform: the site changes iff u <= thr_m1[from][to][cfg] (and never when
/// exp(-dPhi*beta) == 0).
inline KStateRunResult kstate_checkerboard_run(std::vector<int> const &shape, std::vector<int> occ,
                                               KStateModel const &m, double T, double const *mu, uint64_t seed,
                                               uint32_t chain, uint64_t pass0, long n_passes, long sample_period) {
  Lattice L(shape);
  KStateTable tab = make_kstate_table(m, T, mu);
  const long h = L.n0 / 2;
  std::array<uint32_t, 2> key = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  KStateRunResult r;
  long nb[6];
  int cnt[kMaxSpecies];
  for (long p_ = 0; p_ < n_passes; ++p_) {
    const uint64_t t = pass0 + static_cast<uint64_t>(p_);
    for (int colour = 0; colour < 2; ++colour)
      for (long k = 0; k < L.n2; ++k)
        for (long j = 0; j < L.n1; ++j)
          for (long p = 0; p < h; ++p) {
            const long i = 2 * p + ((j + k + colour) & 1);
            const long l = i + L.n0 * (j + L.n1 * k);
            const uint64_t hh = static_cast<uint64_t>((h + 1) / 2);
            const uint64_t q = static_cast<uint64_t>(p / 2) + hh * (static_cast<uint64_t>(j) + static_cast<uint64_t>(L.n1) * static_cast<uint64_t>(k));
            std::array<uint32_t, 4> ctr = {static_cast<uint32_t>(q), (static_cast<uint32_t>(q >> 32) & 0xffu) | (chain << 8),
                                           static_cast<uint32_t>(t), (static_cast<uint32_t>(t >> 32) << 2) | static_cast<uint32_t>(colour)};
            const std::array<uint32_t, 4> w4 = Philox4x32::generate(ctr, key);
            const uint32_t w[2] = {w4[2 * (p & 1)], w4[2 * (p & 1) + 1]};
            const int from = occ[l];
            const int jj = static_cast<int>((static_cast<uint64_t>(w[0]) * static_cast<uint64_t>(m.K - 1)) >> 32);
            const int to = jj + (jj >= from ? 1 : 0);
            L.neighbours(l, nb);
            for (int s = 0; s < m.K; ++s) cnt[s] = 0;
            for (int d = 0; d < 2 * L.dim; ++d) cnt[occ[nb[d]]]++;
            const size_t idx = tab.at(from, to, config_index(cnt, m.K, tab.z));
            if (!tab.never[idx] && w[1] <= tab.thr_m1[idx]) {
              occ[l] = to;
              r.n_accept++;
            } else {
              r.n_reject++;
            }
          }
    if (sample_period > 0 && ((p_ + 1) % sample_period) == 0) r.samples.push_back(kstate_observables(L, occ, m.K));
  }
  r.occupation = occ;
  return r;
}

}  // namespace kstate
}  // namespace monte_oracle

#endif

// ===========================================================================
// N-fold way (rejection-free) driver, include/casm/monte/methods/nfold.hh:80-147,
// for the Ising SGC model.  The reference's loop is: total_rate (before selection);
// (event, time_increment) = event_selector.select_event(); sample if due, with
// sample weight = time_increment; apply the event; time += time_increment; count++.
// The event selector is a template parameter that lives outside the reference tree
// (libcasm-clexmonte supplies one), so the selector is this repo's, the textbook
// Bortz-Kalos-Lebowitz one over rate classes:
//   event l = flip of site l, rate r_l = 1 if dE_l < 0 else exp(-dE_l * beta): the
//   Metropolis acceptance probability, one of 2 * (2 dim + 1) values (class c =
//   2 * n_up + b, the index of the acceptance table);
//   total_rate = sum over classes in index order of n_c * rate_c;
//   class chosen by u1 = random_real(total_rate) against the running sum, member by
//   random_int(n_c - 1) in the class list (lists are kept like OccLocation's: filled
//   in site order, swap-remove, append);
//   time_increment = -log(1 - random_real(1.0)) / total_rate.
// PARITY: unpinned (the reference holds no test of nfold); restatement of the loop.
// ===========================================================================
namespace monte_oracle {
namespace nfold {

struct ClassLists {
  int n_class;
  std::vector<std::vector<long>> members;  // [class] -> sites
  std::vector<int> site_class;
  std::vector<long> site_pos;
  void init(int n_class_, long n_sites) {
    n_class = n_class_;
    members.assign(n_class, {});
    site_class.assign(n_sites, 0);
    site_pos.assign(n_sites, 0);
  }
  void insert(long l, int c) {
    site_class[l] = c;
    site_pos[l] = static_cast<long>(members[c].size());
    members[c].push_back(l);
  }
  void move(long l, int c_new) {
    const int c = site_class[l];
    if (c == c_new) return;
    const long back = members[c].back();
    members[c][site_pos[l]] = back;
    site_pos[back] = site_pos[l];
    members[c].pop_back();
    insert(l, c_new);
  }
};

struct NfoldResult {
  std::vector<int> occupation;
  std::vector<long long> S, B;     // integer observables at the sampled steps (before the event)
  std::vector<double> weight;      // time_increment of the event selected at the sample
  std::vector<double> expected_acceptance_rate;  // total_rate / n_events_possible at the sample
  std::vector<long> event_site;    // every selected site, in order (for trajectory parity)
  double time = 0.0;
  long long n_steps = 0;
};

template <typename EngineType>
NfoldResult nfold_run(std::vector<int> const &shape, std::vector<int> occ, double J, double T, double mu,
                      std::shared_ptr<EngineType> engine, long long n_steps, long long sample_period,
                      bool keep_events) {
  kstate::Lattice L(shape);
  const int dim = L.dim, z = 2 * dim, n_class = 2 * (z + 1);
  AcceptTable tab = make_accept_table(dim, J, T, mu);
  std::vector<double> rate(n_class);
  for (int nu = 0; nu <= z; ++nu)
    for (int b = 0; b < 2; ++b) rate[2 * nu + b] = tab.dE[b][nu] < 0.0 ? 1.0 : tab.prob[b][nu];
  auto n_up_of = [&](long l) {
    long nb[6];
    L.neighbours(l, nb);
    int n = 0;
    for (int d = 0; d < z; ++d) n += occ[nb[d]] > 0;
    return n;
  };
  ClassLists lists;
  lists.init(n_class, L.n_sites());
  for (long l = 0; l < L.n_sites(); ++l) lists.insert(l, 2 * n_up_of(l) + (occ[l] > 0));
  long long S = 0, B = 0;
  integer_observables(occ, shape, S, B);
  RandomNumberGenerator<EngineType> rng(engine);
  NfoldResult r;
  double time = 0.0;
  for (long long step = 0; step < n_steps; ++step) {
    double total_rate = 0.0;
    for (int c = 0; c < n_class; ++c) total_rate += static_cast<double>(lists.members[c].size()) * rate[c];
    // select_event
    const double u1 = rng.random_real(total_rate);
    int chosen = -1;
    double cum = 0.0;
    for (int c = 0; c < n_class; ++c) {
      cum += static_cast<double>(lists.members[c].size()) * rate[c];
      if (u1 < cum) {
        chosen = c;
        break;
      }
    }
    if (chosen < 0)  // rounding at the upper end: the last non-empty class
      for (int c = n_class - 1; c >= 0; --c)
        if (!lists.members[c].empty()) {
          chosen = c;
          break;
        }
    const long l = lists.members[chosen][rng.random_int(static_cast<long>(lists.members[chosen].size()) - 1)];
    const double time_increment = -std::log(1.0 - rng.random_real(1.0)) / total_rate;
    // sample, if due by count (nfold.hh:119-123: before the event is applied, weight = time_increment)
    if (sample_period > 0 && ((step + 1) % sample_period) == 0) {
      r.S.push_back(S);
      r.B.push_back(B);
      r.weight.push_back(time_increment);
      r.expected_acceptance_rate.push_back(total_rate / static_cast<double>(L.n_sites()));
    }
    // apply (nfold.hh:129-131)
    if (keep_events) r.event_site.push_back(l);
    const int n_up = n_up_of(l);
    const int s = occ[l];
    B += static_cast<long long>(-2 * s) * (2 * n_up - z);
    S += -2 * s;
    occ[l] = -s;
    lists.move(l, 2 * n_up + (occ[l] > 0));
    long nb[6];
    L.neighbours(l, nb);
    for (int d = 0; d < z; ++d) lists.move(nb[d], 2 * n_up_of(nb[d]) + (occ[nb[d]] > 0));
    time += time_increment;
  }
  r.occupation = occ;
  r.time = time;
  r.n_steps = n_steps;
  return r;
}

}  // namespace nfold
}  // namespace monte_oracle
