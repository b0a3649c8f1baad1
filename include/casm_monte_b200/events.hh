// casm_monte_b200/events.hh -- host mirror of the general multi-species proposal
// machinery of libcasm-monte: the input side of the path, which
// turns occupant tables into the events the update kernels serve.
//
// Same names, members, ordering rules and error behaviour as
//   include/casm/monte/events/OccCandidate.hh + src/casm/monte/events/OccCandidate.cc
//   include/casm/monte/events/OccEvent.hh
//   include/casm/monte/events/OccLocation.hh + src/casm/monte/events/OccLocation.cc
//     (occupant bookkeeping; the atom-trajectory tracking used by kinetic Monte Carlo is
//      out of scope: update_atoms / track_unique_atoms / save_atom_info must be false)
//   include/casm/monte/events/OccEventProposal.hh:100-348 (canonical and semi-grand
//     canonical swap choice and event proposal)
// over this repo's `Conversions` mirror (monte.hh).  The device counterpart -- the
// same proposal walked on the GPU on the reference's mt19937_64 stream, and the
// coloured k-state sweep -- is cmg_kstate_* in include/casm_monte_gpu.h.
#ifndef CASM_MONTE_B200_EVENTS_HH
#define CASM_MONTE_B200_EVENTS_HH

#include <map>
#include <set>
#include <tuple>

#include "monte.hh"

namespace casm_monte_b200 {

// ---- OccCandidate.hh:20-80 -------------------------------------------------------
struct OccCandidate {
  OccCandidate(Index _asym, Index _species_index) : asym(_asym), species_index(_species_index) {}
  Index asym;
  Index species_index;
  bool operator<(OccCandidate B) const {
    if (asym != B.asym) return asym < B.asym;
    return species_index < B.species_index;
  }
  bool operator==(OccCandidate B) const { return asym == B.asym && species_index == B.species_index; }
  bool operator!=(OccCandidate B) const { return !(*this == B); }
};

class OccSwap {
 public:
  OccSwap(OccCandidate const &_cand_a, OccCandidate const &_cand_b) : cand_a(_cand_a), cand_b(_cand_b) {}
  OccCandidate cand_a;
  OccCandidate cand_b;
  void reverse() { std::swap(cand_a, cand_b); }
  OccSwap &sort() {
    OccSwap B(*this);
    B.reverse();
    if (B < *this) *this = B;
    return *this;
  }
  OccSwap sorted() const {
    OccSwap res(*this);
    res.sort();
    return res;
  }
  bool operator<(OccSwap const &B) const {
    return std::make_tuple(cand_a, cand_b) < std::make_tuple(B.cand_a, B.cand_b);
  }
  bool operator==(OccSwap const &B) const { return cand_a == B.cand_a && cand_b == B.cand_b; }
};

// ---- OccCandidate.hh:135-178, OccCandidate.cc:12-60 -----------------------------------
class OccCandidateList {
 public:
  typedef std::vector<OccCandidate>::const_iterator const_iterator;
  OccCandidateList() {}
  /// custom list of OccCandidate
  OccCandidateList(std::vector<OccCandidate> candidates, Conversions const &convert)
      : m_candidate(std::move(candidates)) {
    _make_lookup(convert);
  }
  /// all possible OccCandidate: every allowed species of every orbit with > 1 occupant
  explicit OccCandidateList(Conversions const &convert) {
    for (Index asym = 0; asym < convert.asym_size(); ++asym) {
      if (convert.occ_size(asym) < 2) continue;
      for (Index i = 0; i < convert.occ_size(asym); ++i)
        m_candidate.push_back(OccCandidate(asym, convert.species_index(asym, i)));
    }
    _make_lookup(convert);
  }
  /// index into the candidate list, or size() if not allowed
  Index index(OccCandidate const &cand) const { return m_species_to_cand_index.at(cand.asym).at(cand.species_index); }
  Index index(Index asym, Index species_index) const { return m_species_to_cand_index.at(asym).at(species_index); }
  OccCandidate const &operator[](Index candidate_index) const { return m_candidate.at(candidate_index); }
  const_iterator begin() const { return m_candidate.begin(); }
  const_iterator end() const { return m_candidate.end(); }
  Index size() const { return m_end; }

 private:
  void _make_lookup(Conversions const &convert) {
    m_end = static_cast<Index>(m_candidate.size());
    m_species_to_cand_index.assign(convert.asym_size(), std::vector<Index>(convert.species_size(), m_end));
    Index index = 0;
    for (auto const &cand : m_candidate) m_species_to_cand_index.at(cand.asym).at(cand.species_index) = index++;
  }
  std::vector<std::vector<Index>> m_species_to_cand_index;
  std::vector<OccCandidate> m_candidate;
  Index m_end = 0;
};

// ---- OccCandidate.cc:62-182 ------------------------------------------------------------
inline bool is_valid(Conversions const &convert, OccCandidate const &cand) {
  return cand.asym >= 0 && cand.asym < convert.asym_size() && cand.species_index >= 0 &&
         cand.species_index < convert.species_size() && convert.species_allowed(cand.asym, cand.species_index);
}
inline bool is_valid(Conversions const &convert, OccCandidate const &cand_a, OccCandidate const &cand_b) {
  return is_valid(convert, cand_a) && is_valid(convert, cand_b);
}
inline bool is_valid(Conversions const &convert, OccSwap const &swap) {
  return is_valid(convert, swap.cand_a, swap.cand_b);
}
inline bool allowed_canonical_swap(Conversions const &convert, OccCandidate cand_a, OccCandidate cand_b) {
  return is_valid(convert, cand_a) && is_valid(convert, cand_b) && cand_a.species_index != cand_b.species_index &&
         convert.species_allowed(cand_a.asym, cand_b.species_index) &&
         convert.species_allowed(cand_b.asym, cand_a.species_index);
}
/// a->b only (no reverse swaps)
inline std::vector<OccSwap> make_canonical_swaps(Conversions const &convert, OccCandidateList const &occ_candidate_list) {
  std::vector<OccSwap> canonical_swaps;
  for (auto const &cand_a : occ_candidate_list)
    for (auto const &cand_b : occ_candidate_list)
      if (cand_a < cand_b && allowed_canonical_swap(convert, cand_a, cand_b))
        canonical_swaps.push_back(OccSwap(cand_a, cand_b));
  return canonical_swaps;
}
inline bool allowed_semigrand_canonical_swap(Conversions const &convert, OccCandidate cand_a, OccCandidate cand_b) {
  return is_valid(convert, cand_a) && is_valid(convert, cand_b) && cand_a.asym == cand_b.asym &&
         cand_a.species_index != cand_b.species_index && convert.species_allowed(cand_a.asym, cand_b.species_index);
}
/// a->b and b->a
inline std::vector<OccSwap> make_semigrand_canonical_swaps(Conversions const &convert,
                                                           OccCandidateList const &occ_candidate_list) {
  std::vector<OccSwap> swaps;
  for (auto const &cand_a : occ_candidate_list)
    for (auto const &cand_b : occ_candidate_list)
      if (allowed_semigrand_canonical_swap(convert, cand_a, cand_b)) swaps.push_back(OccSwap(cand_a, cand_b));
  return swaps;
}
inline Index get_n_allowed_per_unitcell(Conversions const &convert, std::vector<OccSwap> const &semigrand_canonical_swaps) {
  std::map<Index, Index> asym_to_n_swaps;
  for (Index asym = 0; asym < convert.asym_size(); ++asym) asym_to_n_swaps.emplace(asym, 0);
  for (OccSwap const &swap : semigrand_canonical_swaps) asym_to_n_swaps[swap.cand_a.asym]++;
  Index n_allowed_per_unitcell = 0;
  for (auto const &pair : asym_to_n_swaps)
    if (pair.second > 0)
      n_allowed_per_unitcell += (pair.second - 1) * static_cast<Index>(convert.asym_to_b(pair.first).size());
  return n_allowed_per_unitcell;
}

// ---- OccEvent.hh:20-73 (occupant bookkeeping part) ------------------------------------------
struct Mol {
  Index id = 0;             ///< Location in OccLocation.m_mol
  Index l = 0;              ///< Location in config
  Index asym = 0;           ///< Asym unit index (must be consistent with l)
  Index species_index = 0;  ///< Species type index (must be consistent with config.occ(l))
  Index loc = 0;            ///< Location in OccLocation.m_loc
};
// OccTransform and OccEvent (with occ_transform) are in monte.hh

// ---- OccLocation -----------------------------------------------------------------------------
class OccLocation {
 public:
  typedef Index size_type;
  OccLocation(Conversions const &_convert, OccCandidateList const &_candidate_list, bool _update_atoms = false,
              bool _track_unique_atoms = false, bool _save_atom_info = false)
      : m_convert(_convert), m_candidate_list(_candidate_list), m_loc(_candidate_list.size()) {
    if (_update_atoms || _track_unique_atoms || _save_atom_info)
      throw std::runtime_error(
          "Error constructing OccLocation: atom trajectory tracking (kinetic Monte Carlo) is out of scope");
  }
  /// OccLocation.cc:39-116
  void initialize(std::vector<int> const &occupation) {
    m_mol.clear();
    m_l_to_mol.clear();
    for (auto &vec : m_loc) vec.clear();
    if (static_cast<Index>(occupation.size()) != m_convert.l_size())
      throw std::runtime_error("Error in OccLocation::initialize: occupation.size() != l_size()");
    Index Nmut = 0;
    for (Index l = 0; l < static_cast<Index>(occupation.size()); ++l)
      if (m_convert.occ_size(m_convert.l_to_asym(l)) > 1) Nmut++;
    m_mol.resize(Nmut);
    m_l_to_mol.reserve(occupation.size());
    Index mol_id = 0;
    for (Index l = 0; l < static_cast<Index>(occupation.size()); ++l) {
      Index asym = m_convert.l_to_asym(l);
      if (m_convert.occ_size(asym) > 1) {
        Index species_index = m_convert.species_index(asym, occupation[l]);
        Index cand_index = m_candidate_list.index(asym, species_index);
        Mol &mol = m_mol[mol_id];
        mol.id = mol_id;
        mol.l = l;
        mol.asym = asym;
        mol.species_index = species_index;
        mol.loc = static_cast<Index>(m_loc.at(cand_index).size());
        m_loc[cand_index].push_back(mol_id);
        m_l_to_mol.push_back(mol_id);
        mol_id++;
      } else {
        m_l_to_mol.push_back(Nmut);
      }
    }
  }
  /// OccLocation.cc:253-283: update occupation and the lists to reflect that `e` occurred
  void apply(OccEvent const &e, std::vector<int> &occupation) {
    for (auto const &occ : e.occ_transform) {
      Mol &mol = m_mol.at(occ.mol_id);
      if (mol.species_index != occ.from_species)
        throw std::runtime_error("Error in OccLocation::apply: species mismatch");
      occupation.at(mol.l) = static_cast<int>(m_convert.occ_index(mol.asym, occ.to_species));
      Index cand_index = m_candidate_list.index(mol.asym, mol.species_index);
      Index back = m_loc[cand_index].back();
      m_loc[cand_index][mol.loc] = back;
      m_mol[back].loc = mol.loc;
      m_loc[cand_index].pop_back();
      mol.species_index = occ.to_species;
      cand_index = m_candidate_list.index(mol.asym, mol.species_index);
      mol.loc = static_cast<Index>(m_loc.at(cand_index).size());
      m_loc[cand_index].push_back(mol.id);
    }
  }
  /// OccLocation.hh:255-291
  template <typename GeneratorType>
  Mol const &choose_mol(Index cand_index, GeneratorType &random_number_generator) const {
    return mol(m_loc.at(cand_index)[random_number_generator.random_int(
        static_cast<Index>(m_loc.at(cand_index).size()) - 1)]);
  }
  template <typename GeneratorType>
  Mol const &choose_mol(Index cand_index, std::set<Index> exclude, GeneratorType &random_number_generator) const {
    Index loc;
    do {
      loc = random_number_generator.random_int(static_cast<Index>(m_loc.at(cand_index).size()) - 1);
    } while (exclude.count(loc));
    return mol(m_loc[cand_index][loc]);
  }
  template <typename GeneratorType>
  Mol const &choose_mol(OccCandidate const &cand, GeneratorType &random_number_generator) const {
    return choose_mol(m_candidate_list.index(cand), random_number_generator);
  }
  size_type mol_size() const { return static_cast<size_type>(m_mol.size()); }
  Mol &mol(Index mol_id) { return m_mol.at(mol_id); }
  Mol const &mol(Index mol_id) const { return m_mol.at(mol_id); }
  OccCandidateList const &candidate_list() const { return m_candidate_list; }
  size_type cand_size(Index cand_index) const { return static_cast<size_type>(m_loc.at(cand_index).size()); }
  size_type cand_size(OccCandidate const &cand) const { return cand_size(m_candidate_list.index(cand)); }
  Index mol_id(Index cand_index, Index loc) const { return m_loc.at(cand_index).at(loc); }
  Index mol_id(OccCandidate const &cand, Index loc) const { return mol_id(m_candidate_list.index(cand), loc); }
  Index l_to_mol_id(Index l) const { return m_l_to_mol.at(l); }
  Conversions const &convert() const { return m_convert; }

 private:
  Conversions const &m_convert;
  OccCandidateList const &m_candidate_list;
  std::vector<std::vector<Index>> m_loc;
  std::vector<Mol> m_mol;
  std::vector<Index> m_l_to_mol;
};

// ---- OccEventProposal.hh:108-348 ---------------------------------------------------------------
template <typename GeneratorType>
OccSwap const &choose_canonical_swap(OccLocation const &occ_location, std::vector<OccSwap> const &canonical_swap,
                                     GeneratorType &random_number_generator) {
  Index tsize = static_cast<Index>(canonical_swap.size());
  std::vector<double> m_tsum(tsize + 1);
  m_tsum[0] = 0.;
  for (Index i = 0; i < tsize; ++i)
    m_tsum[i + 1] = m_tsum[i] + ((double)occ_location.cand_size(canonical_swap[i].cand_a)) *
                                    ((double)occ_location.cand_size(canonical_swap[i].cand_b));
  if (m_tsum.back() == 0.0) throw std::runtime_error("Error in choose_canonical_swap: No events possible.");
  double rand = random_number_generator.random_real(m_tsum.back());
  for (Index i = 0; i < tsize; ++i)
    if (rand < m_tsum[i + 1]) return canonical_swap[i];
  throw std::runtime_error("Error in choose_canonical_swap");
}
template <typename GeneratorType>
OccEvent &propose_canonical_event_from_swap(OccEvent &e, OccLocation const &occ_location,
                                                         OccSwap const &swap, GeneratorType &random_number_generator) {
  e.occ_transform.resize(2);
  e.linear_site_index.resize(2);
  e.new_occ.resize(2);
  OccTransform &transform_a = e.occ_transform[0];
  Mol const &mol_a = occ_location.choose_mol(swap.cand_a, random_number_generator);
  transform_a.mol_id = mol_a.id;
  transform_a.l = mol_a.l;
  transform_a.asym = swap.cand_a.asym;
  transform_a.from_species = swap.cand_a.species_index;
  transform_a.to_species = swap.cand_b.species_index;
  OccTransform &transform_b = e.occ_transform[1];
  Mol const &mol_b = occ_location.choose_mol(swap.cand_b, random_number_generator);
  transform_b.mol_id = mol_b.id;
  transform_b.l = mol_b.l;
  transform_b.asym = swap.cand_b.asym;
  transform_b.from_species = swap.cand_b.species_index;
  transform_b.to_species = swap.cand_a.species_index;
  for (Index i = 0; i < 2; ++i) {
    OccTransform const &t = e.occ_transform[i];
    e.linear_site_index[i] = t.l;
    e.new_occ[i] = static_cast<int>(occ_location.convert().occ_index(t.asym, t.to_species));
  }
  return e;
}
template <typename GeneratorType>
OccEvent &propose_canonical_event(OccEvent &e, OccLocation const &occ_location,
                                               std::vector<OccSwap> const &canonical_swap,
                                               GeneratorType &random_number_generator) {
  auto const &swap = choose_canonical_swap(occ_location, canonical_swap, random_number_generator);
  return propose_canonical_event_from_swap(e, occ_location, swap, random_number_generator);
}
template <typename GeneratorType>
OccSwap const &choose_semigrand_canonical_swap(OccLocation const &occ_location,
                                               std::vector<OccSwap> const &semigrand_canonical_swap,
                                               GeneratorType &random_number_generator) {
  Index tsize = static_cast<Index>(semigrand_canonical_swap.size());
  std::vector<double> m_tsum(tsize + 1);
  m_tsum[0] = 0.;
  for (Index i = 0; i < tsize; ++i)
    m_tsum[i + 1] = m_tsum[i] + ((double)occ_location.cand_size(semigrand_canonical_swap[i].cand_a));
  if (m_tsum.back() == 0.0)
    throw std::runtime_error("Error in choose_semigrand_canonical_swap: No events possible.");
  double rand = random_number_generator.random_real(m_tsum.back());
  for (Index i = 0; i < tsize; ++i)
    if (rand < m_tsum[i + 1]) return semigrand_canonical_swap[i];
  throw std::runtime_error("Error in choose_semigrand_canonical_swap");
}
template <typename GeneratorType>
OccEvent &propose_semigrand_canonical_event_from_swap(OccEvent &e,
                                                                   OccLocation const &occ_location,
                                                                   OccSwap const &swap,
                                                                   GeneratorType &random_number_generator) {
  e.occ_transform.resize(1);
  e.linear_site_index.resize(1);
  e.new_occ.resize(1);
  OccTransform &transform = e.occ_transform[0];
  Mol const &mol = occ_location.choose_mol(swap.cand_a, random_number_generator);
  transform.mol_id = mol.id;
  transform.l = mol.l;
  transform.asym = swap.cand_a.asym;
  transform.from_species = swap.cand_a.species_index;
  transform.to_species = swap.cand_b.species_index;
  e.linear_site_index[0] = transform.l;
  e.new_occ[0] = static_cast<int>(occ_location.convert().occ_index(transform.asym, transform.to_species));
  return e;
}
template <typename GeneratorType>
OccEvent &propose_semigrand_canonical_event(OccEvent &e, OccLocation const &occ_location,
                                                         std::vector<OccSwap> const &semigrand_canonical_swap,
                                                         GeneratorType &random_number_generator) {
  auto const &swap = choose_semigrand_canonical_swap(occ_location, semigrand_canonical_swap, random_number_generator);
  return propose_semigrand_canonical_event_from_swap(e, occ_location, swap, random_number_generator);
}

}  // namespace casm_monte_b200

#endif
