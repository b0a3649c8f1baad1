// snf.hh -- unit-cell index arithmetic of a supercell with a general integer
// transformation matrix T (S = P * T), shared by the host mirror of
// `Conversions` (monte.hh) and by the C-ABI library (batched device forms).
//
// The reference delegates this to xtal::UnitCellIndexConverter /
// xtal::UnitCellCoordIndexConverter of CASMcode_crystallography v2.2.0
// (casm/crystallography/LinearIndexConverter.hh; call sites
// src/casm/monte/Conversions.cc:120-127, :185-229), which is absent from this
// image.  Its published scheme is restated here: with the Smith normal form
// T = U * S * V (U, V unimodular, S = diag(s0, s1, s2), s0 | s1 | s2), the
// unit cell of linear index ix is U * (m, n, p) brought within the supercell,
//     m = ix % s0,  n = (ix / s0) % s1,  p = ix / (s0 * s1),
// and the inverse is (m, n, p) = (U^-1 * ijk) mod (s0, s1, s2).  The linear site
// index is l = b * n_unitcells + ix (pinned by python/tests/events/
// test_Conversions.py:65-72).  U is not unique, so for a T that is not already
// in Smith normal form the ORDER of unit cells may differ from libcasm-xtal's:
// parity unpinned (no test of the reference holds a value); for a diagonal T
// with s0 | s1 | s2 (e.g. n x n x n) U = I and the order is first-index-fastest.
#ifndef CASM_MONTE_B200_SNF_HH
#define CASM_MONTE_B200_SNF_HH

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <utility>

namespace casm_monte_b200 {

struct Mat3l {
  int64_t a[3][3];
  static Mat3l identity() {
    Mat3l m;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m.a[r][c] = r == c;
    return m;
  }
  static Mat3l from_row_major(const int64_t *v) {
    Mat3l m;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m.a[r][c] = v[3 * r + c];
    return m;
  }
};
inline Mat3l mul(Mat3l const &x, Mat3l const &y) {
  Mat3l z;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      z.a[r][c] = 0;
      for (int k = 0; k < 3; ++k) z.a[r][c] += x.a[r][k] * y.a[k][c];
    }
  return z;
}
inline int64_t det(Mat3l const &m) {
  return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]) -
         m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]) +
         m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}
inline Mat3l adjugate(Mat3l const &m) {
  Mat3l c;
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (k + 1) % 3, c2 = (k + 2) % 3;
      // cofactor of (r, k), transposed into (k, r)
      c.a[k][r] = m.a[r1][c1] * m.a[r2][c2] - m.a[r1][c2] * m.a[r2][c1];
    }
  return c;
}
inline int64_t floor_div(int64_t x, int64_t y) {
  int64_t q = x / y, r = x % y;
  return (r != 0 && ((r < 0) != (y < 0))) ? q - 1 : q;
}
inline int64_t floor_mod(int64_t x, int64_t y) { return x - y * floor_div(x, y); }

/// Smith normal form by row / column operations: S = Uinv * T * Vinv diagonal,
/// positive, s0 | s1 | s2.  A T already in that form is left untouched (U = I).
struct SmithNormalForm {
  Mat3l S, Uinv, U;
};
inline SmithNormalForm smith_normal_form(Mat3l const &T) {
  Mat3l S = T, L = Mat3l::identity();  // L accumulates the row operations: S = L * T * (column ops)
  auto swap_rows = [&](int i, int j) {
    for (int c = 0; c < 3; ++c) {
      std::swap(S.a[i][c], S.a[j][c]);
      std::swap(L.a[i][c], L.a[j][c]);
    }
  };
  auto add_row = [&](int dst, int src, int64_t f) {  // row dst += f * row src
    for (int c = 0; c < 3; ++c) {
      S.a[dst][c] += f * S.a[src][c];
      L.a[dst][c] += f * L.a[src][c];
    }
  };
  auto negate_row = [&](int i) {
    for (int c = 0; c < 3; ++c) {
      S.a[i][c] = -S.a[i][c];
      L.a[i][c] = -L.a[i][c];
    }
  };
  auto swap_cols = [&](int i, int j) {
    for (int r = 0; r < 3; ++r) std::swap(S.a[r][i], S.a[r][j]);
  };
  auto add_col = [&](int dst, int src, int64_t f) {
    for (int r = 0; r < 3; ++r) S.a[r][dst] += f * S.a[r][src];
  };
  for (int t = 0; t < 3; ++t) {
    for (;;) {
      // pivot: the non-zero entry of smallest magnitude in the trailing block
      int pr = -1, pc = -1;
      for (int r = t; r < 3; ++r)
        for (int c = t; c < 3; ++c)
          if (S.a[r][c] != 0 && (pr < 0 || std::llabs(S.a[r][c]) < std::llabs(S.a[pr][pc]))) {
            pr = r;
            pc = c;
          }
      if (pr < 0) throw std::runtime_error("Conversions: transformation matrix is singular");
      if (pr != t) swap_rows(pr, t);
      if (pc != t) swap_cols(pc, t);
      bool clean = true;
      for (int r = t + 1; r < 3; ++r)
        if (S.a[r][t] != 0) {
          add_row(r, t, -floor_div(S.a[r][t], S.a[t][t]));
          clean = clean && S.a[r][t] == 0;
        }
      for (int c = t + 1; c < 3; ++c)
        if (S.a[t][c] != 0) {
          add_col(c, t, -floor_div(S.a[t][c], S.a[t][t]));
          clean = clean && S.a[t][c] == 0;
        }
      if (!clean) continue;
      // every remaining entry must be a multiple of the pivot
      bool divides = true;
      for (int r = t + 1; r < 3 && divides; ++r)
        for (int c = t + 1; c < 3; ++c)
          if (S.a[r][c] % S.a[t][t] != 0) {
            add_row(t, r, 1);  // bring the offending row up and start over
            divides = false;
            break;
          }
      if (divides) break;
    }
    if (S.a[t][t] < 0) negate_row(t);
  }
  SmithNormalForm out;
  out.S = S;
  out.Uinv = L;
  const int64_t d = det(L);  // +-1
  Mat3l adj = adjugate(L);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) adj.a[r][c] *= d;  // inverse of a unimodular matrix
  out.U = adj;
  return out;
}

/// xtal::UnitCellCoordIndexConverter restated (see the header comment).
struct SiteIndexConverter {
  Mat3l T, adjT, U, Uinv;
  int64_t detT = 1, s[3] = {1, 1, 1}, n_unitcells = 1, n_basis = 1;
  SiteIndexConverter() : T(Mat3l::identity()), adjT(Mat3l::identity()), U(Mat3l::identity()), Uinv(Mat3l::identity()) {}
  SiteIndexConverter(Mat3l const &_T, int64_t _n_basis) : T(_T), n_basis(_n_basis) {
    detT = det(T);
    if (detT == 0) throw std::runtime_error("Conversions: transformation matrix is singular");
    if (n_basis < 1) throw std::runtime_error("Conversions: need n_basis >= 1");
    adjT = adjugate(T);
    SmithNormalForm f = smith_normal_form(T);
    U = f.U;
    Uinv = f.Uinv;
    for (int d = 0; d < 3; ++d) s[d] = f.S.a[d][d];
    n_unitcells = s[0] * s[1] * s[2];
  }
  int64_t total_sites() const { return n_basis * n_unitcells; }
  /// ijk -> the equivalent unit cell inside the supercell: ijk - T * floor(T^-1 ijk)
  void bring_within(int64_t ijk[3]) const {
    int64_t f[3];
    for (int r = 0; r < 3; ++r) {
      const int64_t num = adjT.a[r][0] * ijk[0] + adjT.a[r][1] * ijk[1] + adjT.a[r][2] * ijk[2];
      f[r] = floor_div(num, detT);
    }
    for (int r = 0; r < 3; ++r) ijk[r] -= T.a[r][0] * f[0] + T.a[r][1] * f[1] + T.a[r][2] * f[2];
  }
  void unitcell(int64_t ix, int64_t ijk[3]) const {
    const int64_t mnp[3] = {ix % s[0], (ix / s[0]) % s[1], ix / (s[0] * s[1])};
    for (int r = 0; r < 3; ++r) ijk[r] = U.a[r][0] * mnp[0] + U.a[r][1] * mnp[1] + U.a[r][2] * mnp[2];
    bring_within(ijk);
  }
  int64_t unitcell_index(const int64_t ijk[3]) const {
    int64_t mnp[3];
    for (int r = 0; r < 3; ++r)
      mnp[r] = floor_mod(Uinv.a[r][0] * ijk[0] + Uinv.a[r][1] * ijk[1] + Uinv.a[r][2] * ijk[2], s[r]);
    return mnp[0] + s[0] * (mnp[1] + s[1] * mnp[2]);
  }
  void bijk(int64_t l, int64_t out[4]) const {
    if (l < 0 || l >= total_sites()) throw std::runtime_error("Conversions: linear site index out of range");
    out[0] = l / n_unitcells;
    unitcell(l % n_unitcells, out + 1);
  }
  int64_t linear_site_index(const int64_t bijk[4]) const {
    if (bijk[0] < 0 || bijk[0] >= n_basis) throw std::runtime_error("Conversions: sublattice index out of range");
    return bijk[0] * n_unitcells + unitcell_index(bijk + 1);
  }
};

}  // namespace casm_monte_b200

#endif
