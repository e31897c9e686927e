// oracle/ad.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// A minimal tape-based reverse-mode AD for the CPU oracle.  The reference's
// adjoint kernels (`*.grad`, plb/engine/mpm_simulator.py:335-345) are produced
// by taichi==0.7.26's source-transform autodiff, which is not in the tree.  The
// oracle therefore records the forward arithmetic of each reference kernel on a
// Wengert tape and sweeps it in reverse with Taichi's published per-op rules:
//   * min(a,b): adjoint goes to a iff a < b, else to b   (ties -> b)
//   * max(a,b): adjoint goes to a iff b < a, else to b   (ties -> b)
//   * abs(a):   sgn(a) * adj  (sgn(0) == 0)
//   * casts to int, comparisons and `select` predicates carry no gradient
//   * `if` bodies are differentiated under the forward predicate
// This makes "differentiate through the finite-difference normal stencil" and
// the min/max tie-breaking hold by construction rather than by hand derivation.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace ad {

template <class T>
struct Tape {
  struct Node {
    int a, b;
    T da, db;
  };
  std::vector<Node> nodes;
  std::vector<T> adj;
  int push(int a, T da, int b, T db) {
    nodes.push_back(Node{a, b, da, db});
    return (int)nodes.size() - 1;
  }
  void clear() {
    nodes.clear();
    adj.clear();
  }
  void begin_reverse() { adj.assign(nodes.size(), T(0)); }
  void reverse() {
    for (int i = (int)nodes.size() - 1; i >= 0; --i) {
      T g = adj[i];
      if (g == T(0)) continue;
      const Node& n = nodes[i];
      if (n.a >= 0) adj[n.a] += n.da * g;
      if (n.b >= 0) adj[n.b] += n.db * g;
    }
  }
};

template <class T>
inline Tape<T>& tape() {
  static thread_local Tape<T> t;
  return t;
}

template <class T>
struct Var {
  T v;
  int id;
  Var() : v(0), id(-1) {}
  Var(T c) : v(c), id(-1) {}
  Var(T c, int i) : v(c), id(i) {}
  static Var input(T c) { return Var(c, tape<T>().push(-1, T(0), -1, T(0))); }
  void seed(T g) const {
    if (id >= 0) tape<T>().adj[id] += g;
  }
  T grad() const { return id >= 0 ? tape<T>().adj[id] : T(0); }
};

template <class T>
inline Var<T> mk(T v, int a, T da, int b, T db) {
  if (a < 0 && b < 0) return Var<T>(v);
  return Var<T>(v, tape<T>().push(a, da, b, db));
}

#define AD_BIN(op, expr, dda, ddb)                                              \
  template <class T>                                                            \
  inline Var<T> operator op(const Var<T>& x, const Var<T>& y) {                 \
    return mk<T>(expr, x.id, dda, y.id, ddb);                                   \
  }                                                                             \
  template <class T>                                                            \
  inline Var<T> operator op(const Var<T>& x, T yc) {                            \
    Var<T> y(yc);                                                               \
    return mk<T>(expr, x.id, dda, -1, T(0));                                    \
  }                                                                             \
  template <class T>                                                            \
  inline Var<T> operator op(T xc, const Var<T>& y) {                            \
    Var<T> x(xc);                                                               \
    return mk<T>(expr, -1, T(0), y.id, ddb);                                    \
  }
AD_BIN(+, x.v + y.v, T(1), T(1))
AD_BIN(-, x.v - y.v, T(1), T(-1))
AD_BIN(*, x.v* y.v, y.v, x.v)
AD_BIN(/, x.v / y.v, T(1) / y.v, -x.v / (y.v * y.v))
#undef AD_BIN

template <class T>
inline Var<T> operator-(const Var<T>& x) {
  return mk<T>(-x.v, x.id, T(-1), -1, T(0));
}
template <class T>
inline Var<T>& operator+=(Var<T>& x, const Var<T>& y) {
  x = x + y;
  return x;
}
template <class T>
inline Var<T>& operator-=(Var<T>& x, const Var<T>& y) {
  x = x - y;
  return x;
}
template <class T>
inline Var<T>& operator+=(Var<T>& x, T y) {
  x = x + y;
  return x;
}
template <class T>
inline Var<T>& operator-=(Var<T>& x, T y) {
  x = x - y;
  return x;
}

// ---- scalar function layer shared by plain T and Var<T> ----
inline float val(float x) { return x; }
inline double val(double x) { return x; }
template <class T>
inline T val(const Var<T>& x) {
  return x.v;
}

inline float s_sqrt(float x) { return std::sqrt(x); }
inline double s_sqrt(double x) { return std::sqrt(x); }
template <class T>
inline Var<T> s_sqrt(const Var<T>& x) {
  T r = std::sqrt(x.v);
  return mk<T>(r, x.id, T(0.5) / r, -1, T(0));
}
// Emulation of fast-math transcendentals (test infrastructure).  The reference runs ti.init(fast_math=True)
// (plb/engine/taichi_env.py:20): on its CUDA backend log/exp are the hardware approximations (absolute error of log up to
// 2^-21.4 on [0.5, 2], exp about 2 ulp).  With amplitude a > 0 the oracle's log carries a deterministic pseudo-random
// absolute error in [-a, a] and exp a relative one in [-a/2, a/2] (hash of the argument, so a recompute sees the same
// value); the spread over a few amplitudes / salts is the formulation's sensitivity to that freedom.  0 = exact (default).
inline double& fast_math_noise() {
  static double a = 0.0;
  return a;
}
inline int& fast_math_salt() {
  static int s = 0;
  return s;
}
inline double fm_hash(double x) {   // in [-1, 1], deterministic in (x, salt)
  float xf = (float)x;
  unsigned h;
  std::memcpy(&h, &xf, 4);
  h ^= (unsigned)fast_math_salt() * 0x9E3779B1u;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return (double)(h & 0xFFFFFF) / (double)0x7FFFFF - 1.0;
}
template <class T>
inline T fm_exp(T x) {
  T r = std::exp(x);
  double a = fast_math_noise();
  return a > 0.0 ? (T)((double)r * (1.0 + 0.5 * a * fm_hash((double)x))) : r;
}
template <class T>
inline T fm_log(T x) {
  T r = std::log(x);
  double a = fast_math_noise();
  return a > 0.0 ? (T)((double)r + a * fm_hash((double)x)) : r;
}
inline float s_exp(float x) { return fm_exp(x); }
inline double s_exp(double x) { return fm_exp(x); }
template <class T>
inline Var<T> s_exp(const Var<T>& x) {
  T r = fm_exp(x.v);
  return mk<T>(r, x.id, r, -1, T(0));
}
inline float s_log(float x) { return fm_log(x); }
inline double s_log(double x) { return fm_log(x); }
template <class T>
inline Var<T> s_log(const Var<T>& x) {
  return mk<T>(fm_log(x.v), x.id, T(1) / x.v, -1, T(0));
}
inline float s_sin(float x) { return std::sin(x); }
inline double s_sin(double x) { return std::sin(x); }
template <class T>
inline Var<T> s_sin(const Var<T>& x) {
  return mk<T>(std::sin(x.v), x.id, std::cos(x.v), -1, T(0));
}
inline float s_cos(float x) { return std::cos(x); }
inline double s_cos(double x) { return std::cos(x); }
template <class T>
inline Var<T> s_cos(const Var<T>& x) {
  return mk<T>(std::cos(x.v), x.id, -std::sin(x.v), -1, T(0));
}
inline float s_abs(float x) { return std::fabs(x); }
inline double s_abs(double x) { return std::fabs(x); }
template <class T>
inline Var<T> s_abs(const Var<T>& x) {
  T s = x.v > T(0) ? T(1) : (x.v < T(0) ? T(-1) : T(0));
  return mk<T>(std::fabs(x.v), x.id, s, -1, T(0));
}
// Taichi min/max rule (ties -> second operand)
inline float s_min(float a, float b) { return a < b ? a : b; }
inline double s_min(double a, double b) { return a < b ? a : b; }
inline float s_max(float a, float b) { return b < a ? a : b; }
inline double s_max(double a, double b) { return b < a ? a : b; }
template <class T>
inline Var<T> s_min(const Var<T>& a, const Var<T>& b) {
  bool c = a.v < b.v;
  return mk<T>(c ? a.v : b.v, a.id, c ? T(1) : T(0), b.id, c ? T(0) : T(1));
}
template <class T>
inline Var<T> s_max(const Var<T>& a, const Var<T>& b) {
  bool c = b.v < a.v;
  return mk<T>(c ? a.v : b.v, a.id, c ? T(1) : T(0), b.id, c ? T(0) : T(1));
}
template <class T>
inline Var<T> s_min(const Var<T>& a, T b) {
  return s_min(a, Var<T>(b));
}
template <class T>
inline Var<T> s_min(T a, const Var<T>& b) {
  return s_min(Var<T>(a), b);
}
template <class T>
inline Var<T> s_max(const Var<T>& a, T b) {
  return s_max(a, Var<T>(b));
}
template <class T>
inline Var<T> s_max(T a, const Var<T>& b) {
  return s_max(Var<T>(a), b);
}

template <class S>
struct real_of {
  typedef S type;
};
template <class T>
struct real_of<Var<T>> {
  typedef T type;
};

}  // namespace ad
