// tpetra_shim.hpp -- the few Teuchos/Tpetra/Thyra types the reference's hot-path classes
// expose in their signatures (Trilinos is not available here).  Same names, same argument
// meaning, host storage.  With a real Trilinos, delete this header and bind the C ABI
// directly.
#pragma once
#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Teuchos {
enum ETransp { NO_TRANS = 0, TRANS = 1, CONJ_TRANS = 2 };
template <class T>
using RCP = std::shared_ptr<T>;
template <class T>
RCP<T> rcp(T *p) { return RCP<T>(p); }
template <class T>
RCP<T> rcp(const std::shared_ptr<T> &p) { return p; }
template <class T>
using Array = std::vector<T>;
}  // namespace Teuchos

namespace Tpetra {

// contiguous local index space of `n` entries, 1-based global ids like the reference's maps
// (src/mesh.cpp:608-627); `first_gid` = global id of local entry 0
template <class LO = int, class GO = int>
class Map {
public:
  Map(std::size_t n_local, std::size_t n_global, GO first_gid) : n_(n_local), ng_(n_global), first_(first_gid) {}
  std::size_t getNodeNumElements() const { return n_; }
  std::size_t getGlobalNumElements() const { return ng_; }
  GO getGlobalElement(LO k) const { return first_ + k; }
  bool isSameAs(const Map &o) const { return n_ == o.n_ && ng_ == o.ng_ && first_ == o.first_; }

private:
  std::size_t n_, ng_;
  GO first_;
};

template <class S = double, class LO = int, class GO = int>
class MultiVector {
public:
  MultiVector(const Teuchos::RCP<const Map<LO, GO>> &map, std::size_t nvec, bool zero = true)
      : map_(map), n_(map->getNodeNumElements()), nv_(nvec), data_(n_ * nvec, zero ? S(0) : S(0)) {}
  virtual ~MultiVector() = default;
  Teuchos::RCP<const Map<LO, GO>> getMap() const { return map_; }
  std::size_t getLocalLength() const { return n_; }
  std::size_t getNumVectors() const { return nv_; }
  std::size_t getStride() const { return n_; }
  S *getDataNonConst(std::size_t j = 0) { return data_.data() + j * n_; }
  const S *getData(std::size_t j = 0) const { return data_.data() + j * n_; }
  void putScalar(S v) { data_.assign(data_.size(), v); }
  void replaceLocalValue(LO k, S v) { data_[k] = v; }

protected:
  Teuchos::RCP<const Map<LO, GO>> map_;
  std::size_t n_, nv_;
  std::vector<S> data_;
};

template <class S = double, class LO = int, class GO = int>
class Vector : public MultiVector<S, LO, GO> {
public:
  explicit Vector(const Teuchos::RCP<const Map<LO, GO>> &map, bool zero = true)
      : MultiVector<S, LO, GO>(map, 1, zero) {}
  S &operator[](std::size_t k) { return this->data_[k]; }
  const S &operator[](std::size_t k) const { return this->data_[k]; }
  // serial (one rank) reductions; the partition-independent device versions are
  // nosh_dot / nosh_norm2 of the C ABI
  S dot(const Vector &o) const {
    S s = 0;
    for (std::size_t k = 0; k < this->n_; k++) s += this->data_[k] * o.data_[k];
    return s;
  }
  S norm1() const {
    S s = 0;
    for (S v : this->data_) s += std::fabs(v);
    return s;
  }
  S norm2() const { return std::sqrt(dot(*this)); }
  S normInf() const {
    S s = 0;
    for (S v : this->data_) s = std::fmax(s, std::fabs(v));
    return s;
  }
};

template <class S = double, class LO = int, class GO = int>
class Operator {
public:
  virtual ~Operator() = default;
  virtual void apply(const MultiVector<S, LO, GO> &X, MultiVector<S, LO, GO> &Y,
                     Teuchos::ETransp mode = Teuchos::NO_TRANS, S alpha = S(1), S beta = S(0)) const = 0;
  virtual Teuchos::RCP<const Map<LO, GO>> getDomainMap() const = 0;
  virtual Teuchos::RCP<const Map<LO, GO>> getRangeMap() const = 0;
};

}  // namespace Tpetra
