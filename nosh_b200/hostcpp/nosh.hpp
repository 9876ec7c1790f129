// nosh.hpp -- C++ mirror of the reference's hot-path classes over the C ABI
// (include/nosh_b200.h).  Same class names, constructor arguments, method names and error
// behaviour as the reference, so that code written against
//   nosh::parameter_matrix::keo        (src/parameter_matrix_keo.hpp:37-60)
//   nosh::parameter_matrix::DkeoDP     (src/parameter_matrix_dkeo_dp.hpp)
//   nosh::jacobian_operator            (src/jacobian_operator.hpp:26-65)
//   nosh::keo_regularized              (src/keo_regularized.hpp:37-74)
//   nosh::model_evaluator::nls         (src/model_evaluator_nls.hpp:59-167)
//   nosh::scalar_field::constant / explicit_values, nosh::vector_field::explicit_values /
//   constantCurl                       (src/scalar_field_*.hpp, src/vector_field_*.hpp)
// compiles and runs with every operation executed by the sm_100a kernels.  No arithmetic
// happens in this header: each method forwards to one C-ABI call.
//
// Differences, all forced by the device boundary and documented in DESIGN.md section 2:
//  * a nosh::mesh owns ONE device context, hence one set of fields and one KEO buffer: the
//    field objects handed to the first operator built on a mesh are bound to it (the
//    reference shares one keo_ between the model evaluator and its Jacobian anyway,
//    src/model_evaluator_nls.cpp:258-263)
//  * meshes come from arrays or the synthetic generator (nosh::read needs MOAB)
//  * the Thyra InArgs/OutArgs protocol is reduced to plain structs with the same members
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <set>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nosh_b200.h"
#include "tpetra_shim.hpp"

namespace nosh {

// status -> the exception type the reference throws at the corresponding place
inline void check(nosh_ctx *ctx, nosh_status st) {
  if (st == NOSH_OK) return;
  const std::string msg = nosh_last_error(ctx);
  switch (st) {
    case NOSH_EKEY: throw std::out_of_range(msg);        // std::map::at
    case NOSH_EINVAL: throw std::logic_error(msg);       // TEUCHOS_TEST_FOR_EXCEPT_MSG
    default: throw std::runtime_error(msg);              // moab_wrap / CUDA
  }
}

struct param_list {
  std::vector<const char *> names;
  std::vector<double> values;
  explicit param_list(const std::map<std::string, double> &m) {
    for (const auto &kv : m) {
      names.push_back(kv.first.c_str());
      values.push_back(kv.second);
    }
  }
  int size() const { return (int)names.size(); }
};

// ---------------------------------------------------------------------------------------
// The communicator the reference's mesh carries (Teuchos::Comm<int>, src/mesh_reader.cpp:53-57; the tests run
// under mpiexec -n 2 / -n 7, test/CMakeLists.txt:18-23): rank, size and an all-gather of host records -- with MPI,
//   [](void *u, const void *s, void *r, int64_t n) { return MPI_Allgather(s, n, MPI_BYTE, r, n, MPI_BYTE, *(MPI_Comm *)u); }
// One rank (the default) needs no callback.  The device data path never goes through it (include/nosh_b200.h).
struct comm {
  int rank = 0, size = 1;
  nosh_allgather_fn allgather = nullptr;
  void *user = nullptr;
};

class mesh {
public:
  // general mesh from arrays (coords: n x 3, cells: nc x (dim+1), 0-based, GLOBAL numbering on every rank).
  // Several ranks: every rank keeps and uploads only its part (nosh_mesh_set_local, the READ_PART analogue).
  mesh(int dim, const std::vector<double> &coords, const std::vector<int> &cells, int device = 0,
       const comm &c = comm()) {
    create(device, c);
    set_from_global(dim, coords, cells);
    finish();
  }
  // synthetic structured tetrahedral grid
  mesh(int nx, int ny, int nz, double jitter = 0.2, uint64_t seed = 1234, int device = 0, const comm &c = comm()) {
    create(device, c);
    const double lo[3] = {-5, -5, -5}, hi[3] = {5, 5, 5};
    check(ctx_, nosh_mesh_tetgrid(ctx_, nx, ny, nz, lo, hi, jitter, seed));
    finish();
  }
  // nosh::read(file) (src/mesh_reader.cpp:19-162): mesh + vertex tags from a legacy VTK file (MOAB's
  // .h5m / Exodus need libraries that are not available: std::runtime_error with the conversion hint)
  explicit mesh(const std::string &file_name, int device = 0, const comm &c = comm()) {
    nosh_meshfile *f = nullptr;
    const nosh_status st = nosh_meshfile_read(file_name.c_str(), &f);
    if (st != NOSH_OK) throw std::runtime_error(std::string("nosh::read: ") + nosh_meshfile_last_error());
    int32_t dim = 0, nf = 0;
    int64_t nv = 0, nc = 0;
    nosh_meshfile_info(f, &dim, &nv, &nc, &nf);
    file_dim_ = dim;
    file_coords_.resize((size_t)nv * 3);
    file_cells_.resize((size_t)nc * (dim + 1));
    nosh_meshfile_get(f, file_coords_.data(), file_cells_.data());
    for (int32_t i = 0; i < nf; i++) {
      const char *name = nullptr;
      int32_t ncomp = 0;
      nosh_meshfile_field_name(f, i, &name, &ncomp);
      std::vector<double> v((size_t)nv * ncomp);
      nosh_meshfile_get_field(f, name, nullptr, v.data());
      tags_[name] = {ncomp, std::move(v)};
    }
    nosh_meshfile_free(f);
    create(device, c);
    set_from_global(dim, file_coords_, file_cells_);
    finish();
  }
  // vertex tags of the file (src/mesh.cpp:249-446) on the OWNED vertices of this rank (the file holds them in
  // global numbering; every rank reads the whole file -- legacy VTK has no partition information)
  std::shared_ptr<Tpetra::Vector<double, int, int>> get_vector(const std::string &tag) const {
    const auto &t = tag_at(tag, 1);
    auto v = std::make_shared<Tpetra::Vector<double, int, int>>(map_);
    const size_t b = (size_t)info_.owned_begin, n = (size_t)info_.n_owned;
    std::copy(t.begin() + b, t.begin() + b + n, v->getDataNonConst());
    return v;
  }
  std::shared_ptr<Tpetra::Vector<double, int, int>> get_complex_vector(const std::string &tag) const {
    const auto &t = tag_at(tag, 2);  // (re, im) per vertex = the interleaved layout of complex_map
    auto v = std::make_shared<Tpetra::Vector<double, int, int>>(complex_map_);
    const size_t b = 2 * (size_t)info_.owned_begin, n = 2 * (size_t)info_.n_owned;
    std::copy(t.begin() + b, t.begin() + b + n, v->getDataNonConst());
    return v;
  }
  std::shared_ptr<Tpetra::MultiVector<double, int, int>> get_multi_vector(const std::string &tag) const {
    const auto &t = tag_at(tag, 3);  // file: (x,y,z) per vertex; MultiVector: one column per component
    auto v = std::make_shared<Tpetra::MultiVector<double, int, int>>(map_, 3);
    const size_t b = (size_t)info_.owned_begin, n = (size_t)info_.n_owned;
    for (size_t k = 0; k < n; k++)
      for (int c = 0; c < 3; c++) v->getDataNonConst(c)[k] = t[3 * (b + k) + c];
    return v;
  }
  const std::vector<double> &tag_data(const std::string &tag) const { return tags_.at(tag).second; }
  // a vertex tag in LOCAL numbering (owned vertices first, then ghosts): what the field set-up calls of the C ABI take
  std::vector<double> tag_local(const std::string &tag, int ncomp) const {
    const auto &t = tag_at(tag, ncomp);
    const auto &g = local_gids();
    std::vector<double> out(g.size() * (size_t)ncomp);
    for (size_t k = 0; k < g.size(); k++)
      for (int c = 0; c < ncomp; c++) out[k * ncomp + c] = t[(size_t)g[k] * ncomp + c];
    return out;
  }
  // global (0-based) ids of the local vertices: owned first, then ghosts
  const std::vector<int64_t> &local_gids() const {
    if (gids_.empty() && info_.n_owned + info_.n_ghost > 0) {
      gids_.resize((size_t)(info_.n_owned + info_.n_ghost));
      check(ctx_, nosh_mesh_local_gids(ctx_, gids_.data()));
    }
    return gids_;
  }
  const comm &get_comm() const { return comm_; }

  // ---- subdomains (src/mesh.cpp:142-196, src/subdomain.hpp): vertex sets selected by is_inside(x), restricted
  // to the boundary skin for is_boundary_only subdomains; "everywhere" and "boundary" exist by default (:59-75)
  struct subdomain_def {
    std::string id;
    bool is_boundary_only;
    std::function<bool(const double *x)> is_inside;
  };
  void mark_subdomains(const std::vector<subdomain_def> &subdomains) const {
    ensure_vertex_data();
    for (const auto &sd : subdomains) {
      std::vector<int32_t> flags((size_t)info_.n_owned, 0);
      for (int64_t k = 0; k < info_.n_owned; k++)
        if ((!sd.is_boundary_only || boundary_[k]) && sd.is_inside(&coords_[3 * (size_t)k])) flags[k] = 1;
      subdomains_[sd.id] = std::move(flags);
    }
  }
  const std::vector<int32_t> &get_vertices(const std::string &subdomain_id) const {
    ensure_vertex_data();
    auto it = subdomains_.find(subdomain_id);
    if (it == subdomains_.end())  // src/mesh.cpp:199-210
      throw std::logic_error("Subdomain \"" + subdomain_id + "\" not found. Did you call mark_subdomains({...}) on the mesh?");
    return it->second;
  }
  const std::vector<double> &local_coords() const {
    ensure_vertex_data();
    return coords_;
  }
  // edge list in local vertex ids with length and covolume (src/mesh.hpp: get_edge_data / my_edges)
  struct edge_arrays {
    std::vector<int32_t> vertices;  // E x 2
    std::vector<double> length, covolume;
  };
  const edge_arrays &edge_data() const {
    if (edges_.length.empty() && info_.n_edges > 0) {
      edges_.vertices.resize(2 * (size_t)info_.n_edges);
      edges_.length.resize((size_t)info_.n_edges);
      edges_.covolume.resize((size_t)info_.n_edges);
      check(ctx_, nosh_mesh_get_edges(ctx_, edges_.vertices.data(), edges_.length.data(), edges_.covolume.data()));
    }
    return edges_;
  }
  // mesh::write (src/mesh.cpp:249-263), the outNNNN dumps of continuation_data_saver.hpp:24-50: the mesh of
  // the file with `psi` replaced by the given state
  void write(const std::string &file_name, const Tpetra::Vector<double, int, int> *psi = nullptr) const {
    if (file_coords_.empty()) throw std::logic_error("mesh::write: this mesh was not read from a file");
    std::vector<const char *> names;
    std::vector<int32_t> ncomps;
    std::vector<const double *> values;
    for (const auto &kv : tags_) {
      if (psi && kv.first == "psi") continue;
      names.push_back(kv.first.c_str());
      ncomps.push_back(kv.second.first);
      values.push_back(kv.second.second.data());
    }
    if (psi) {
      names.push_back("psi");
      ncomps.push_back(2);
      values.push_back(psi->getData());
    }
    if (nosh_meshfile_write(file_name.c_str(), file_dim_, (int64_t)file_coords_.size() / 3, file_coords_.data(),
                            (int64_t)file_cells_.size() / (file_dim_ + 1), file_cells_.data(), (int32_t)names.size(),
                            names.data(), ncomps.data(), values.data(), 1) != NOSH_OK)
      throw std::runtime_error(std::string("mesh::write: ") + nosh_meshfile_last_error());
  }
  ~mesh() { nosh_ctx_destroy(ctx_); }
  mesh(const mesh &) = delete;
  mesh &operator=(const mesh &) = delete;

  nosh_ctx *ctx() const { return ctx_; }
  std::shared_ptr<const Tpetra::Map<int, int>> map() const { return map_; }
  std::shared_ptr<const Tpetra::Map<int, int>> complex_map() const { return complex_map_; }
  // src/mesh.hpp:178: control volumes on the owned map
  std::shared_ptr<const Tpetra::Vector<double, int, int>> control_volumes() const {
    if (!cv_) {
      auto v = std::make_shared<Tpetra::Vector<double, int, int>>(map_);
      check(ctx_, nosh_mesh_get_control_volumes(ctx_, v->getDataNonConst()));
      cv_ = v;
    }
    return cv_;
  }
  const nosh_mesh_info_t &info() const { return info_; }

  // field binding (one field set per device context)
  mutable const void *bound_thickness = nullptr, *bound_mvp = nullptr, *bound_potential = nullptr;

private:
  void create(int device, const comm &c) {
    comm_ = c;
    if (nosh_ctx_create(device, nullptr, &ctx_) != NOSH_OK)
      throw std::runtime_error("nosh_ctx_create failed: no CUDA device (there is no CPU fallback)");
    if (c.size > 1) check(ctx_, nosh_ctx_comm_init_host(ctx_, c.rank, c.size, c.allgather, c.user));
  }
  // one rank: upload everything; several ranks: keep the cells touching my vertex range and their vertices
  template <class Int>
  void set_from_global(int dim, const std::vector<double> &coords, const std::vector<Int> &cells) {
    const int nvc = dim + 1;
    const int64_t nv = (int64_t)coords.size() / 3, nc = (int64_t)cells.size() / nvc;
    if (comm_.size == 1) {
      std::vector<int32_t> c32(cells.begin(), cells.end());
      check(ctx_, nosh_mesh_set(ctx_, dim, nv, coords.data(), nc, c32.data()));
      return;
    }
    int64_t b = 0, e = 0, g = 0, group = 65536;
    if (const char *env = std::getenv("NOSH_B200_GROUP")) group = std::atoll(env);
    if (nosh_partition_range(nv, comm_.size, comm_.rank, group, &b, &e, &g) != NOSH_OK)
      throw std::logic_error("nosh_partition_range failed");
    std::vector<int64_t> lid((size_t)nv, -1), gids;
    std::vector<int32_t> lc;
    for (int64_t c = 0; c < nc; c++) {
      bool mine = false;
      for (int k = 0; k < nvc; k++) mine = mine || ((int64_t)cells[c * nvc + k] >= b && (int64_t)cells[c * nvc + k] < e);
      if (!mine) continue;
      for (int k = 0; k < nvc; k++) {
        const int64_t v = cells[c * nvc + k];
        if (v < 0 || v >= nv) throw std::runtime_error("Illegal mesh: vertex index outside [0, n)");
        if (lid[v] < 0) {
          lid[v] = (int64_t)gids.size();
          gids.push_back(v);
        }
        lc.push_back((int32_t)lid[v]);
      }
    }
    std::vector<double> lx(gids.size() * 3);
    for (size_t k = 0; k < gids.size(); k++)
      for (int d = 0; d < 3; d++) lx[3 * k + d] = coords[3 * (size_t)gids[k] + d];
    check(ctx_, nosh_mesh_set_local(ctx_, dim, nv, (int64_t)gids.size(), gids.data(), lx.data(), (int64_t)lc.size() / nvc,
                                    lc.data()));
  }
  void finish() {
    check(ctx_, nosh_mesh_info(ctx_, &info_));
    map_ = std::make_shared<Tpetra::Map<int, int>>(info_.n_owned, info_.n_global, 1 + (int)info_.owned_begin);
    complex_map_ = std::make_shared<Tpetra::Map<int, int>>(2 * info_.n_owned, 2 * info_.n_global,
                                                           2 * (1 + (int)info_.owned_begin));
  }
  const std::vector<double> &tag_at(const std::string &tag, int ncomp) const {
    auto it = tags_.find(tag);
    if (it == tags_.end()) throw std::runtime_error("tag \"" + tag + "\" not found");  // moab_wrap.hpp:52-63
    if (it->second.first != ncomp) throw std::logic_error("tag \"" + tag + "\" has the wrong number of components");
    return it->second.second;
  }
  int file_dim_ = 0;
  std::vector<double> file_coords_;
  std::vector<int32_t> file_cells_;
  std::map<std::string, std::pair<int, std::vector<double>>> tags_;
  void ensure_vertex_data() const {
    if (!coords_.empty() || info_.n_owned + info_.n_ghost == 0) return;
    coords_.resize(3 * (size_t)(info_.n_owned + info_.n_ghost));
    check(ctx_, nosh_mesh_get_coords(ctx_, coords_.data()));
    boundary_.assign((size_t)info_.n_owned, 0);
    if (comm_.size == 1) check(ctx_, nosh_mesh_boundary_vertices(ctx_, boundary_.data()));
    subdomains_["everywhere"] = std::vector<int32_t>((size_t)info_.n_owned, 1);
    subdomains_["boundary"] = boundary_;
  }
  mutable std::vector<double> coords_;
  mutable std::vector<int32_t> boundary_;
  mutable std::map<std::string, std::vector<int32_t>> subdomains_;
  mutable edge_arrays edges_;
  nosh_ctx *ctx_ = nullptr;
  comm comm_;
  mutable std::vector<int64_t> gids_;
  nosh_mesh_info_t info_;
  std::shared_ptr<const Tpetra::Map<int, int>> map_, complex_map_;
  mutable std::shared_ptr<const Tpetra::Vector<double, int, int>> cv_;
};

// src/mesh_reader.hpp: std::shared_ptr<nosh::mesh> read(const std::string & file_name)
inline std::shared_ptr<nosh::mesh> read(const std::string &file_name, const comm &c = comm()) {
  return std::make_shared<nosh::mesh>(file_name, 0, c);
}

// ---------------------------------------------------------------------------------------
namespace scalar_field {
class base {
public:
  virtual ~base() = default;
  virtual const std::map<std::string, double> get_scalar_parameters() const = 0;
  virtual void bind_as_thickness(const mesh &m) const = 0;
  virtual void bind_as_potential(const mesh &m) const = 0;
};
// src/scalar_field_constant.hpp: constant(mesh, c, param1_name = "", param1_init_value = 0)
class constant : public base {
public:
  constant(const nosh::mesh &, double c, std::string param1_name = "", double param1_init_value = 0.0)
      : c_(c), name_(std::move(param1_name)), init_(param1_init_value) {}
  const std::map<std::string, double> get_scalar_parameters() const override {
    std::map<std::string, double> m;
    if (!name_.empty()) m[name_] = init_;
    return m;
  }
  void bind_as_thickness(const mesh &m) const override { check(m.ctx(), nosh_set_thickness(m.ctx(), nullptr, c_)); }
  void bind_as_potential(const mesh &m) const override {
    check(m.ctx(), nosh_set_potential_constant(m.ctx(), c_, name_.empty() ? nullptr : name_.c_str()));
  }

private:
  double c_;
  std::string name_;
  double init_;
};
// src/scalar_field_explicit_values.hpp: by tag name like the reference, or values given directly (local numbering)
class explicit_values : public base {
public:
  explicit_values(const nosh::mesh &, std::vector<double> values) : v_(std::move(values)) {}
  // the reference's constructor (src/scalar_field_explicit_values.hpp:20-23): the values are the vertex tag
  // `field_name` of the mesh file
  explicit_values(const nosh::mesh &m, const std::string &field_name) : v_(m.tag_local(field_name, 1)) {}
  const std::map<std::string, double> get_scalar_parameters() const override { return {{"beta", 1.0}}; }
  void bind_as_thickness(const mesh &m) const override { check(m.ctx(), nosh_set_thickness(m.ctx(), v_.data(), 0.0)); }
  void bind_as_potential(const mesh &m) const override { check(m.ctx(), nosh_set_potential_values(m.ctx(), v_.data())); }

private:
  std::vector<double> v_;
};
}  // namespace scalar_field

namespace vector_field {
class base {
public:
  virtual ~base() = default;
  virtual void set_parameters(const std::map<std::string, double> &params) = 0;
  virtual const std::map<std::string, double> get_scalar_parameters() const = 0;
  virtual void bind(const mesh &m) const = 0;
};
// src/vector_field_explicit_values.hpp: explicit_values(mesh, field_name, mu) reads the vertex tag (mesh tag "A"
// in the reference); the overload with a vector takes the nodal values directly, n_local x 3
class explicit_values : public base {
public:
  explicit_values(const nosh::mesh &, std::vector<double> A, double mu) : A_(std::move(A)), mu_(mu) {}
  // the reference's constructor (src/vector_field_explicit_values.hpp:17-21): the nodal values are the vertex tag
  // `field_name` of the mesh file
  explicit_values(const nosh::mesh &m, const std::string &field_name, double mu)
      : A_(m.tag_local(field_name, 3)), mu_(mu) {}
  void set_parameters(const std::map<std::string, double> &p) override { mu_ = p.at("mu"); }
  const std::map<std::string, double> get_scalar_parameters() const override { return {{"mu", mu_}}; }
  void bind(const mesh &m) const override { check(m.ctx(), nosh_set_mvp_explicit(m.ctx(), A_.data())); }

private:
  std::vector<double> A_;
  double mu_;
};
// src/vector_field_constant_curl.hpp: constantCurl(mesh, b, u)
class constantCurl : public base {
public:
  constantCurl(const std::shared_ptr<nosh::mesh> &m, const std::vector<double> &b, const std::vector<double> &u = {})
      : b_(b), u_(u) {
    // normalisation is checked at construction like the reference (constant_curl.cpp:35-44)
    check(m->ctx(), nosh_set_mvp_constcurl(m->ctx(), b_.data(), u_.empty() ? nullptr : u_.data()));
    m->bound_mvp = this;
  }
  void set_parameters(const std::map<std::string, double> &p) override {
    mu_ = p.at("mu");
    theta_ = p.at("theta");
  }
  const std::map<std::string, double> get_scalar_parameters() const override { return {{"mu", mu_}, {"theta", theta_}}; }
  void bind(const mesh &m) const override {
    check(m.ctx(), nosh_set_mvp_constcurl(m.ctx(), b_.data(), u_.empty() ? nullptr : u_.data()));
  }

private:
  std::vector<double> b_, u_;
  double mu_ = 0.0, theta_ = 0.0;
};
}  // namespace vector_field

inline void bind_fields(const mesh &m, const scalar_field::base *thickness, const vector_field::base *mvp,
                        const scalar_field::base *potential) {
  if (thickness && m.bound_thickness != thickness) {
    thickness->bind_as_thickness(m);
    m.bound_thickness = thickness;
  }
  if (mvp && m.bound_mvp != mvp) {
    mvp->bind(m);
    m.bound_mvp = mvp;
  }
  if (potential && m.bound_potential != potential) {
    potential->bind_as_potential(m);
    m.bound_potential = potential;
  }
}

// ---------------------------------------------------------------------------------------
// src/parameter_object.hpp:26-34
class parameter_object {
public:
  virtual ~parameter_object() = default;
  void set_parameters(const std::map<std::string, double> &scalar_params,
                      const std::map<std::string, std::shared_ptr<const Tpetra::Vector<double, int, int>>> &vector_params) {
    this->refill_(scalar_params, vector_params);  // the reference's cache never hits (parameter_object.cpp:17-44)
  }
  virtual std::map<std::string, double> get_scalar_parameters() const { return {}; }

protected:
  virtual void refill_(const std::map<std::string, double> &,
                       const std::map<std::string, std::shared_ptr<const Tpetra::Vector<double, int, int>>> &) = 0;
};

namespace parameter_matrix {

class matrix_base : public parameter_object, public Tpetra::Operator<double, int, int> {
public:
  matrix_base(const std::shared_ptr<const nosh::mesh> &mesh, const std::shared_ptr<const nosh::scalar_field::base> &thickness,
              const std::shared_ptr<nosh::vector_field::base> &mvp, nosh_matrix_id id)
      : mesh_(mesh), thickness_(thickness), mvp_(mvp), id_(id) {
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), nullptr);
  }
  std::map<std::string, double> get_scalar_parameters() const override { return mvp_->get_scalar_parameters(); }
  // Tpetra::CrsMatrix::apply: Y = alpha*op(A)*X + beta*Y
  void apply(const Tpetra::MultiVector<double, int, int> &X, Tpetra::MultiVector<double, int, int> &Y,
             Teuchos::ETransp mode = Teuchos::NO_TRANS, double alpha = 1.0, double beta = 0.0) const override {
    check(mesh_->ctx(), nosh_matrix_apply(mesh_->ctx(), id_, X.getData(), (int64_t)X.getStride(), Y.getDataNonConst(),
                                          (int64_t)Y.getStride(), (int)X.getNumVectors(), (nosh_transp)mode, alpha, beta));
  }
  Teuchos::RCP<const Tpetra::Map<int, int>> getDomainMap() const override { return mesh_->complex_map(); }
  Teuchos::RCP<const Tpetra::Map<int, int>> getRangeMap() const override { return mesh_->complex_map(); }
  const std::shared_ptr<const nosh::mesh> &get_mesh() const { return mesh_; }

  // ---- Tpetra::CrsMatrix row access (the reference's keo IS a CrsMatrix, src/parameter_matrix_keo.hpp:37).  The
  // device stores complex blocks; a local row of the real 2N x 2N matrix is unfolded from them: block K_ij = a + ib
  // contributes [a, -b] to row 2i and [b, a] to row 2i+1 at columns 2j, 2j+1 (src/parameter_matrix_keo.cpp:150-166).
  // Column indices are LOCAL (owned vertices first, then ghosts); values are fetched from the device once per fill.
  size_t getNodeNumRows() const { return 2 * (size_t)mesh_->info().n_owned; }
  size_t getNumEntriesInLocalRow(int row) const {
    fetch_blocks();
    const int64_t i = row / 2;
    return 2 * (size_t)(rowptr_[i + 1] - rowptr_[i]);
  }
  void getLocalRowCopy(int row, std::vector<int> &cols, std::vector<double> &vals, size_t &num) const {
    fetch_blocks();
    const int64_t i = row / 2;
    const bool imag_row = row & 1;
    num = 2 * (size_t)(rowptr_[i + 1] - rowptr_[i]);
    cols.resize(num);
    vals.resize(num);
    size_t k = 0;
    for (int64_t p = rowptr_[i]; p < rowptr_[i + 1]; p++, k += 2) {
      const double a = bvals_[2 * p], b = bvals_[2 * p + 1];
      cols[k] = 2 * bcols_[p];
      cols[k + 1] = 2 * bcols_[p] + 1;
      vals[k] = imag_row ? b : a;
      vals[k + 1] = imag_row ? a : -b;
    }
  }

protected:
  void invalidate_rows() { rowptr_.clear(); }
  void fetch_blocks() const {
    if (!rowptr_.empty()) return;
    const auto &mi = mesh_->info();
    rowptr_.resize((size_t)mi.n_owned + 1);
    bcols_.resize((size_t)mi.n_blocks);
    bvals_.resize(2 * (size_t)mi.n_blocks);
    check(mesh_->ctx(), nosh_get_block_csr(mesh_->ctx(), id_, rowptr_.data(), bcols_.data(), bvals_.data()));
  }
  const std::shared_ptr<const nosh::mesh> mesh_;
  const std::shared_ptr<const nosh::scalar_field::base> thickness_;
  const std::shared_ptr<nosh::vector_field::base> mvp_;
  nosh_matrix_id id_;
  mutable std::vector<int64_t> rowptr_;
  mutable std::vector<int32_t> bcols_;
  mutable std::vector<double> bvals_;
};

// src/parameter_matrix_keo.hpp:37-60
class keo : public matrix_base {
public:
  keo(const std::shared_ptr<const nosh::mesh> &mesh, const std::shared_ptr<const nosh::scalar_field::base> &thickness,
      const std::shared_ptr<nosh::vector_field::base> &mvp)
      : matrix_base(mesh, thickness, mvp, NOSH_MAT_KEO) {}

protected:
  void refill_(const std::map<std::string, double> &p,
               const std::map<std::string, std::shared_ptr<const Tpetra::Vector<double, int, int>>> &) override {
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), nullptr);
    mvp_->set_parameters(p);  // src/parameter_matrix_keo.cpp:88
    param_list pl(p);
    check(mesh_->ctx(), nosh_keo_fill(mesh_->ctx(), pl.size(), pl.names.data(), pl.values.data()));
    invalidate_rows();
  }
};

// src/parameter_matrix_dkeo_dp.hpp
class DkeoDP : public matrix_base {
public:
  DkeoDP(const std::shared_ptr<const nosh::mesh> &mesh, const std::shared_ptr<const nosh::scalar_field::base> &thickness,
         const std::shared_ptr<nosh::vector_field::base> &mvp, const std::string &param_name)
      : matrix_base(mesh, thickness, mvp, NOSH_MAT_DKEO), param_name_(param_name) {}

protected:
  void refill_(const std::map<std::string, double> &p,
               const std::map<std::string, std::shared_ptr<const Tpetra::Vector<double, int, int>>> &) override {
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), nullptr);
    mvp_->set_parameters(p);
    param_list pl(p);
    check(mesh_->ctx(), nosh_dkeo_fill(mesh_->ctx(), pl.size(), pl.names.data(), pl.values.data(), param_name_.c_str()));
    invalidate_rows();
  }
  const std::string param_name_;
};
}  // namespace parameter_matrix

// ---------------------------------------------------------------------------------------
// src/jacobian_operator.hpp:26-65
class jacobian_operator : public Tpetra::Operator<double, int, int> {
public:
  jacobian_operator(const std::shared_ptr<const nosh::mesh> &mesh,
                    const std::shared_ptr<const nosh::scalar_field::base> &scalar_potential,
                    const std::shared_ptr<const nosh::scalar_field::base> &thickness,
                    const std::shared_ptr<nosh::parameter_matrix::keo> &keo)
      : mesh_(mesh), scalar_potential_(scalar_potential), thickness_(thickness), keo_(keo) {
    bind_fields(*mesh_, thickness_.get(), nullptr, scalar_potential_.get());
  }
  nosh_ctx *ctx() const { return mesh_->ctx(); }
  void apply(const Tpetra::MultiVector<double, int, int> &X, Tpetra::MultiVector<double, int, int> &Y,
             Teuchos::ETransp mode = Teuchos::NO_TRANS, double alpha = 1.0, double beta = 0.0) const override {
    // unsupported mode/alpha/beta -> NOSH_EINVAL -> std::logic_error, as jacobian_operator.cpp:48-59
    check(mesh_->ctx(), nosh_jac_apply(mesh_->ctx(), X.getData(), (int64_t)X.getStride(), Y.getDataNonConst(),
                                       (int64_t)Y.getStride(), (int)X.getNumVectors(), (nosh_transp)mode, alpha, beta));
  }
  Teuchos::RCP<const Tpetra::Map<int, int>> getDomainMap() const override { return keo_->getDomainMap(); }
  Teuchos::RCP<const Tpetra::Map<int, int>> getRangeMap() const override { return keo_->getRangeMap(); }
  void rebuild(const std::map<std::string, double> &params, const Tpetra::Vector<double, int, int> &current_x) {
    bind_fields(*mesh_, thickness_.get(), nullptr, scalar_potential_.get());
    param_list pl(params);
    check(mesh_->ctx(), nosh_jac_rebuild(mesh_->ctx(), pl.size(), pl.names.data(), pl.values.data(), current_x.getData()));
  }

private:
  const std::shared_ptr<const nosh::mesh> mesh_;
  const std::shared_ptr<const nosh::scalar_field::base> scalar_potential_, thickness_;
  const std::shared_ptr<nosh::parameter_matrix::keo> keo_;
};

// src/keo_regularized.hpp:37-74
class keo_regularized : public Tpetra::Operator<double, int, int> {
public:
  keo_regularized(const std::shared_ptr<const nosh::mesh> &mesh, const std::shared_ptr<const nosh::scalar_field::base> &thickness,
                  const std::shared_ptr<nosh::vector_field::base> &mvp)
      : mesh_(mesh), thickness_(thickness), mvp_(mvp) {
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), nullptr);
  }
  nosh_ctx *ctx() const { return mesh_->ctx(); }
  // one AMG V-cycle on the regularised KEO (src/keo_regularized.cpp:88-165; MueLu in the reference, built on
  // the device here); unsupported mode/alpha/beta -> std::logic_error (:98-100)
  void apply(const Tpetra::MultiVector<double, int, int> &X, Tpetra::MultiVector<double, int, int> &Y,
             Teuchos::ETransp mode = Teuchos::NO_TRANS, double alpha = 1.0, double beta = 0.0) const override {
    check(mesh_->ctx(), nosh_keoreg_apply(mesh_->ctx(), X.getData(), (int64_t)X.getStride(), Y.getDataNonConst(),
                                          (int64_t)Y.getStride(), (int)X.getNumVectors(), (nosh_transp)mode, alpha, beta));
  }
  // the matrix K + diagonal blocks itself
  void apply_matrix(const Tpetra::MultiVector<double, int, int> &X, Tpetra::MultiVector<double, int, int> &Y) const {
    check(mesh_->ctx(), nosh_keoreg_matrix_apply(mesh_->ctx(), X.getData(), (int64_t)X.getStride(), Y.getDataNonConst(),
                                                 (int64_t)Y.getStride(), (int)X.getNumVectors()));
  }
  Teuchos::RCP<const Tpetra::Map<int, int>> getDomainMap() const override { return mesh_->complex_map(); }
  Teuchos::RCP<const Tpetra::Map<int, int>> getRangeMap() const override { return mesh_->complex_map(); }
  void rebuild(const std::map<std::string, double> &params, const Tpetra::Vector<double, int, int> &x) {
    mvp_->set_parameters(params);
    param_list pl(params);
    check(mesh_->ctx(), nosh_keoreg_rebuild(mesh_->ctx(), pl.size(), pl.names.data(), pl.values.data(), x.getData()));
  }

private:
  const std::shared_ptr<const nosh::mesh> mesh_;
  const std::shared_ptr<const nosh::scalar_field::base> thickness_;
  const std::shared_ptr<nosh::vector_field::base> mvp_;
};

// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// What nls::get_W_factory hands to NOX/LOCA: a Stratimikos "Belos" linear-solve strategy
// (src/model_evaluator_nls.cpp:269-299: "Pseudo Block CG", no preconditioner; the alternatives it keeps
// commented out -- "MINRES", "Pseudo Block GMRES" -- are what examples/conf.xml:103-136 configures).
// Stratimikos/Belos are not installed; this is the image of that parameter list over the device solvers.
// ---------------------------------------------------------------------------------------------
struct linear_solve_status {
  int iterations = 0;
  bool converged = false;
  double achieved_tol = 0.0;
};
class belos_solve_strategy {
public:
  std::string solver_type = "Pseudo Block CG";  // model_evaluator_nls.cpp:282
  std::string preconditioner_type = "None";     // :291
  double convergence_tolerance = 1e-8;          // Belos default; conf.xml sets 1e-10
  int maximum_iterations = 1000;                // Belos default
  int num_blocks = 300;                         // GMRES restart length, Belos default
  // solve op x = b; op: the jacobian_operator / keo / keo_regularized of `mesh`, prec: a rebuilt
  // keo_regularized (create_W_prec) or null
  linear_solve_status solve(const Tpetra::Operator<double, int, int> &op, const Tpetra::Vector<double, int, int> &b,
                            Tpetra::Vector<double, int, int> &x,
                            const std::shared_ptr<Tpetra::Operator<double, int, int>> &prec = nullptr) const {
    nosh_ctx *c = nullptr;
    nosh_operator_id id;
    if (auto j = dynamic_cast<const nosh::jacobian_operator *>(&op)) {
      c = j->ctx();
      id = NOSH_OP_JACOBIAN;
    } else if (auto k = dynamic_cast<const nosh::keo_regularized *>(&op)) {
      c = k->ctx();
      id = NOSH_OP_KEOREG;
    } else {
      throw std::logic_error("belos_solve_strategy: unknown operator type");
    }
    const bool use_prec = prec && preconditioner_type != "None";
    if (use_prec && !std::dynamic_pointer_cast<nosh::keo_regularized>(prec))
      throw std::logic_error("belos_solve_strategy: the preconditioner must be a nosh::keo_regularized");
    const nosh_precond pc = use_prec ? NOSH_PREC_KEOREG_AMG : NOSH_PREC_NONE;
    nosh_krylov_result r;
    if (solver_type == "Pseudo Block CG")
      check(c, nosh_cg_prec(c, id, pc, b.getData(), x.getDataNonConst(), convergence_tolerance, maximum_iterations, &r,
                            nullptr));
    else if (solver_type == "MINRES")
      check(c, nosh_minres_prec(c, id, pc, b.getData(), x.getDataNonConst(), convergence_tolerance, maximum_iterations,
                                &r, nullptr));
    else if (solver_type == "Pseudo Block GMRES")
      check(c, nosh_gmres(c, id, pc, b.getData(), x.getDataNonConst(), convergence_tolerance, maximum_iterations,
                          num_blocks, &r, nullptr));
    else
      throw std::logic_error("belos_solve_strategy: unknown \"Solver Type\" " + solver_type);
    linear_solve_status st;
    st.iterations = r.iterations;
    st.converged = r.converged != 0;
    st.achieved_tol = r.relres;
    return st;
  }
};

// ---------------------------------------------------------------------------------------------
// Generic finite-volume matrix (src/fvm_matrix.hpp, src/matrix_core_{edge,vertex,boundary,dirichlet}.hpp).  The
// reference's cores are virtual eval(moab::EntityHandle) objects generated by nfc; MOAB handles do not exist here,
// so eval receives what the generated bodies look up through the handle (edge index, end points, length,
// covolume / vertex index, coordinates, control volume).  fill() evaluates every core ONCE on the host into
// per-edge 2x2 blocks and per-vertex pairs and hands them to the device (nosh_fvm_matrix_fill: slot scatter, no
// atomics); apply() is the device SpMV.  One rank.
// ---------------------------------------------------------------------------------------------
struct edge_ref {
  int64_t index;
  int32_t v0, v1;            // local vertex ids, gid(v0) < gid(v1)
  const double *x0, *x1;     // coordinates
  double length, covolume;   // mesh->get_edge_data()[k]
};
struct vertex_ref {
  int64_t index;             // local (owned) vertex id
  const double *x;
  double control_volume;
};
struct matrix_core_edge_data {
  double lhs[2][2];
  double rhs[2];
};
class matrix_core_edge {
public:
  explicit matrix_core_edge(std::set<std::string> _subdomain_ids = {"everywhere"}) : subdomain_ids(std::move(_subdomain_ids)) {}
  virtual ~matrix_core_edge() = default;
  virtual matrix_core_edge_data eval(const edge_ref &edge) const = 0;
  const std::set<std::string> subdomain_ids;
};
struct vertex_data {
  double lhs, rhs;
};
class matrix_core_vertex {
public:
  explicit matrix_core_vertex(std::set<std::string> _subdomain_ids = {"everywhere"}) : subdomain_ids(std::move(_subdomain_ids)) {}
  virtual ~matrix_core_vertex() = default;
  virtual vertex_data eval(const vertex_ref &vertex) const = 0;
  const std::set<std::string> subdomain_ids;
};
using matrix_core_boundary = matrix_core_vertex;  // same shape (src/matrix_core_boundary.hpp)
class matrix_core_dirichlet {
public:
  explicit matrix_core_dirichlet(std::set<std::string> _subdomain_ids) : subdomain_ids(std::move(_subdomain_ids)) {}
  virtual ~matrix_core_dirichlet() = default;
  virtual double eval(const vertex_ref &vertex) const = 0;
  const std::set<std::string> subdomain_ids;
};

class fvm_matrix : public Tpetra::Operator<double, int, int> {
public:
  fvm_matrix(const std::shared_ptr<const nosh::mesh> &_mesh, std::vector<std::shared_ptr<const matrix_core_edge>> matrix_core_edges,
             std::vector<std::shared_ptr<const matrix_core_vertex>> matrix_core_vertexs,
             std::vector<std::shared_ptr<const matrix_core_boundary>> matrix_core_boundarys,
             std::vector<std::shared_ptr<const matrix_core_dirichlet>> dbcs)
      : mesh(_mesh), matrix_core_edges_(std::move(matrix_core_edges)), matrix_core_vertexs_(std::move(matrix_core_vertexs)),
        matrix_core_boundarys_(std::move(matrix_core_boundarys)), dbcs_(std::move(dbcs)) {}

  // src/fvm_matrix.hpp:44-70: zero, edge / vertex / boundary contributions, Dirichlet rows; rhs optional
  void fill(const std::shared_ptr<Tpetra::Vector<double, int, int>> &rhs = nullptr) {
    const auto &ed = mesh->edge_data();
    const auto &xc = mesh->local_coords();
    const int64_t E = mesh->info().n_edges, N = mesh->info().n_owned;
    auto cv = mesh->control_volumes();
    std::vector<double> elhs(4 * (size_t)E, 0.0), erhs(2 * (size_t)E, 0.0), vlhs((size_t)N, 0.0), vrhs((size_t)N, 0.0), dval((size_t)N, 0.0);
    std::vector<int32_t> dmask((size_t)N, 0);
    for (const auto &core : matrix_core_edges_)
      for (const auto &sd : core->subdomain_ids) {
        const auto &in = mesh->get_vertices(sd);
        for (int64_t k = 0; k < E; k++) {
          const int32_t v0 = ed.vertices[2 * k], v1 = ed.vertices[2 * k + 1];
          const bool in0 = v0 < N && in[v0], in1 = v1 < N && in[v1];
          if (!in0 && !in1) continue;
          const auto d = core->eval({k, v0, v1, &xc[3 * (size_t)v0], &xc[3 * (size_t)v1], ed.length[k], ed.covolume[k]});
          // interior edges of the subdomain contribute both rows, half edges the row of the vertex inside (:96-146)
          if (in0) {
            elhs[4 * k] += d.lhs[0][0];
            elhs[4 * k + 1] += d.lhs[0][1];
            erhs[2 * k] += d.rhs[0];
          }
          if (in1) {
            elhs[4 * k + 2] += d.lhs[1][0];
            elhs[4 * k + 3] += d.lhs[1][1];
            erhs[2 * k + 1] += d.rhs[1];
          }
        }
      }
    auto vertex_loop = [&](const std::vector<std::shared_ptr<const matrix_core_vertex>> &cores) {
      for (const auto &core : cores)
        for (const auto &sd : core->subdomain_ids) {
          const auto &in = mesh->get_vertices(sd);
          for (int64_t k = 0; k < N; k++)
            if (in[k]) {
              const auto d = core->eval({k, &xc[3 * (size_t)k], (*cv)[k]});
              vlhs[k] += d.lhs;
              vrhs[k] += d.rhs;
            }
        }
    };
    vertex_loop(matrix_core_vertexs_);
    vertex_loop(matrix_core_boundarys_);
    for (const auto &bc : dbcs_)
      for (const auto &sd : bc->subdomain_ids) {
        const auto &in = mesh->get_vertices(sd);
        for (int64_t k = 0; k < N; k++)
          if (in[k]) {
            dmask[k] = 1;
            dval[k] = bc->eval({k, &xc[3 * (size_t)k], (*cv)[k]});
          }
      }
    const bool any_d = !dbcs_.empty();
    check(mesh->ctx(), nosh_fvm_matrix_fill(mesh->ctx(), nullptr, elhs.data(), erhs.data(), vlhs.data(), vrhs.data(),
                                            any_d ? dmask.data() : nullptr, any_d ? dval.data() : nullptr,
                                            rhs ? rhs->getDataNonConst() : nullptr));
  }
  void apply(const Tpetra::MultiVector<double, int, int> &X, Tpetra::MultiVector<double, int, int> &Y,
             Teuchos::ETransp mode = Teuchos::NO_TRANS, double alpha = 1.0, double beta = 0.0) const override {
    if (mode != Teuchos::NO_TRANS || alpha != 1.0 || beta != 0.0) throw std::logic_error("fvm_matrix::apply: only y = A x");
    for (size_t j = 0; j < X.getNumVectors(); j++)
      check(mesh->ctx(), nosh_fvm_matrix_apply(mesh->ctx(), X.getData(j), Y.getDataNonConst(j)));
  }
  Teuchos::RCP<const Tpetra::Map<int, int>> getDomainMap() const override { return mesh->map(); }
  Teuchos::RCP<const Tpetra::Map<int, int>> getRangeMap() const override { return mesh->map(); }
  // mikado::linear_solve(A, b, x, {"package": "Belos", "method": "Pseudo Block CG"}) of examples/poisson/poisson.cpp
  linear_solve_status solve_cg(const Tpetra::Vector<double, int, int> &b, Tpetra::Vector<double, int, int> &x, double tol = 1e-10,
                               int maxit = 1000) const {
    nosh_krylov_result r;
    check(mesh->ctx(), nosh_fvm_cg(mesh->ctx(), b.getData(), x.getDataNonConst(), tol, maxit, &r));
    linear_solve_status st;
    st.iterations = r.iterations;
    st.converged = r.converged != 0;
    st.achieved_tol = r.relres;
    return st;
  }
  const std::shared_ptr<const nosh::mesh> mesh;

private:
  const std::vector<std::shared_ptr<const matrix_core_edge>> matrix_core_edges_;
  const std::vector<std::shared_ptr<const matrix_core_vertex>> matrix_core_vertexs_;
  const std::vector<std::shared_ptr<const matrix_core_boundary>> matrix_core_boundarys_;
  const std::vector<std::shared_ptr<const matrix_core_dirichlet>> dbcs_;
};

namespace model_evaluator {

// the members of Thyra::ModelEvaluatorBase::InArgs / OutArgs that nls supports
// (src/model_evaluator_nls.cpp:331-390)
struct InArgs {
  std::shared_ptr<const Tpetra::Vector<double, int, int>> x;
  std::vector<double> p;  // p(0), ordered like get_p_names(0)
  double alpha = 0.0, beta = 0.0;
  void set_x(const std::shared_ptr<const Tpetra::Vector<double, int, int>> &v) { x = v; }
  void set_p(int l, const std::vector<double> &v) {
    if (l != 0) throw std::logic_error("LOCA can only deal with one parameter vector.");
    p = v;
  }
};
struct OutArgs {
  std::shared_ptr<Tpetra::Vector<double, int, int>> f;
  std::shared_ptr<Tpetra::MultiVector<double, int, int>> DfDp;  // DERIV_MV_BY_COL
  std::shared_ptr<Tpetra::Operator<double, int, int>> W_op, W_prec;
  void set_f(const std::shared_ptr<Tpetra::Vector<double, int, int>> &v) { f = v; }
  void set_DfDp(int, const std::shared_ptr<Tpetra::MultiVector<double, int, int>> &v) { DfDp = v; }
  void set_W_op(const std::shared_ptr<Tpetra::Operator<double, int, int>> &v) { W_op = v; }
  void set_W_prec(const std::shared_ptr<Tpetra::Operator<double, int, int>> &v) { W_prec = v; }
};

// src/model_evaluator_nls.hpp:59-167
class nls {
public:
  nls(const std::shared_ptr<const nosh::mesh> &mesh, const std::shared_ptr<nosh::vector_field::base> &mvp,
      const std::shared_ptr<const nosh::scalar_field::base> &scalar_potential, const double g,
      const std::shared_ptr<const nosh::scalar_field::base> &thickness,
      const std::shared_ptr<const Tpetra::Vector<double, int, int>> &initial_x, const std::string &deriv_parameter)
      : mesh_(mesh), mvp_(mvp), scalar_potential_(scalar_potential), thickness_(thickness),
        keo_(std::make_shared<nosh::parameter_matrix::keo>(mesh_, thickness_, mvp_)),
        dkeo_dp_(std::make_shared<nosh::parameter_matrix::DkeoDP>(mesh_, thickness_, mvp_, deriv_parameter)),
        initial_x_(initial_x) {
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), scalar_potential_.get());
    // merge all parameters; earlier keys win; std::map keeps them name-sorted (:101-130)
    std::map<std::string, double> params;
    params["g"] = g;
    auto sp = scalar_potential_->get_scalar_parameters();
    params.insert(sp.begin(), sp.end());
    auto mb = keo_->get_scalar_parameters();
    params.insert(mb.begin(), mb.end());
    for (const auto &kv : params) {
      p_names_.push_back(kv.first);
      p_init_.push_back(kv.second);
    }
  }
  Teuchos::RCP<const Teuchos::Array<std::string>> get_p_names(int l) const {
    if (l != 0) throw std::logic_error("LOCA can only deal with one parameter vector.");
    return std::make_shared<const Teuchos::Array<std::string>>(p_names_);
  }
  InArgs createInArgs() const { return InArgs(); }
  OutArgs createOutArgs() const { return OutArgs(); }
  InArgs getNominalValues() const {
    InArgs a;
    a.x = initial_x_;
    a.p = p_init_;
    return a;
  }
  std::shared_ptr<Tpetra::Operator<double, int, int>> create_W_op() const {
    return std::make_shared<nosh::jacobian_operator>(mesh_, scalar_potential_, thickness_, keo_);  // :253-267
  }
  std::shared_ptr<Tpetra::Operator<double, int, int>> create_W_prec() const {
    return std::make_shared<nosh::keo_regularized>(mesh_, thickness_, mvp_);  // :301-314
  }
  // :269-299 (live settings: "Pseudo Block CG", "Preconditioner Type" = "None")
  std::shared_ptr<nosh::belos_solve_strategy> get_W_factory() const {
    return std::make_shared<nosh::belos_solve_strategy>();
  }
  const std::shared_ptr<const nosh::mesh> mesh() const { return mesh_; }

  // model_evaluator::base scalars (src/model_evaluator_base.hpp:50-60, src/model_evaluator_nls.cpp:699-770): what the
  // reference's commented code states (its live code returns 0.0); partition independent
  double inner_product(const Tpetra::Vector<double, int, int> &phi, const Tpetra::Vector<double, int, int> &psi) const {
    double r = 0.0;
    check(mesh_->ctx(), nosh_inner_product(mesh_->ctx(), phi.getData(), psi.getData(), &r));
    return r;
  }
  double norm(const Tpetra::Vector<double, int, int> &psi) const { return std::sqrt(inner_product(psi, psi)); }
  double gibbs_energy(const Tpetra::Vector<double, int, int> &psi) const {
    double r = 0.0;
    check(mesh_->ctx(), nosh_gibbs_energy(mesh_->ctx(), psi.getData(), &r));
    return r;
  }

  // evalModelImpl (:392-525)
  void evalModel(const InArgs &in, const OutArgs &out) const {
    double alpha = in.alpha, beta = in.beta;
    if (alpha == 0.0 && beta == 0.0) beta = 1.0;
    if (alpha != 0.0 || beta != 1.0) throw std::logic_error("nls::evalModel: only alpha == 0, beta == 1 supported");
    if (!in.x) throw std::logic_error("nls::evalModel: x is null");
    if (in.p.size() != p_names_.size()) throw std::logic_error("nls::evalModel: p(0) has the wrong size");
    std::map<std::string, double> params;
    for (size_t k = 0; k < p_names_.size(); k++) params[p_names_[k]] = in.p[k];
    param_list pl(params);
    nosh_ctx *c = mesh_->ctx();
    bind_fields(*mesh_, thickness_.get(), mvp_.get(), scalar_potential_.get());
    if (out.f)  // compute_f_ (:527-628)
      check(c, nosh_compute_f(c, pl.size(), pl.names.data(), pl.values.data(), in.x->getData(), out.f->getDataNonConst()));
    if (out.DfDp) {  // computeDFDP_ for every parameter (:464-489)
      if (out.DfDp->getNumVectors() != p_names_.size()) throw std::logic_error("DfDp has the wrong number of columns");
      for (size_t k = 0; k < p_names_.size(); k++)
        check(c, nosh_compute_dfdp(c, pl.size(), pl.names.data(), pl.values.data(), p_names_[k].c_str(), in.x->getData(),
                                   out.DfDp->getDataNonConst(k)));
    }
    if (out.W_op) {  // :492-504
      auto jac = std::dynamic_pointer_cast<nosh::jacobian_operator>(out.W_op);
      if (!jac) throw std::logic_error("W_op is not a nosh::jacobian_operator");
      jac->rebuild(params, *in.x);
    }
    if (out.W_prec) {  // :507-522
      auto prec = std::dynamic_pointer_cast<nosh::keo_regularized>(out.W_prec);
      if (!prec) throw std::logic_error("W_prec is not a nosh::keo_regularized");
      prec->rebuild(params, *in.x);
    }
  }

private:
  const std::shared_ptr<const nosh::mesh> mesh_;
  const std::shared_ptr<nosh::vector_field::base> mvp_;
  const std::shared_ptr<const nosh::scalar_field::base> scalar_potential_, thickness_;
  const std::shared_ptr<nosh::parameter_matrix::keo> keo_;
  const std::shared_ptr<nosh::parameter_matrix::DkeoDP> dkeo_dp_;
  const std::shared_ptr<const Tpetra::Vector<double, int, int>> initial_x_;
  std::vector<std::string> p_names_;
  std::vector<double> p_init_;
};
}  // namespace model_evaluator

// ---------------------------------------------------------------------------------------------
// src/observer.hpp / src/continuation_data_saver.hpp: what Piro/LOCA call after every continuation step -- a CSV row
// "(0) step,(1) <param>,(2) Gibbs energy,(2) ||x||_2 scaled" and an outNNNN dump of the state -- and the entry
// point that drives them: nosh-cont's LOCA run (executables/nosh-cont/nosh-cont.cpp:206-344, examples/conf.xml:35-75:
// "Arc Length" stepper, "Tangent" predictor, adaptive step size) on the device drivers.
// ---------------------------------------------------------------------------------------------
class continuation_data_saver {
public:
  explicit continuation_data_saver(const std::shared_ptr<nosh::mesh> &mesh, std::string prefix = "out")
      : mesh_(mesh), prefix_(std::move(prefix)), index_(0) {}
  void saveSolution(const Tpetra::Vector<double, int, int> &x, double /*p*/) {
    char name[32];
    std::snprintf(name, sizeof(name), "%04zu.vtk", index_);
    mesh_->write(prefix_ + name, &x);
    index_++;
  }
  size_t count() const { return index_; }

private:
  const std::shared_ptr<nosh::mesh> mesh_;
  const std::string prefix_;
  size_t index_;
};

class observer {
public:
  observer(const std::shared_ptr<const nosh::model_evaluator::nls> &model_eval, const std::string &csv_filename = "",
           const std::string &cont_param_name = "")
      : model_eval_(model_eval), csv_(csv_filename.empty() ? nullptr : std::fopen(csv_filename.c_str(), "w")),
        cont_param_name_(cont_param_name), index_(0) {}
  ~observer() {
    if (csv_) std::fclose(csv_);
  }
  observer(const observer &) = delete;
  // src/observer.cpp:134-159 (save_continuation_statistics_)
  void observeSolution(const Tpetra::Vector<double, int, int> &soln, double param_val) {
    if (!csv_) return;
    if (index_ == 0) std::fprintf(csv_, "(0) step,(1) %s,(2) Gibbs energy,(2) ||x||_2 scaled\n", cont_param_name_.c_str());
    std::fprintf(csv_, "%d,%.15e,%.15e,%.15e\n", index_, param_val, model_eval_->gibbs_energy(soln), model_eval_->norm(soln));
    std::fflush(csv_);
    index_++;
  }

private:
  const std::shared_ptr<const nosh::model_evaluator::nls> model_eval_;
  std::FILE *csv_;
  const std::string cont_param_name_;
  int index_;
};

struct continuation_options {  // examples/conf.xml:35-75
  double initial_value = 0.0, min_value = -100.0, max_value = 100.0;
  double initial_step_size = 1.0e-3, min_step_size = 1.0e-7, max_step_size = 1.0e-2, aggressiveness = 2.0;
  int max_steps = 10, max_nonlinear_iterations = 20, max_linear_iterations = 1000;
  double nonlinear_tolerance = 1.0e-8, linear_tolerance = 1.0e-10;
  // LOCA's stepper defaults, which nosh-cont inherits (examples/conf.xml sets neither)
  bool enable_arc_length_scaling = true, hit_continuation_bound = true;
};
// Arc-length continuation in `param_name` from the state x (updated in place); the other parameters keep the values
// in `params`.  observer / saver (either may be null) see every accepted step.  Returns the step records.
inline std::vector<nosh_arclength_step> continuation(const std::shared_ptr<const nosh::model_evaluator::nls> &model,
                                                     std::map<std::string, double> params, const std::string &param_name,
                                                     Tpetra::Vector<double, int, int> &x, const continuation_options &o,
                                                     nosh::observer *obs = nullptr, nosh::continuation_data_saver *saver = nullptr) {
  params[param_name] = o.initial_value;
  param_list pl(params);
  nosh_arclength_options opt;
  opt.initial_step_size = o.initial_step_size;
  opt.min_step_size = o.min_step_size;
  opt.max_step_size = o.max_step_size;
  opt.aggressiveness = o.aggressiveness;
  opt.max_steps = o.max_steps;
  opt.nl_maxit = o.max_nonlinear_iterations;
  opt.nl_tol = o.nonlinear_tolerance;
  opt.lin_tol = o.linear_tolerance;
  opt.lin_maxit = o.max_linear_iterations;
  opt.flags = (o.enable_arc_length_scaling ? NOSH_ARC_SCALING : 0) | (o.hit_continuation_bound ? NOSH_ARC_HIT_BOUND : 0);
  opt.min_value = o.min_value;
  opt.max_value = o.max_value;
  opt.goal_contribution = opt.max_contribution = opt.min_scale = opt.initial_scale = 0.0;  // LOCA's defaults
  struct hook {
    nosh::observer *obs;
    nosh::continuation_data_saver *saver;
    std::shared_ptr<const Tpetra::Map<int, int>> map;
  } h{obs, saver, model->mesh()->complex_map()};
  nosh_ctx *c = model->mesh()->ctx();
  check(c, nosh_ctx_set_step_observer(
               c,
               [](void *user, int, double param, double, double, const double *psi, int64_t n) -> int {
                 auto *hk = static_cast<hook *>(user);
                 Tpetra::Vector<double, int, int> v(hk->map);
                 std::copy(psi, psi + n, v.getDataNonConst());
                 if (hk->saver) hk->saver->saveSolution(v, param);
                 if (hk->obs) hk->obs->observeSolution(v, param);
                 return 0;
               },
               &h));
  std::vector<nosh_arclength_step> steps((size_t)o.max_steps + 2);
  int n = 0;
  const nosh_status st = nosh_continuation_arclength(c, pl.size(), pl.names.data(), pl.values.data(), param_name.c_str(), &opt,
                                                     x.getDataNonConst(), steps.data(), &n);
  nosh_ctx_set_step_observer(c, nullptr, nullptr);
  check(c, st);
  steps.resize((size_t)n);
  return steps;
}
}  // namespace nosh
