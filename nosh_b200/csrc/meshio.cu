// meshio.cu -- host-only mesh file I/O and vertex ordering ("next" row f3 of SURVEY.md section 8).
//
// The reference reads .h5m / Exodus files through MOAB (src/mesh_reader.cpp:19-162), takes the vertex
// tags "psi" (2 doubles), "A" (3 doubles), "V" from them (src/mesh.cpp:249-446: get_vector,
// get_complex_vector, get_multi_vector) and dumps states as outNNNN.h5m (mesh::write :249-263,
// src/continuation_data_saver.hpp:24-50).  MOAB, HDF5 and netCDF are not available offline, so the formats
// supported here are
//  * the legacy VTK unstructured grid (ASCII or BINARY), read and written -- what `meshio-convert in.e out.vtk`
//    produces from the reference's meshes, and a format MOAB itself reads and writes;
//  * Exodus II in the netCDF CLASSIC container (CDF-1, 64-bit-offset CDF-2, CDF-5), read only, with a reader of
//    that container written here (exodus.inc) -- the format of the reference's own test meshes
//    (test/data/*.e.md5).  Exodus files in the netCDF-4 container are HDF5 files and stay NOSH_EUNSUPPORTED,
//    like .h5m;
//  * gmsh MSH, ASCII, formats 2.x and 4.1, read only (msh.inc) -- what `gmsh -3 examples/meshes/pacman.geo`
//    writes, i.e. the reference's meshes before any conversion.
// Cells other than triangles / tetrahedra are ignored, like the reference's tri/tet dispatch
// (src/mesh_reader.cpp:127-160).
//
// nosh_morton_order: the contiguous vertex ranges of nosh_partition_range only make a good partition if
// the numbering is spatially local; this gives the permutation along a Morton (Z-order) curve through
// the bounding box -- the stand-in for the `mbpart` step of test/data/CMakeLists.txt:36-52.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/nosh_b200.h"

struct nosh_meshfile {
  int dim = 0;
  std::vector<double> coords;    // nv x 3
  std::vector<int32_t> cells;    // nc x (dim+1)
  std::map<std::string, std::pair<int, std::vector<double>>> fields;  // name -> (ncomp, nv x ncomp)
  std::string err;
};

namespace {

thread_local std::string g_err;

template <typename T>
T byteswap(T v) {
  unsigned char *p = reinterpret_cast<unsigned char *>(&v);
  std::reverse(p, p + sizeof(T));
  return v;
}

struct Reader {
  std::ifstream f;
  bool binary = false;
  uint64_t fsize = 0;  // counts in the file are bounded by it before anything is sized with them
  std::string token() {
    std::string t;
    f >> t;
    return t;
  }
  static std::string upper(std::string s) {
    for (auto &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
  }
  void skip_eol() {
    int c = f.peek();
    while (c == ' ' || c == '\t' || c == '\r') {
      f.get();
      c = f.peek();
    }
    if (c == '\n') f.get();
  }
  // n values of VTK type `type` converted to T
  template <typename T>
  bool read_array(const std::string &type, size_t n, std::vector<T> &out) {
    if (n > fsize) return false;  // a value takes at least one byte
    out.resize(n);
    const std::string ty = upper(type);
    if (!binary) {
      for (size_t i = 0; i < n; i++) {
        double v;
        if (!(f >> v)) return false;
        out[i] = (T)v;
      }
      return true;
    }
    skip_eol();
    auto rd = [&](auto tag) {
      using S = decltype(tag);
      std::vector<S> buf(n);
      f.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)(n * sizeof(S)));
      if (!f) return false;
      for (size_t i = 0; i < n; i++) out[i] = (T)byteswap(buf[i]);  // legacy VTK binary is big endian
      return true;
    };
    if (ty == "DOUBLE") return rd(double());
    if (ty == "FLOAT") return rd(float());
    if (ty == "INT") return rd(int32_t());
    if (ty == "UNSIGNED_INT") return rd(uint32_t());
    if (ty == "LONG" || ty == "VTKTYPEINT64") return rd(int64_t());
    if (ty == "UNSIGNED_CHAR") return rd((unsigned char)0);
    return false;
  }
};

nosh_status fail(nosh_status code, const std::string &msg) {
  g_err = msg;
  return code;
}

bool ends_with(const std::string &s, const char *suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && std::equal(s.end() - n, s.end(), suf, [](char a, char b) {
           return std::tolower((unsigned char)a) == std::tolower((unsigned char)b);
         });
}

}  // namespace

#include "exodus.inc"
#include "msh.inc"

extern "C" {

const char *nosh_meshfile_last_error(void) { return g_err.c_str(); }

static nosh_status meshfile_read_impl(const char *path, nosh_meshfile **out) {
  if (!path || !out) return fail(NOSH_EINVAL, "NULL argument");
  *out = nullptr;
  const std::string p(path);
  if (ends_with(p, ".h5m") || ends_with(p, ".h5"))
    return fail(NOSH_EUNSUPPORTED,
                "MOAB/HDF5 files need libraries that are not available here; convert with "
                "`meshio-convert in out.vtk` (legacy VTK) or to Exodus II in the netCDF classic container");
  if (ends_with(p, ".e") || ends_with(p, ".exo") || ends_with(p, ".ex2") || ends_with(p, ".exii") || ends_with(p, ".g") ||
      ends_with(p, ".gen"))
    return exodus_read_impl(p, out);
  if (ends_with(p, ".msh")) return msh_read_impl(p, out);
  Reader R;
  R.f.open(path, std::ios::binary);
  if (!R.f) return fail(NOSH_EINVAL, "cannot open " + p);
  R.f.seekg(0, std::ios::end);
  R.fsize = (uint64_t)R.f.tellg();
  R.f.seekg(0);
  std::string line;
  std::getline(R.f, line);
  if (line.find("# vtk DataFile") != 0) return fail(NOSH_EINVAL, p + ": not a legacy VTK file");
  std::getline(R.f, line);  // title
  std::string fmt = Reader::upper(R.token());
  if (fmt == "BINARY") R.binary = true;
  else if (fmt != "ASCII") return fail(NOSH_EINVAL, p + ": bad format line");
  if (Reader::upper(R.token()) != "DATASET" || Reader::upper(R.token()) != "UNSTRUCTURED_GRID")
    return fail(NOSH_EINVAL, p + ": only DATASET UNSTRUCTURED_GRID is supported");
  std::unique_ptr<nosh_meshfile> holder(new nosh_meshfile());  // freed on every exit path, exceptions included
  nosh_meshfile *M = holder.get();
  std::vector<int64_t> conn;  // raw CELLS stream
  std::vector<int64_t> offsets;
  bool new_layout = false;
  std::vector<int32_t> types;
  size_t ncells = 0, nv = 0;
  enum { NONE, POINT, CELL } section = NONE;
  std::string kw;
  while (R.f >> kw) {
    kw = Reader::upper(kw);
    if (kw == "POINTS") {
      size_t n;
      std::string ty;
      R.f >> n >> ty;
      nv = n;
      if (!R.read_array(ty, 3 * n, M->coords)) goto bad;
    } else if (kw == "CELLS") {
      size_t a, b;
      R.f >> a >> b;
      // classic layout: "CELLS ncells size" then (k, v0..vk-1)*; VTK >= 9: "CELLS noffsets nconn" + OFFSETS/CONNECTIVITY
      R.skip_eol();
      std::streampos pos = R.f.tellg();
      std::string nxt;
      R.f >> nxt;
      if (Reader::upper(nxt) == "OFFSETS") {
        new_layout = true;
        std::string ty;
        R.f >> ty;
        if (!R.read_array(ty, a, offsets)) goto bad;
        R.f >> nxt >> ty;  // CONNECTIVITY type
        if (!R.read_array(ty, b, conn)) goto bad;
        ncells = a - 1;
      } else {
        R.f.seekg(pos);
        ncells = a;
        if (!R.read_array(std::string("int"), b, conn)) goto bad;
      }
    } else if (kw == "CELL_TYPES") {
      size_t n;
      R.f >> n;
      if (!R.read_array(std::string("int"), n, types)) goto bad;
    } else if (kw == "POINT_DATA") {
      size_t n;
      R.f >> n;
      section = POINT;
    } else if (kw == "CELL_DATA") {
      size_t n;
      R.f >> n;
      section = CELL;
    } else if (kw == "SCALARS") {
      std::string name, ty;
      R.f >> name >> ty;
      int ncomp = 1;
      R.skip_eol();
      std::streampos pos = R.f.tellg();
      std::string t;
      R.f >> t;
      if (Reader::upper(t) != "LOOKUP_TABLE") {
        ncomp = std::atoi(t.c_str());
        R.f >> t;  // LOOKUP_TABLE
      }
      (void)pos;
      R.f >> t;  // table name
      std::vector<double> v;
      const size_t cnt = (section == CELL ? ncells : nv) * (size_t)ncomp;
      if (!R.read_array(ty, cnt, v)) goto bad;
      if (section == POINT) M->fields[name] = {ncomp, std::move(v)};
    } else if (kw == "VECTORS" || kw == "NORMALS") {
      std::string name, ty;
      R.f >> name >> ty;
      std::vector<double> v;
      const size_t cnt = (section == CELL ? ncells : nv) * 3;
      if (!R.read_array(ty, cnt, v)) goto bad;
      if (section == POINT) M->fields[name] = {3, std::move(v)};
    } else if (kw == "FIELD") {
      std::string fname;
      int narr;
      R.f >> fname >> narr;
      for (int a = 0; a < narr; a++) {
        std::string name, ty;
        int ncomp;
        size_t ntup;
        R.f >> name >> ncomp >> ntup >> ty;
        std::vector<double> v;
        if (!R.read_array(ty, ntup * (size_t)ncomp, v)) goto bad;
        if (section == POINT && ntup == nv) M->fields[name] = {ncomp, std::move(v)};
      }
    } else if (kw == "METADATA") {
      // skip the INFORMATION block (up to the next blank line)
      std::getline(R.f, line);
      while (std::getline(R.f, line) && !line.empty() && line != "\r") {
      }
    } else if (kw == "LOOKUP_TABLE") {
      std::string name;
      size_t n;
      R.f >> name >> n;
      std::vector<double> v;
      if (!R.read_array(std::string(R.binary ? "unsigned_char" : "float"), 4 * n, v)) goto bad;
    } else {
      holder.reset();
      return fail(NOSH_EINVAL, p + ": unsupported VTK section " + kw);
    }
  }
  {
    if (types.size() != ncells) {
      holder.reset();
      return fail(NOSH_EINVAL, p + ": CELL_TYPES does not match CELLS");
    }
    // keep the highest-dimensional simplex type present (tetrahedra, else triangles)
    bool has_tet = false;
    for (int t : types) has_tet = has_tet || t == 10;
    M->dim = has_tet ? 3 : 2;
    const int want = has_tet ? 10 : 5, nvc = M->dim + 1;
    size_t pos = 0;
    for (size_t c = 0; c < ncells; c++) {
      size_t k, start;
      if (new_layout) {
        start = (size_t)offsets[c];
        k = (size_t)(offsets[c + 1] - offsets[c]);
      } else {
        if (pos >= conn.size()) goto bad2;
        k = (size_t)conn[pos];
        start = pos + 1;
        pos += k + 1;
      }
      if (types[c] != want) continue;
      if (k != (size_t)nvc || start + k > conn.size()) goto bad2;
      for (size_t i = 0; i < k; i++) {
        const int64_t v = conn[start + i];
        if (v < 0 || (size_t)v >= nv) goto bad2;
        M->cells.push_back((int32_t)v);
      }
    }
    if (M->cells.empty()) {
      holder.reset();
      return fail(NOSH_EMESH, p + ": no triangles or tetrahedra");
    }
    *out = holder.release();
    return NOSH_OK;
  }
bad2:
  holder.reset();
  return fail(NOSH_EINVAL, p + ": inconsistent cell connectivity");
bad:
  holder.reset();
  return fail(NOSH_EINVAL, p + ": truncated or malformed data array");
}

void nosh_meshfile_free(nosh_meshfile *m) { delete m; }

nosh_status nosh_meshfile_info(const nosh_meshfile *m, int32_t *dim, int64_t *n_vertices, int64_t *n_cells,
                               int32_t *n_fields) {
  if (!m) return fail(NOSH_EINVAL, "NULL mesh file");
  if (dim) *dim = m->dim;
  if (n_vertices) *n_vertices = (int64_t)(m->coords.size() / 3);
  if (n_cells) *n_cells = (int64_t)(m->cells.size() / (size_t)(m->dim + 1));
  if (n_fields) *n_fields = (int32_t)m->fields.size();
  return NOSH_OK;
}

nosh_status nosh_meshfile_get(const nosh_meshfile *m, double *coords, int32_t *cells) {
  if (!m) return fail(NOSH_EINVAL, "NULL mesh file");
  if (coords) std::copy(m->coords.begin(), m->coords.end(), coords);
  if (cells) std::copy(m->cells.begin(), m->cells.end(), cells);
  return NOSH_OK;
}

nosh_status nosh_meshfile_field_name(const nosh_meshfile *m, int32_t index, const char **name, int32_t *ncomp) {
  if (!m || index < 0 || index >= (int32_t)m->fields.size()) return fail(NOSH_EINVAL, "bad field index");
  auto it = m->fields.begin();
  std::advance(it, index);
  if (name) *name = it->first.c_str();
  if (ncomp) *ncomp = it->second.first;
  return NOSH_OK;
}

nosh_status nosh_meshfile_get_field(const nosh_meshfile *m, const char *name, int32_t *ncomp, double *values) {
  if (!m || !name) return fail(NOSH_EINVAL, "NULL argument");
  auto it = m->fields.find(name);
  if (it == m->fields.end()) return fail(NOSH_EKEY, std::string("no vertex tag \"") + name + "\" in the file");
  if (ncomp) *ncomp = it->second.first;
  if (values) std::copy(it->second.second.begin(), it->second.second.end(), values);
  return NOSH_OK;
}

static nosh_status meshfile_write_impl(const char *path, int32_t dim, int64_t n_vertices, const double *coords,
                                      int64_t n_cells, const int32_t *cells, int32_t n_fields,
                                      const char *const *names, const int32_t *ncomps, const double *const *values,
                                      int32_t binary) {
  if (!path || !coords || !cells || (dim != 2 && dim != 3) || n_vertices <= 0 || n_cells <= 0 ||
      (n_fields > 0 && (!names || !ncomps || !values)))
    return fail(NOSH_EINVAL, "bad argument");
  std::ofstream f(path, std::ios::binary);
  if (!f) return fail(NOSH_EINVAL, std::string("cannot open ") + path + " for writing");
  const int nvc = dim + 1;
  f << "# vtk DataFile Version 3.0\nnosh_b200 state\n" << (binary ? "BINARY" : "ASCII") << "\nDATASET UNSTRUCTURED_GRID\n";
  f.precision(17);
  auto put_d = [&](const double *v, size_t n, int per_line) {
    if (binary) {
      std::vector<double> b(n);
      for (size_t i = 0; i < n; i++) b[i] = byteswap(v[i]);
      f.write(reinterpret_cast<const char *>(b.data()), (std::streamsize)(n * sizeof(double)));
      f << "\n";
    } else {
      for (size_t i = 0; i < n; i++) f << v[i] << (((i + 1) % per_line == 0) ? "\n" : " ");
      if (n % per_line) f << "\n";
    }
  };
  auto put_i = [&](const std::vector<int32_t> &v, int per_line) {
    if (binary) {
      std::vector<int32_t> b(v.size());
      for (size_t i = 0; i < v.size(); i++) b[i] = byteswap(v[i]);
      f.write(reinterpret_cast<const char *>(b.data()), (std::streamsize)(b.size() * sizeof(int32_t)));
      f << "\n";
    } else {
      for (size_t i = 0; i < v.size(); i++) f << v[i] << (((i + 1) % per_line == 0) ? "\n" : " ");
      if (v.size() % per_line) f << "\n";
    }
  };
  f << "POINTS " << n_vertices << " double\n";
  put_d(coords, (size_t)n_vertices * 3, 3);
  f << "CELLS " << n_cells << " " << n_cells * (nvc + 1) << "\n";
  {
    std::vector<int32_t> c((size_t)n_cells * (nvc + 1));
    for (int64_t i = 0; i < n_cells; i++) {
      c[(size_t)i * (nvc + 1)] = nvc;
      for (int k = 0; k < nvc; k++) c[(size_t)i * (nvc + 1) + 1 + k] = cells[i * nvc + k];
    }
    put_i(c, nvc + 1);
  }
  f << "CELL_TYPES " << n_cells << "\n";
  put_i(std::vector<int32_t>((size_t)n_cells, dim == 3 ? 10 : 5), 1);
  if (n_fields > 0) {
    f << "POINT_DATA " << n_vertices << "\n";
    for (int a = 0; a < n_fields; a++) {
      if (ncomps[a] == 3) {
        f << "VECTORS " << names[a] << " double\n";
      } else {
        f << "SCALARS " << names[a] << " double " << ncomps[a] << "\nLOOKUP_TABLE default\n";
      }
      put_d(values[a], (size_t)n_vertices * (size_t)ncomps[a], ncomps[a]);
    }
  }
  f.flush();
  if (!f) return fail(NOSH_EINVAL, std::string("write to ") + path + " failed");
  return NOSH_OK;
}

static nosh_status morton_order_impl(int64_t n_vertices, const double *coords, int64_t *perm) {
  if (n_vertices <= 0 || !coords || !perm) return fail(NOSH_EINVAL, "bad argument");
  double lo[3], hi[3];
  for (int d = 0; d < 3; d++) lo[d] = hi[d] = coords[d];
  for (int64_t i = 0; i < n_vertices; i++)
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], coords[3 * i + d]);
      hi[d] = std::max(hi[d], coords[3 * i + d]);
    }
  double ext = 0.0;
  for (int d = 0; d < 3; d++) ext = std::max(ext, hi[d] - lo[d]);
  if (!(ext > 0.0)) ext = 1.0;
  auto spread = [](uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
  };
  std::vector<std::pair<uint64_t, int64_t>> keys((size_t)n_vertices);
  for (int64_t i = 0; i < n_vertices; i++) {
    uint64_t q[3];
    for (int d = 0; d < 3; d++) {
      double t = (coords[3 * i + d] - lo[d]) / ext;  // same scale on every axis: cubes, not bricks
      t = std::min(std::max(t, 0.0), 1.0);
      q[d] = (uint64_t)std::min(2097151.0, std::floor(t * 2097152.0));
    }
    keys[(size_t)i] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), i};
  }
  std::sort(keys.begin(), keys.end());
  for (int64_t i = 0; i < n_vertices; i++) perm[i] = keys[(size_t)i].second;  // new position i <- old vertex perm[i]
  return NOSH_OK;
}

// nothing may throw across the C boundary (std::bad_alloc on a huge file, stream failures)
#define GUARDED(call)                                   \
  try {                                                 \
    return call;                                        \
  } catch (const std::exception &e) {                   \
    return fail(NOSH_EINVAL, std::string(e.what()));    \
  } catch (...) {                                       \
    return fail(NOSH_EINVAL, "unknown failure");        \
  }
nosh_status nosh_meshfile_read(const char *path, nosh_meshfile **out) { GUARDED(meshfile_read_impl(path, out)) }
nosh_status nosh_meshfile_write(const char *path, int32_t dim, int64_t n_vertices, const double *coords,
                                int64_t n_cells, const int32_t *cells, int32_t n_fields, const char *const *names,
                                const int32_t *ncomps, const double *const *values, int32_t binary) {
  GUARDED(meshfile_write_impl(path, dim, n_vertices, coords, n_cells, cells, n_fields, names, ncomps, values, binary))
}
nosh_status nosh_morton_order(int64_t n_vertices, const double *coords, int64_t *perm) {
  GUARDED(morton_order_impl(n_vertices, coords, perm))
}

}  // extern "C"
