// apply.cu -- see apply.cuh
#include "apply.cuh"

#include "kmath.cuh"

namespace nosh {

namespace {

template <int EPI>
__device__ __forceinline__ double2 epilogue(double2 acc, double2 xi, int64_t i, const ApplyArgs &A) {
  if (EPI == EPI_DIAG) {
    acc = diag_epilogue(acc, ld_stream2(A.d0 + i), __ldg(A.d1 + i), xi);
  } else if (EPI == EPI_F) {
    const double al = A.cv[i] * A.thick[i] * (A.V[i] + A.g * (xi.x * xi.x + xi.y * xi.y));
    acc.x += al * xi.x;
    acc.y += al * xi.y;
  } else if (EPI == EPI_DG) {
    const double al = A.cv[i] * A.thick[i] * (xi.x * xi.x + xi.y * xi.y);
    acc.x += al * xi.x;
    acc.y += al * xi.y;
  } else if (EPI == EPI_DV) {
    const double al = A.cv[i] * A.thick[i] * A.V[i];
    acc.x += al * xi.x;
    acc.y += al * xi.y;
  }
  return acc;
}

// AMG smoother tails on the finest level (amg.cu): yi = A x on entry
template <int FUSE>
__device__ __forceinline__ double2 smoother_tail(double2 yi, double2 xi, int64_t row, const ApplyArgs &A) {
  const double2 bb = ld_stream2(A.bvec + row);
  if (FUSE == FUSE_RESID) return make_double2(bb.x - yi.x, bb.y - yi.y);
  const double2 di = ld_stream2(A.dinv + row);
  double2 dd = make_double2(A.c2 * (di.x * (bb.x - yi.x)), A.c2 * (di.y * (bb.y - yi.y)));
  if (A.c1 != 0.0) {
    const double2 dold = A.dvec[row];
    dd.x += A.c1 * dold.x;
    dd.y += A.c1 * dold.y;
  }
  if (A.dvec) A.dvec[row] = dd;
  return make_double2(xi.x + dd.x, xi.y + dd.y);
}

template <int FUSE>
__device__ __forceinline__ bool krylov_skip(const ApplyArgs &A) {
  if (FUSE == FUSE_MINRES || FUSE == FUSE_CG) {
    // device-side control flow: after convergence the remaining launches are no-ops
    return A.st->done || A.st->iter != A.host_iter - 1;
  }
  return false;
}

// -------------------------------------------------------------------------------------------
// SELL-32, one thread per row, one CTA per CHUNK (=512) rows = 16 slices.
// -------------------------------------------------------------------------------------------
// U: independent (column, value) pairs in flight per thread; MINB: CTAs per SM the register budget is
// sized for (2 -> 64 registers, 1 -> up to 128); PF: load the next batch's columns before the current
// batch's gathers are consumed (software pipelining of the col -> x dependency).  The variants exist
// for measurement (profiles/apply_variants.py); launch2 picks the one that measured best.
// GH: ghost columns (c >= No) are read from A.xg, the landing slot the neighbours stored into over NVLink
// (L2 loads: written by another device), instead of from x[No..).
template <bool GH>
__device__ __forceinline__ double2 gather_x(const ApplyArgs &A, int c) {
  if (GH) {
    // branch-free: one pointer select, one load instruction for both cases (a branch per gather would serialise
    // the U independent loads a thread keeps in flight).  Plain (coherent) load: the ghosts land DURING this
    // kernel, after the boundary CTAs' wait + fence, so the non-coherent path of __ldg is not allowed for them.
    const double2 *p = c < A.No ? A.x + c : A.xg + (c - A.No);
    return *p;
  }
  return __ldg(A.x + c);
}

template <bool F32>
struct ValT {
  using type = double2;
  static __device__ __forceinline__ type load(const ApplyArgs &A, int p) { return ld_stream2(A.val + p); }
};
template <>
struct ValT<true> {
  using type = float2;
  static __device__ __forceinline__ type load(const ApplyArgs &A, int p) { return ld_stream_f2(A.val32 + p); }
};

template <int EPI, int FUSE, int U, bool PF, bool GH, bool F32 = false>
__device__ __forceinline__ void apply_sell_cta(const ApplyArgs &A, double *red) {
  const int chunk = A.chunk_list ? __ldg(A.chunk_list + blockIdx.x) : (int)blockIdx.x;
  const int64_t pos = (int64_t)chunk * CHUNK + threadIdx.x;  // SELL position; the row stored there:
  const int64_t row = A.sell_row ? (int64_t)__ldg(A.sell_row + pos) : pos;
  const int64_t slice = pos >> 5;
  const int lane = threadIdx.x & 31;
  const double scale = (FUSE == FUSE_MINRES) ? A.st->inv_beta : 1.0;
  double2 acc = make_double2(0.0, 0.0);
  if (slice < A.nslices) {
    int p = __ldg(A.slice_off + slice) + lane;
    const int pend = __ldg(A.slice_off + slice + 1);
    if (PF) {
      int c[U];
      bool have = p + 32 * (U - 1) < pend;
      if (have) {
#pragma unroll
        for (int u = 0; u < U; u++) c[u] = ld_stream_i32(A.col + p + 32 * u);
      }
      while (have) {
        typename ValT<F32>::type v[U];
        double2 xv[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ValT<F32>::load(A, p + 32 * u);
#pragma unroll
        for (int u = 0; u < U; u++) xv[u] = gather_x<GH>(A, c[u]);
        p += 32 * U;
        have = p + 32 * (U - 1) < pend;
        if (have) {
#pragma unroll
          for (int u = 0; u < U; u++) c[u] = ld_stream_i32(A.col + p + 32 * u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (FUSE == FUSE_MINRES) {
            xv[u] = scaled(xv[u], scale);
          }
          cfma(acc, v[u], xv[u]);
        }
      }
    } else {
      // U independent (column, value) loads, then U gathers, then the FMAs
      for (; p + 32 * (U - 1) < pend; p += 32 * U) {
        int c[U];
        typename ValT<F32>::type v[U];
        double2 xv[U];
#pragma unroll
        for (int u = 0; u < U; u++) c[u] = ld_stream_i32(A.col + p + 32 * u);
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ValT<F32>::load(A, p + 32 * u);
#pragma unroll
        for (int u = 0; u < U; u++) xv[u] = gather_x<GH>(A, c[u]);
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (FUSE == FUSE_MINRES) {
            xv[u] = scaled(xv[u], scale);
          }
          cfma(acc, v[u], xv[u]);
        }
      }
    }
    if (!PF && p < pend) {
      // remainder: ONE masked batch (same ascending order => same bits) instead of a loop of single, dependent
      // col -> x loads; masked entries gather x[0], a valid address, and are never used
      int c[U];
      typename ValT<F32>::type v[U];
      double2 xv[U];
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = p + 32 * u < pend ? ld_stream_i32(A.col + p + 32 * u) : 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p + 32 * u < pend) v[u] = ValT<F32>::load(A, p + 32 * u);
        else v[u] = typename ValT<F32>::type();
      }
#pragma unroll
      for (int u = 0; u < U; u++) xv[u] = gather_x<GH>(A, c[u]);
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (p + 32 * u < pend) {
          if (FUSE == FUSE_MINRES) {
            xv[u] = scaled(xv[u], scale);
          }
          cfma(acc, v[u], xv[u]);
        }
      }
      p = pend;
    }
    for (; p < pend; p += 32) {  // PF variants only
      const int c = ld_stream_i32(A.col + p);
      const typename ValT<F32>::type v = ValT<F32>::load(A, p);
      double2 xv = gather_x<GH>(A, c);
      if (FUSE == FUSE_MINRES) {
        xv = scaled(xv, scale);
      }
      cfma(acc, v, xv);
    }
  }
  double contrib = 0.0;
  if (row < A.No) {
    double2 xi = __ldg(A.x + row);
    if (FUSE == FUSE_MINRES) {
      xi = scaled(xi, scale);
    }
    double2 yi = epilogue<EPI>(acc, xi, row, A);
    if (FUSE == FUSE_AXPBY) {
      yi.x *= A.a;
      yi.y *= A.a;
      if (A.b != 0.0) {
        const double2 yo = A.y[row];
        yi.x += A.b * yo.x;
        yi.y += A.b * yo.y;
      }
    } else if (FUSE == FUSE_MINRES) {
      const double f = A.st->f_r1;
      if (f != 0.0) {
        yi = sub_scaled(yi, f, ld_stream2(A.r1 + row));
      }
      contrib = cdot(xi, yi);
    } else if (FUSE == FUSE_CG) {
      contrib = cdot(xi, yi);
    } else if (FUSE == FUSE_RESID || FUSE == FUSE_CHEB) {
      yi = smoother_tail<FUSE>(yi, xi, row, A);
    }
    A.y[row] = yi;
  }
  if (FUSE == FUSE_MINRES || FUSE == FUSE_CG) {
    const double s = block_sum<CHUNK / 32>(contrib, red);
    if (threadIdx.x == 0) A.partials[chunk] = s;
    last_cta_finalize(A.fin, red);
  }
}

// GHK: the launch carries a separate ghost vector (A.xg).  Only the CTAs at or behind halo_first_block -- the
// chunks that reference ghosts, listed last -- wait for the neighbours' flags and use the ghost-aware gather;
// all other CTAs run the plain body (the pointer select per gather costs ~15 % on the interior rows: measured).
template <int EPI, int FUSE, int U, int MINB, bool PF, bool GHK = false, bool F32 = false>
__global__ void __launch_bounds__(CHUNK, MINB) k_apply_sell(const ApplyArgs A) {
  if (krylov_skip<FUSE>(A)) return;
  if (A.gate && A.gate->done) return;
  __shared__ double red[32];
  if (GHK) {
    if ((int)blockIdx.x < A.push_blocks) halo_push_cta(A.halo, A.x, A.send_idx, A.n_send, A.halo_epoch, A.push_blocks);
    if ((int)blockIdx.x >= A.halo_first_block) {
      if (A.halo) halo_wait_cta(A.halo, A.halo_epoch);
      apply_sell_cta<EPI, FUSE, U, PF, true>(A, red);
      return;
    }
  }
  apply_sell_cta<EPI, FUSE, U, PF, false, F32>(A, red);
}

// -------------------------------------------------------------------------------------------
// block-CSR, LPR lanes per row, one CTA per CHUNK rows.
// -------------------------------------------------------------------------------------------
template <int EPI, int FUSE, int LPR>
__global__ void __launch_bounds__(CHUNK, 2) k_apply_csr(const ApplyArgs A) {
  if (krylov_skip<FUSE>(A)) return;
  if (A.gate && A.gate->done) return;
  __shared__ double red[32];
  if (A.xg && A.halo && (int)blockIdx.x >= A.halo_first_block) halo_wait_cta(A.halo, A.halo_epoch);
  constexpr int RPP = CHUNK / LPR;  // rows per pass
  const int chunk = A.chunk_list ? __ldg(A.chunk_list + blockIdx.x) : (int)blockIdx.x;
  const int sub = threadIdx.x / LPR, sl = threadIdx.x % LPR;
  const double scale = (FUSE == FUSE_MINRES) ? A.st->inv_beta : 1.0;
  double contrib = 0.0;
#pragma unroll 2
  for (int pass = 0; pass < LPR; pass++) {
    const int64_t row = (int64_t)chunk * CHUNK + pass * RPP + sub;
    double2 acc = make_double2(0.0, 0.0);
    if (row < A.No) {
      const int b = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
      for (int p = b + sl; p < e; p += LPR) {
        const int c = ld_stream_i32(A.col + p);
        const double2 v = ld_stream2(A.val + p);
        double2 xv = (A.xg && c >= A.No) ? __ldcg(A.xg + (c - A.No)) : __ldg(A.x + c);
        if (FUSE == FUSE_MINRES) {
          xv = scaled(xv, scale);
        }
        cfma(acc, v, xv);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (sl == 0 && row < A.No) {
      double2 xi = __ldg(A.x + row);
      if (FUSE == FUSE_MINRES) {
        xi = scaled(xi, scale);
      }
      double2 yi = epilogue<EPI>(acc, xi, row, A);
      if (FUSE == FUSE_AXPBY) {
        yi.x *= A.a;
        yi.y *= A.a;
        if (A.b != 0.0) {
          const double2 yo = A.y[row];
          yi.x += A.b * yo.x;
          yi.y += A.b * yo.y;
        }
      } else if (FUSE == FUSE_MINRES) {
        const double f = A.st->f_r1;
        if (f != 0.0) {
          yi = sub_scaled(yi, f, ld_stream2(A.r1 + row));
        }
        contrib += cdot(xi, yi);
      } else if (FUSE == FUSE_CG) {
        contrib += cdot(xi, yi);
      } else if (FUSE == FUSE_RESID || FUSE == FUSE_CHEB) {
        yi = smoother_tail<FUSE>(yi, xi, row, A);
      }
      A.y[row] = yi;
    }
  }
  if (FUSE == FUSE_MINRES || FUSE == FUSE_CG) {
    const double s = block_sum<CHUNK / 32>(contrib, red);
    if (threadIdx.x == 0) A.partials[chunk] = s;
    last_cta_finalize(A.fin, red);
  }
}

// pairs in flight per thread of the fp32-value variant: 12 B per pair instead of 20, so more of them are needed to
// keep the same bytes in flight (profiles/r2_summary.md section 8)
constexpr int F32_U = 5;
// ... and of every fp64 kernel: 5 divides the 15 blocks per row of a Kuhn tet mesh (three full batches, no
// remainder); other row lengths end in one masked batch (apply_sell_cta).  Round 1 used 4.
constexpr int DEF_U = 5;

template <int EPI, int FUSE>
void launch2(Ctx *ctx, const ApplyArgs &A) {
  const unsigned grid = A.chunk_list ? (unsigned)A.n_list : (unsigned)cdiv(A.No, CHUNK);
  if (grid == 0) return;
  constexpr bool ghost_ok = FUSE == FUSE_NONE || FUSE == FUSE_AXPBY;
  if (A.xg && !ghost_ok) NOSH_THROW(NOSH_EINVAL, "internal: separate ghost vector with a fused Krylov/smoother apply");
  constexpr bool f32_ok = EPI == EPI_DIAG && (FUSE == FUSE_RESID || FUSE == FUSE_CHEB);
  if (A.val32 && !(f32_ok && ctx->layout == NOSH_LAYOUT_SELL32)) NOSH_THROW(NOSH_EINVAL, "internal: fp32 values with this apply variant");
  if (A.val32) {
    if constexpr (f32_ok) {
      switch (ctx->apply_variant) {  // measurement variants, like the fp64 kernels below
        case 1: k_apply_sell<EPI, FUSE, 4, 2, false, false, true><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
        case 2: k_apply_sell<EPI, FUSE, 5, 2, false, false, true><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
        case 3: k_apply_sell<EPI, FUSE, 8, 2, false, false, true><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
        default: k_apply_sell<EPI, FUSE, F32_U, 2, false, false, true><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      }
    }
  } else if (ctx->layout == NOSH_LAYOUT_SELL32 && A.xg) {
    if constexpr (ghost_ok) k_apply_sell<EPI, FUSE, DEF_U, 2, false, true><<<grid, CHUNK, 0, ctx->stream>>>(A);
  } else if (ctx->layout == NOSH_LAYOUT_SELL32) {
    // measurement variants exist for the two kernels of the MINRES loop only (compile time)
    constexpr bool tunable = EPI == EPI_DIAG && (FUSE == FUSE_NONE || FUSE == FUSE_MINRES);
    const int v = tunable ? ctx->apply_variant : 0;
    switch (v) {
      case 1: k_apply_sell<EPI, FUSE, tunable ? 8 : DEF_U, 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 2: k_apply_sell<EPI, FUSE, tunable ? 4 : DEF_U, 2, tunable><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 3: k_apply_sell<EPI, FUSE, tunable ? 8 : DEF_U, tunable ? 1 : 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 4: k_apply_sell<EPI, FUSE, tunable ? 8 : DEF_U, tunable ? 1 : 2, tunable><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 5: k_apply_sell<EPI, FUSE, tunable ? 6 : DEF_U, 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 6: k_apply_sell<EPI, FUSE, tunable ? 4 : DEF_U, 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      case 7: k_apply_sell<EPI, FUSE, tunable ? 12 : DEF_U, tunable ? 1 : 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
      default: k_apply_sell<EPI, FUSE, DEF_U, 2, false><<<grid, CHUNK, 0, ctx->stream>>>(A); break;
    }
  } else
    k_apply_csr<EPI, FUSE, 8><<<grid, CHUNK, 0, ctx->stream>>>(A);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}

template <int EPI>
void launch1(Ctx *ctx, int fuse, const ApplyArgs &A) {
  switch (fuse) {
    case FUSE_NONE: launch2<EPI, FUSE_NONE>(ctx, A); break;
    case FUSE_AXPBY: launch2<EPI, FUSE_AXPBY>(ctx, A); break;
    case FUSE_MINRES: launch2<EPI, FUSE_MINRES>(ctx, A); break;
    case FUSE_CG: launch2<EPI, FUSE_CG>(ctx, A); break;
    default: NOSH_THROW(NOSH_EINVAL, "bad fuse mode");
  }
}

}  // namespace

void launch_apply(Ctx *ctx, int epi, int fuse, const ApplyArgs &A) {
  switch (epi) {
    case EPI_NONE: launch1<EPI_NONE>(ctx, fuse, A); break;
    case EPI_DIAG:
      if (fuse == FUSE_RESID) launch2<EPI_DIAG, FUSE_RESID>(ctx, A);
      else if (fuse == FUSE_CHEB) launch2<EPI_DIAG, FUSE_CHEB>(ctx, A);
      else launch1<EPI_DIAG>(ctx, fuse, A);
      break;
    case EPI_F: launch2<EPI_F, FUSE_NONE>(ctx, A); break;
    case EPI_DG: launch2<EPI_DG, FUSE_NONE>(ctx, A); break;
    case EPI_DV: launch2<EPI_DV, FUSE_NONE>(ctx, A); break;
    default: NOSH_THROW(NOSH_EINVAL, "bad epilogue");
  }
}

}  // namespace nosh
