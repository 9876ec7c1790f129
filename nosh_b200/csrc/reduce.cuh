// reduce.cuh -- levels 2 and 3 of the fixed reduction tree (common.cuh) plus the scalar recurrences
// of the Krylov solvers, shared by the stand-alone k_finalize kernel (multi-GPU) and by the
// "last CTA" tail of the producing kernels (single GPU: no extra launch per reduction).
// The tree shape does not depend on the CTA size, so both paths give the same bits.
#pragma once
#include "common.cuh"

namespace nosh {

enum { FIN_DOT = 0, FIN_MINRES_INIT, FIN_MINRES_ALPHA, FIN_MINRES_BETA, FIN_CG_INIT, FIN_CG_PAP, FIN_CG_RHO,
       FIN_PCG_INIT_RHO, FIN_PCG_RR, FIN_PCG_RHO };

struct FinArgs {
  const double *partials;
  int64_t n_chunks;
  int cpg;
  int64_t group_begin, n_groups_local;
  int n_groups_global;
  double *gsend;        // MAX_GROUPS, global group index; zero outside the local groups
  const double *grecv;  // all-reduced copy (== gsend on one GPU)
  KrylovState *st;
  double *out;
  double *hist;
  int what, host_iter, stage;  // stage 0: single GPU; 1: level 2 only; 2: level 3 + scalars;
                               // 3: one kernel, group sums exchanged through peer memory
  double tol;
  int maxit;
  P2PView p2p;
  unsigned int *counter;  // last-CTA ticket (single GPU): the producing kernel finishes the reduction itself
};

static __device__ __forceinline__ void sym_ortho(double a, double b, double &c, double &s, double &r) {
  const double absA = fabs(a), absB = fabs(b);
  if (absB == 0.0) {
    s = 0.0;
    r = absA;
    c = (absA == 0.0) ? 1.0 : (a >= 0.0 ? 1.0 : -1.0);
  } else if (absA == 0.0) {
    c = 0.0;
    s = (b >= 0.0 ? 1.0 : -1.0);
    r = absB;
  } else if (absB >= absA) {
    const double tau = a / b;
    s = (b >= 0.0 ? 1.0 : -1.0) / sqrt(1.0 + tau * tau);
    c = s * tau;
    r = b / s;
  } else {
    const double tau = b / a;
    c = (a >= 0.0 ? 1.0 : -1.0) / sqrt(1.0 + tau * tau);
    s = c * tau;
    r = a / c;
  }
}

static __device__ __forceinline__ void fin_scalars(const FinArgs &F, double total) {
  KrylovState *st = F.st;
  switch (F.what) {
    case FIN_DOT:
      F.out[0] = total;
      break;
    case FIN_MINRES_INIT: {
      KrylovState s;
      memset(&s, 0, sizeof(s));
      s.tol = F.tol;
      s.maxit = F.maxit;
      if (total <= 0.0) {
        s.done = 1;
        s.converged = 1;
        s.inv_beta = 0.0;
      } else {
        s.beta1 = sqrt(total);
        s.beta = s.beta1;
        s.phibar = s.beta1;
        s.cs = -1.0;
        s.sn = 0.0;
        s.inv_beta = 1.0 / s.beta1;
        s.relres = 1.0;
        if (F.maxit <= 0 || 1.0 <= F.tol) {
          s.done = 1;
          s.converged = 1.0 <= F.tol;
        }
      }
      *st = s;
      if (F.hist) F.hist[0] = 1.0;
      break;
    }
    case FIN_MINRES_ALPHA:
      st->alpha = total;
      st->f_r2 = total / st->beta;
      break;
    case FIN_MINRES_BETA: {
      if (total < 0.0) {
        st->done = 1;
        st->breakdown = 1;
        break;
      }
      const double betaNew = sqrt(total);
      const double oldeps = st->epsln;
      const double delta = st->cs * st->dbar + st->sn * st->alpha;
      const double gbar = st->sn * st->dbar - st->cs * st->alpha;
      st->epsln = st->sn * betaNew;
      st->dbar = -st->cs * betaNew;
      double cs, sn, gamma;
      sym_ortho(gbar, betaNew, cs, sn, gamma);
      st->cs = cs;
      st->sn = sn;
      st->gamma = gamma;
      st->gbar = gbar;
      st->phi = cs * st->phibar;
      st->phibar = sn * st->phibar;
      if (gamma == 0.0) {
        st->done = 1;
        st->breakdown = 1;
        break;
      }
      st->oldeps = oldeps;
      st->delta = delta;
      st->inv_gamma = 1.0 / gamma;
      st->inv_beta_prev = st->inv_beta;
      st->oldBeta = st->beta;
      st->beta = betaNew;
      st->inv_beta = 1.0 / betaNew;
      st->f_r1 = betaNew / st->oldBeta;
      st->iter += 1;
      st->relres = st->phibar / st->beta1;
      if (F.hist) F.hist[st->iter] = st->relres;
      if (st->relres <= st->tol) {
        st->done = 1;
        st->converged = 1;
      } else if (st->iter >= st->maxit) {
        st->done = 1;
      }
      break;
    }
    case FIN_CG_INIT: {
      KrylovState s;
      memset(&s, 0, sizeof(s));
      s.tol = F.tol;
      s.maxit = F.maxit;
      s.rho = total;
      s.r0norm = sqrt(total);
      s.relres = 1.0;
      if (s.r0norm == 0.0) {
        s.done = 1;
        s.converged = 1;
        s.relres = 0.0;
      } else if (F.maxit <= 0 || 1.0 <= F.tol) {
        s.done = 1;
        s.converged = 1.0 <= F.tol;
      }
      *st = s;
      if (F.hist) F.hist[0] = 1.0;
      break;
    }
    case FIN_CG_PAP:
      st->pAp = total;
      st->cg_alpha = st->rho / total;
      break;
    case FIN_PCG_INIT_RHO:  // preconditioned CG: rho_0 = <r_0, M r_0> (r0norm comes from FIN_CG_INIT)
      st->rho = total;
      break;
    case FIN_PCG_RR: {  // ||r_k|| -> convergence test; rho / beta follow in FIN_PCG_RHO
      st->iter += 1;
      st->relres = sqrt(total) / st->r0norm;
      if (F.hist) F.hist[st->iter] = st->relres;
      if (st->relres <= st->tol) {
        st->done = 1;
        st->converged = 1;
      } else if (st->iter >= st->maxit) {
        st->done = 1;
      }
      break;
    }
    case FIN_PCG_RHO:
      st->cg_beta = total / st->rho;
      st->rho = total;
      break;
    case FIN_CG_RHO: {
      st->cg_beta = total / st->rho;
      st->rho = total;
      st->iter += 1;
      st->relres = sqrt(total) / st->r0norm;
      if (F.hist) F.hist[st->iter] = st->relres;
      if (st->relres <= st->tol) {
        st->done = 1;
        st->converged = 1;
      } else if (st->iter >= st->maxit) {
        st->done = 1;
      }
      break;
    }
  }
}


static __device__ __forceinline__ bool fin_is_iterative(int what) {
  return what == FIN_MINRES_ALPHA || what == FIN_MINRES_BETA || what == FIN_CG_PAP || what == FIN_CG_RHO ||
         what == FIN_PCG_RR || what == FIN_PCG_RHO;
}

// Called by ALL threads of one CTA (any multiple of 32 threads up to 1024).  sm: 32 doubles of
// shared memory.  Level 2: one warp per group, each lane a fixed run of consecutive chunk partials;
// level 3: warp sums of 32 consecutive group sums, then one warp sum over those 32 values.
template <bool WITH_P2P>
static __device__ __forceinline__ void finalize_levels(const FinArgs &F, double *sm) {
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (F.stage != 2) {
    const int per = (F.cpg + 31) / 32;
    for (int64_t g = w; g < F.n_groups_local; g += nw) {
      const int64_t base = g * F.cpg;
      double s = 0.0;
      for (int t = 0; t < per; t++) {
        const int k = l * per + t;
        if (k < F.cpg && base + k < F.n_chunks) s += __ldcg(F.partials + base + k);
      }
      s = warp_sum(s);
      if (l == 0) F.gsend[F.group_begin + g] = s;
    }
    if (F.stage == 1) return;
    __syncthreads();
  }
  const double *gs = F.stage == 2 ? F.grecv : F.gsend;
  bool comm_failed = false;
  if (WITH_P2P && F.stage == 3) {
    // all-gather of the group sums over NVLink: store mine into every rank's slot (own included),
    // publish an epoch flag to every rank, wait for everybody's flag.  Ranks are never more
    // than one reduction apart, so two slots are enough; the epoch counts the reductions that were
    // EXECUTED (device-side counter: launches skipped after convergence do not advance it, so two
    // consecutive executed reductions never share a slot).  Peer r's flag also tells me that all
    // NVLink stores r issued before it (its halo push) have landed.
    const P2PView &q = F.p2p;
    const unsigned long long epoch = *q.epoch_ctr + 1ull;  // read by all threads before thread 0 bumps it below
    const int slot = (int)(epoch & 1ull);
    const int nloc = (int)F.n_groups_local;
    for (int i = threadIdx.x; i < nloc * q.P; i += blockDim.x) {
      const int r = i / nloc, g = (int)F.group_begin + i % nloc;
      q.red[r][slot * MAX_GROUPS + g] = F.gsend[g];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < q.P) {
      *((volatile unsigned long long *)&q.flags[threadIdx.x][q.me]) = epoch;
      const volatile unsigned long long *mine = (const volatile unsigned long long *)&q.flags[q.me][threadIdx.x];
      const long long t0 = clock64();
      while (*mine < epoch) {
        if (clock64() - t0 > q.timeout) {  // a peer is gone; fail instead of hanging
          *q.err = 1;
          break;
        }
      }
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0) *q.epoch_ctr = epoch;
    comm_failed = *((volatile int *)q.err) != 0;
    gs = q.red[q.me] + slot * MAX_GROUPS;
  }
  for (int seg = w; seg < 32; seg += nw) {
    const int i = 32 * seg + l;
    double v = i < F.n_groups_global ? __ldcg(gs + i) : 0.0;
    v = warp_sum(v);
    if (l == 0) sm[seg] = v;
  }
  __syncthreads();
  if (w == 0) {
    double t = sm[l];
    t = warp_sum(t);
    if (l == 0) {
      if (comm_failed) {
        // stale group sums: do not advance the recurrences; stop the solver / flag the scalar
        if (F.what == FIN_DOT) {
          F.out[0] = t;
          F.out[1] = 1.0;
        } else {
          F.st->done = 1;
          F.st->converged = 0;
        }
      } else {
        fin_scalars(F, t);
      }
    }
  }
}

// Tail of a producing kernel: thread 0 has just stored this CTA's chunk partial.  The CTA that
// draws the last ticket finishes the reduction (all other CTAs' partials are visible: each fenced
// before taking its ticket).  sm: >= 32 doubles of shared memory.  No-op when F.counter == NULL.
static __device__ __forceinline__ void last_cta_finalize(const FinArgs &F, double *sm) {
  if (!F.counter) return;
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(F.counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    finalize_levels<false>(F, sm);  // no peer-memory stage here: keeps the kernel free of a stack frame
    if (threadIdx.x == 0) *F.counter = 0u;
  }
}

}  // namespace nosh
