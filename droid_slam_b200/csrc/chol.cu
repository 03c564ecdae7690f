// Damped SPD solve of the reduced pose system on the device, fp64:  (H + diag(ep + lm*diag(H))) x = b.
//
// Replaces the reference's host-side SparseBlock::solve (src/droid_kernels.cu:1201-1222: Eigen::SimplicialLLT in
// fp64 on the CPU behind two PCIe round trips).  Same contract: fp64 arithmetic, a non-positive pivot means
// "not SPD" and yields x = 0.
//
// One thread-block CLUSTER (up to 16 CTAs, i.e. up to 16 SMs of one GPC) runs a right-looking tiled Cholesky with
// 32x32 fp64 tiles that live in global memory (L2 resident: 6P x 6P doubles is 1.5 MB for P = 71).  Per panel k:
//     TRSM of the column-k tiles (one warp per tile, lane = row, forward substitution against L_kk in shared memory)
//       -- cluster barrier --
//     trailing update A_ij -= L_ik L_jk^T: one warp per tile with an 8x4 register block per lane (operands staged in the
//     warp's padded shared-memory slabs, coalesced global I/O).  The NEXT diagonal tile is on the critical path, so CTA 0
//     updates it with all 256 threads and its warp 0 factors it immediately (rows in registers, the pivot column is
//     broadcast through shared memory) while every other warp of the cluster works on the remaining tiles; a spare
//     warp inverts L_kk for the backward pass
//       -- cluster barrier --
// i.e. two hardware cluster barriers per panel instead of kernel launches or grid-wide syncs.
// ENVELOPE: the reduced pose system of a sliding-window / proximity factor graph is block banded (pose a couples to pose b only
// through a common source frame), and a Cholesky factor never fills in left of a row's first nonzero.  The load phase records, per
// 32-row tile row, the first structurally nonzero tile column (`first`); TRSM, trailing updates and the backward substitution then
// skip every tile outside that envelope.  A dense system (the 72-keyframe metric window) does the same work as before; the global-BA
// configs (6P = 2394 ... 5994, half bandwidth ~150) drop from O(n^3) to O(n b^2) -- what Eigen's sparse LLT does for the reference.
// The right-hand side rides along as an extra tile row, so L^-1 b comes out of the factorisation for free; the
// backward substitution uses the inverted diagonal tiles and runs in CTA 0.
// (B200 note, measured: a dependent fp64 op costs ~10-20 cycles and a 64-bit warp shuffle pair is slower than a
//  shared-memory broadcast, which is why the pivot column goes through shared memory and the pivot uses an fp32
//  rsqrt seed + one Newton step -- 3e-14 relative, far below the fp32 rounding of the result.)
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace dba {

constexpr int kT = 32;                 // tile edge
constexpr int kTP = kT + 1;            // padded row length in shared memory
constexpr int kCholThreads = 256;      // 8 warps per CTA
constexpr int kCholWarps = kCholThreads / 32;

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CHOL_STAMP(slot) do { if (p.timing && cta == 0 && tid == 0) p.timing[(slot)] = gtimer(); } while (0)
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void stcg(double* p, double v) { __stcg(p, v); }

// 1/sqrt(d): MUFU.RSQ64H seed (~2^-22) + one third-order correction r += r t (1/2 + 3/8 t), t = 1 - d r^2 (error ~ t^3: full fp64).
// Deliberately branch-free: a branch here splits warp_potrf into basic blocks and stops ptxas from scheduling the rank-1 update
// under the latency of this chain.  d <= 0 yields NaN/inf, which the caller flags through its pivot test.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double t = fma(-d, r * r, 1.0);
  return fma(fma(0.375, t, 0.5), r * t, r);
}

// Cholesky of a 32x32 tile, one row per lane in registers; the pivot column is broadcast through `col` (2 x 32 doubles
// of shared memory private to the warp).  rdiag_out receives 1/L[k][k] (lane k's value).  Returns false on a
// non-positive pivot.
__device__ __forceinline__ bool warp_potrf(double (&a)[kT], int lane, double* col, double& rdiag_out) {
  bool ok = true;
  rdiag_out = 0.0;
  // software-pipelined: the pivot of column k+1 only needs a[k+1] after the rank-1 update of column k, so that element is updated
  // first and its shuffle + rsqrt chain (the latency that bounds this routine) runs under the remaining 30 updates of column k
  double d = __shfl_sync(0xffffffffu, a[0], 0);
  double r = fast_rsqrt(d);
#pragma unroll
  for (int k = 0; k < kT; k++) {
    if (!(d > 0.0)) ok = false;
    const double l = (lane == k) ? d * r : a[k] * r;
    if (lane == k) rdiag_out = r;
    a[k] = l;
    double* cb = col + (k & 1) * kT;
    cb[lane] = l;
    __syncwarp();
    if (k + 1 < kT) {
      a[k + 1] -= l * cb[k + 1];
      d = __shfl_sync(0xffffffffu, a[k + 1], k + 1);
      r = fast_rsqrt(d);
    }
#pragma unroll
    for (int j = k + 2; j < kT; j++) a[j] -= l * cb[j];   // only rows >= j are meaningful
  }
  return ok;
}

struct CholParams {
  const double* H;   // [n][n] fp64, lower triangle valid
  const double* b;   // [n]
  double* L;         // [(nt+1)*32][nt*32] row-major working matrix (tile row nt carries b^T in its row 0)
  double* Linv;      // [nt][32][32] inverses of the diagonal tiles
  double* rdiag;     // [nt*32] reciprocals of diag(L)
  int* first;        // [nt+1] envelope: first nonzero tile column of each tile row (rhs row nt: 0)
  int* fail;         // sticky flag: non-positive pivot
  float* x;          // [n] result (fp32 like the reference's dx)
  int n, nt;
  double lm, ep;
  unsigned long long* timing;   // debug (DBA_CHOL_TIMING=1): globaltimer stamps of CTA 0 / the potrf warp, else nullptr
  CholPeers peers;              // world <= 1: plain local system
};

// one warp: C (32x32 at Ct) -= A (at At) * B^T (at Bt); lane (rg = lane>>3, cg = lane&7) owns rows 8rg..8rg+7, cols 4cg..4cg+3
__device__ __forceinline__ void warp_tile_update(const double* At, const double* Bt, double* Ct, int ld, int lane,
                                                 double (*sA)[kTP], double (*sB)[kTP]) {
  const int rg = lane >> 3, cgp = lane & 7;
  double acc[8][4];
  // the C tile's loads go out first and return under the operand staging (one L2 round trip instead of two)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double2 c01 = __ldcg(reinterpret_cast<const double2*>(Ct + (size_t)(8 * rg + i) * ld + 4 * cgp));
    const double2 c23 = __ldcg(reinterpret_cast<const double2*>(Ct + (size_t)(8 * rg + i) * ld + 4 * cgp + 2));
    acc[i][0] = c01.x; acc[i][1] = c01.y; acc[i][2] = c23.x; acc[i][3] = c23.y;
  }
#pragma unroll 16
  for (int r = 0; r < kT; r++) { sA[r][lane] = ldcg(At + (size_t)r * ld + lane); sB[r][lane] = ldcg(Bt + (size_t)r * ld + lane); }
  __syncwarp();
#pragma unroll 4
  for (int q = 0; q < kT; q++) {
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = sA[8 * rg + i][q];
#pragma unroll
    for (int jx = 0; jx < 4; jx++) bv[jx] = sB[4 * cgp + jx][q];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int jx = 0; jx < 4; jx++) acc[i][jx] -= av[i] * bv[jx];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    __stcg(reinterpret_cast<double2*>(Ct + (size_t)(8 * rg + i) * ld + 4 * cgp), make_double2(acc[i][0], acc[i][1]));
    __stcg(reinterpret_cast<double2*>(Ct + (size_t)(8 * rg + i) * ld + 4 * cgp + 2), make_double2(acc[i][2], acc[i][3]));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kCholThreads, 1) chol_cluster_kernel(CholParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks();
  const int cta = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gw = cta * kCholWarps + warp;          // warp id within the cluster
  const int nwarps = ncta * kCholWarps;
  const int nt = p.nt, n = p.n;
  const int ld = nt * kT;                          // leading dimension of L
  double* __restrict__ L = p.L;

  __shared__ double s_Lkk[kT][kTP];
  __shared__ double s_rdiag[kT];
  __shared__ double s_vec[kT];
  __shared__ double s_col[2 * kT];
  __shared__ int s_act[kCholThreads];                // tile rows with a nonzero tile in panel k (ascending; the rhs row nt is always last)
  __shared__ int s_wc[kCholWarps];
  __shared__ int s_nact;
  extern __shared__ double s_dyn[];                  // per-warp slabs + two CTA-wide tiles, rows padded to 33 doubles
  double (*s_A)[kT][kTP] = reinterpret_cast<double (*)[kT][kTP]>(s_dyn);
  double (*s_B)[kT][kTP] = reinterpret_cast<double (*)[kT][kTP]>(s_dyn + (size_t)kCholWarps * kT * kTP);
  double (*s_D)[kTP] = reinterpret_cast<double (*)[kTP]>(s_dyn + (size_t)2 * kCholWarps * kT * kTP);
  double (*s_T)[kTP] = reinterpret_cast<double (*)[kTP]>(s_dyn + (size_t)2 * kCholWarps * kT * kTP + kT * kTP);

  // ---- fused peer-to-peer reduction: wait until every rank has published its partial system for this epoch ------------
  const int world = p.peers.world;
  if (world > 1) {
    __shared__ int s_timeout;
    if (tid == 0) {
      int bad = 0;
      const unsigned long long want = p.peers.epoch_dev ? *p.peers.epoch_dev : p.peers.epoch;
      for (int r = 0; r < world; r++) {
        unsigned long long v = 0;
        long long spins = 0;
        do {
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.peers.flags + r) : "memory");
        } while (v < want && ++spins < (1ll << 24));
        if (v < want) bad = 1;
      }
      s_timeout = bad;
    }
    __syncthreads();
    if (cta == 0 && tid == 0) *p.fail = s_timeout ? 2 : 0;   // a peer never arrived: give up loudly (dx = 0), never hang
  } else if (cta == 0 && tid == 0) *p.fail = 0;
  // ---- envelope: first[i] starts at the diagonal, the load below lowers it to the first nonzero tile of the row
  const bool envelope = nt < kCholThreads;             // one thread per tile row in the per-panel scan below
  for (int i = cta * kCholThreads + tid; i <= nt; i += ncta * kCholThreads) p.first[i] = (i < nt && envelope) ? i : 0;
  cluster.sync();
  // ---- load: lower tiles of H with damping (reference :1205-1206), identity padding, rhs row ------------------
  {
    const size_t total = (size_t)(nt + 1) * kT * ld;
    const size_t nn = (size_t)n * n;
    const size_t stride = (size_t)ncta * kCholThreads;
    constexpr int kU = 4;                               // elements per thread in flight: with peers, kU x world NVLink loads overlap their ~2 us round trips
    for (size_t base = (size_t)cta * kCholThreads + tid; base < total; base += kU * stride) {
      size_t srcs[kU];
      double vals[kU];
      double t[kU][8];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const size_t idx = base + u * stride;
        srcs[u] = (size_t)-1; vals[u] = 0.0;
        if (idx < total) {
          const int r = (int)(idx / ld), c = (int)(idx - (size_t)r * ld);
          if (r < nt * kT) {
            if (r < n && c < n) {
              if (c <= r) srcs[u] = (size_t)r * n + c;
              else if ((r >> 5) == (c >> 5)) srcs[u] = (size_t)c * n + r;     // diagonal tiles are kept fully symmetric
            } else if (r == c) vals[u] = 1.0;
          } else if (r == nt * kT && c < n) srcs[u] = nn + c;
        }
        if (srcs[u] != (size_t)-1) {                      // element of the [n*n | n] system feeding this entry
          if (world > 1) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
              t[u][q] = 0.0;
              if (q < world) asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(t[u][q]) : "l"(p.peers.sys[q] + srcs[u]) : "memory");
            }
          } else t[u][0] = (srcs[u] < nn) ? p.H[srcs[u]] : p.b[srcs[u] - nn];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const size_t idx = base + u * stride;
        if (idx >= total) continue;
        double v = vals[u];
        if (srcs[u] != (size_t)-1) {
          // all peer loads were issued before the first add, then summed in fixed rank order: every rank computes the identical sum
          if (world > 1) {
#pragma unroll
            for (int q = 0; q < 8; q++)
              if (q < world) v += t[u][q];
          } else v = t[u][0];
          const int r = (int)(idx / ld), c = (int)(idx - (size_t)r * ld);
          if (srcs[u] < nn && r == c) v += p.ep + p.lm * v;
          if (envelope && v != 0.0 && r < nt * kT) {
            const int tr = r >> 5, tc = c >> 5;
            if (tc < tr && tc < *reinterpret_cast<volatile int*>(p.first + tr)) atomicMin(p.first + tr, tc);
          }
        }
        stcg(L + idx, v);
      }
    }
  }
  CHOL_STAMP(0);
  cluster.sync();
  CHOL_STAMP(1);

  // ---- potrf of tile (0,0) --------------------------------------------------------------------------------------
  if (gw == 0) {
    double a[kT], rd;
#pragma unroll
    for (int c = 0; c < kT; c++) a[c] = ldcg(L + (size_t)lane * ld + c);
    if (!warp_potrf(a, lane, s_col, rd) && lane == 0) *p.fail = 1;
#pragma unroll
    for (int c = 0; c < kT; c++) stcg(L + (size_t)lane * ld + c, (c <= lane) ? a[c] : 0.0);
    stcg(p.rdiag + lane, rd);
  }
  cluster.sync();
  CHOL_STAMP(2);

  for (int k = 0; k < nt; k++) {
    // ---- every CTA: L_kk and its reciprocal diagonal into shared memory
    for (int e = tid; e < kT * kT; e += kCholThreads) {
      const int r = e >> 5, c = e & 31;
      s_Lkk[r][c] = ldcg(L + (size_t)(k * kT + r) * ld + k * kT + c);
    }
    if (tid < kT) s_rdiag[tid] = ldcg(p.rdiag + k * kT + tid);
    // active tile rows of panel k (inside the envelope), in ascending order -- every CTA builds the identical list
    {
      const int i_row = k + 1 + tid;
      const bool act = envelope ? (i_row <= nt && (i_row == nt || __ldcg(p.first + i_row) <= k)) : false;
      const unsigned bal = __ballot_sync(0xffffffffu, act);
      if (lane == 0) s_wc[warp] = __popc(bal);
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; w++) base += s_wc[w];
      if (act) s_act[base + __popc(bal & ((1u << lane) - 1u))] = i_row;
      if (tid == 0) { int tot = 0; for (int w = 0; w < kCholWarps; w++) tot += s_wc[w]; s_nact = tot; }
    }
    __syncthreads();
    const int nact = envelope ? s_nact : (nt - k);       // >= 1: the rhs row
    CHOL_STAMP(8 + 8 * k + 0);
    // ---- TRSM: tiles (i,k) of the active rows (tile row nt is the right-hand side)
    for (int ta = gw; ta < nact; ta += nwarps) {
      const int i = envelope ? s_act[ta] : k + 1 + ta;
      double a[kT];
      double* tile = L + (size_t)(i * kT) * ld + k * kT;
#pragma unroll
      for (int r = 0; r < kT; r++) s_A[warp][r][lane] = ldcg(tile + (size_t)r * ld + lane);     // coalesced rows, all 32 loads in flight
      __syncwarp();
#pragma unroll
      for (int c = 0; c < kT; c++) a[c] = s_A[warp][lane][c];                                   // lane = row
#pragma unroll
      for (int c = 0; c < kT; c++) {
        const double xv = a[c] * s_rdiag[c];
        a[c] = xv;
#pragma unroll
        for (int j = c + 1; j < kT; j++) a[j] -= xv * s_Lkk[j][c];
        asm volatile("" ::: "memory");
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < kT; c++) s_A[warp][lane][c] = a[c];
      __syncwarp();
#pragma unroll 8
      for (int r = 0; r < kT; r++) stcg(tile + (size_t)r * ld + lane, s_A[warp][r][lane]);
      __syncwarp();
    }
    CHOL_STAMP(8 + 8 * k + 1);
    cluster.sync();
    CHOL_STAMP(8 + 8 * k + 2);
    // ---- trailing update with panel k
    const int rem = nt - k - 1;                       // remaining tile columns
    const int m1 = nact - 1;                          // active rows without the rhs row
    const int ntri = m1 * (m1 + 1) / 2;
    const int ntasks = ntri + m1;                     // tiles (i,j) of active rows, j <= i < nt, plus the rhs row tiles (nt,j)
    // task 0 = tile (k+1,k+1) when row k+1 is active: CTA 0 updates + factors it below; otherwise that tile needs no update (CTA 0 still
    // factors it) and task 0 is an ordinary tile of the workers
    const bool diag_active = envelope ? (m1 >= 1 && s_act[0] == k + 1) : (rem >= 1);
    if (cta == 0 && rem >= 1) {
      // next diagonal tile (task 0): all 256 threads update it, warp 0 factors it
      const double* At = L + (size_t)((k + 1) * kT) * ld + k * kT;
      double* Ct = L + (size_t)((k + 1) * kT) * ld + (k + 1) * kT;
      for (int e = tid; e < kT * kT; e += kCholThreads) {
        const int r = e >> 5, c = e & 31;
        s_D[r][c] = ldcg(At + (size_t)r * ld + c);
        s_T[r][c] = ldcg(Ct + (size_t)r * ld + c);
      }
      __syncthreads();
      {
        const int r = tid >> 3, c0 = (tid & 7) * 4;
        double acc[4] = {s_T[r][c0], s_T[r][c0 + 1], s_T[r][c0 + 2], s_T[r][c0 + 3]};
#pragma unroll 8
        for (int q = 0; q < kT; q++) {
          const double ar = s_D[r][q];
#pragma unroll
          for (int jx = 0; jx < 4; jx++) acc[jx] -= ar * s_D[c0 + jx][q];
        }
        __syncthreads();
#pragma unroll
        for (int jx = 0; jx < 4; jx++) s_T[r][c0 + jx] = acc[jx];
      }
      __syncthreads();
      if (warp == 0) {
        if (p.timing && lane == 0) p.timing[8 + 8 * k + 5] = gtimer();
        double a[kT], rd;
#pragma unroll
        for (int c = 0; c < kT; c++) a[c] = s_T[lane][c];
        if (!warp_potrf(a, lane, s_col, rd) && lane == 0) *p.fail = 1;
        if (p.timing && lane == 0) p.timing[8 + 8 * k + 6] = gtimer();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < kT; c++) s_T[lane][c] = (c <= lane) ? a[c] : 0.0;
        __syncwarp();
#pragma unroll 8
        for (int r = 0; r < kT; r++) stcg(Ct + (size_t)r * ld + lane, s_T[r][lane]);
        stcg(p.rdiag + (k + 1) * kT + lane, rd);
      }
    }
    // remaining tiles: every warp of the cluster except the factoring one
    if (gw != 0) {
      for (int t = (diag_active ? 1 : 0) + gw - 1; t < ntasks; t += nwarps - 1) {
        int i, j;
        if (t < ntri) {
          int bi = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
          while (bi * (bi + 1) / 2 > t) bi--;
          while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
          const int bj = t - bi * (bi + 1) / 2;
          i = envelope ? s_act[bi] : k + 1 + bi; j = envelope ? s_act[bj] : k + 1 + bj;
        } else { i = nt; j = envelope ? s_act[t - ntri] : k + 1 + (t - ntri); }
        warp_tile_update(L + (size_t)(i * kT) * ld + k * kT, L + (size_t)(j * kT) * ld + k * kT, L + (size_t)(i * kT) * ld + j * kT, ld, lane,
                         s_A[warp], s_B[warp]);
      }
    }
    CHOL_STAMP(8 + 8 * k + 3);
    // inverse of L_kk (for the backward substitution) by the last warp of the cluster: lane j owns column j
    if (gw == nwarps - 1) {
      double xcol[kT];
#pragma unroll
      for (int i = 0; i < kT; i++) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < i; m++) s += (m >= lane) ? s_Lkk[i][m] * xcol[m] : 0.0;
        xcol[i] = (i == lane) ? s_rdiag[i] : ((i > lane) ? -s * s_rdiag[i] : 0.0);
      }
#pragma unroll
      for (int i = 0; i < kT; i++) stcg(p.Linv + ((size_t)k * kT + i) * kT + lane, xcol[i]);
    }
    cluster.sync();
    CHOL_STAMP(8 + 8 * k + 4);
  }

  if (cta != 0) return;
  // ---- backward substitution  L^T x = y  in CTA 0;  y^T = row 0 of tile row nt -------------------------------------
  double* y = L + (size_t)(nt * kT) * ld;            // [ld], overwritten by x
  for (int k = nt - 1; k >= 0; k--) {
    if (warp == 0) {
      // x_k = Linv_kk^T y_k : lane c computes sum_r Linv[r][c] * y[r]
      const double yk = ldcg(y + k * kT + lane);
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < kT; r++) s += ldcg(p.Linv + ((size_t)k * kT + r) * kT + lane) * __shfl_sync(0xffffffffu, yk, r);
      stcg(y + k * kT + lane, s);
      s_vec[lane] = s;
    }
    __syncthreads();
    // y_i -= L_ki^T x_k  for first[k] <= i < k (tiles left of the envelope are zero): lane = column of tile (k,i)
    for (int i = __ldcg(p.first + k) + warp; i < k; i += kCholWarps) {
      double s = 0.0;
#pragma unroll 8
      for (int r = 0; r < kT; r++) s += ldcg(L + (size_t)(k * kT + r) * ld + i * kT + lane) * s_vec[r];
      stcg(y + i * kT + lane, ldcg(y + i * kT + lane) - s);
    }
    __syncthreads();
  }
  CHOL_STAMP(3);
  const bool failed = (*reinterpret_cast<volatile int*>(p.fail)) != 0;
  for (int i = tid; i < n; i += kCholThreads) {
    const double v = ldcg(y + i);
    p.x[i] = (failed || !isfinite(v)) ? 0.f : (float)v;      // reference: solver.info() != Success -> zeros
  }
}

size_t chol_workspace_bytes(int n) {
  const size_t nt = (size_t)(n + kT - 1) / kT;
  const size_t ld = nt * kT;
  return ((nt + 1) * kT * ld + nt * kT * kT + nt * kT) * sizeof(double) + (nt + 2) * sizeof(int) + 256;
}

// H [n][n] fp64, b [n] fp64 -> x [n] fp32; fail flag is a device int
int chol_solve_launch(const double* H, const double* b, int n, double lm, double ep, void* workspace, int* fail, float* x, cudaStream_t st,
                      const CholPeers* peers) {
  if (n <= 0) return DBA_OK;
  CholParams p;
  p.H = H; p.b = b; p.n = n; p.nt = (n + kT - 1) / kT; p.lm = lm; p.ep = ep; p.fail = fail; p.x = x;
  if (peers) p.peers = *peers; else { p.peers.world = 0; p.peers.flags = nullptr; p.peers.epoch = 0; p.peers.epoch_dev = nullptr; for (int k = 0; k < 8; k++) p.peers.sys[k] = nullptr; }
  const size_t ld = (size_t)p.nt * kT;
  p.L = reinterpret_cast<double*>(workspace);
  p.Linv = p.L + (size_t)(p.nt + 1) * kT * ld;
  p.rdiag = p.Linv + (size_t)p.nt * kT * kT;
  p.first = reinterpret_cast<int*>(p.rdiag + (size_t)p.nt * kT);

  const size_t dyn_smem = ((size_t)2 * kCholWarps + 2) * kT * kTP * sizeof(double);
  static int cluster_size = 0;
  if (cluster_size == 0) {
    cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    int best = 8;
    for (int cs = 16; cs >= 8; cs -= 8) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = dyn_smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, chol_cluster_kernel, &cfg) == cudaSuccess && nclusters >= 1) { best = cs; break; }
    }
    cudaGetLastError();
    cluster_size = best;
  }
  // small systems do not need the whole cluster
  int cs = cluster_size;
  const int tiles_first_panel = p.nt * (p.nt + 1) / 2 + 1;
  while (cs > 1 && (cs / 2) * kCholWarps - 1 >= tiles_first_panel) cs /= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = dyn_smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  static unsigned long long* tbuf = nullptr;
  static const bool want_timing = getenv("DBA_CHOL_TIMING") != nullptr;
  p.timing = nullptr;
  if (want_timing) {
    if (!tbuf) cudaMallocHost(&tbuf, 4096 * sizeof(unsigned long long));
    memset(tbuf, 0, 4096 * sizeof(unsigned long long));
    if (8 + 8 * p.nt < 4096) p.timing = tbuf;
  }
  DBA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, chol_cluster_kernel, p), "chol_cluster_kernel launch");
  if (p.timing) {
    cudaStreamSynchronize(st);
    const unsigned long long t0 = tbuf[0];
    fprintf(stderr, "[chol timing] n=%d nt=%d cluster=%d  load %.1f us, potrf0 %.1f us, total-to-backsub-end %.1f us\n", n, p.nt, cs, (tbuf[1] - t0) / 1e3,
            (tbuf[2] - tbuf[1]) / 1e3, (tbuf[3] - t0) / 1e3);
    for (int k = 0; k < p.nt; k++) {
      const unsigned long long* q = tbuf + 8 + 8 * k;
      fprintf(stderr, "  panel %2d: Lkk-load@%.1f trsm(cta0) %.1f  barrier %.1f  update(cta0 thread0) %.1f  inv+barrier %.1f | diag tile: coop-update %.1f potrf %.1f\n", k,
              (q[0] - t0) / 1e3, (q[1] - q[0]) / 1e3, (q[2] - q[1]) / 1e3, (q[3] - q[2]) / 1e3, (q[4] - q[3]) / 1e3,
              q[5] ? (q[5] - q[2]) / 1e3 : 0.0, q[6] ? (q[6] - q[5]) / 1e3 : 0.0);
    }
  }
  return DBA_OK;
}

}  // namespace dba

// standalone entry (used by the solver tests and by callers that already hold a reduced system)
extern "C" size_t dba_solve_workspace_bytes(int n) { return dba::chol_workspace_bytes(n) + 64; }

extern "C" int dba_solve_spd(const double* H, const double* b, int n, float lm, float ep, float* x, int* fail_flag_device,
                             void* workspace, size_t workspace_bytes, dba_stream_t stream) {
  DBA_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return DBA_OK;
  DBA_CHECK_ARG(H && b && x && fail_flag_device && workspace, "null pointer");
  if (workspace_bytes < dba::chol_workspace_bytes(n)) { dba::set_error("solve workspace too small"); return DBA_ERR_WORKSPACE; }
  return dba::chol_solve_launch(H, b, n, (double)lm, (double)ep, workspace, fail_flag_device, x, (cudaStream_t)stream);
}
