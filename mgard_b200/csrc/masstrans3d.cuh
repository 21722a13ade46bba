// Fused 3-D load-vector kernel (sm_100a): mass matrix x restriction along f, c
// and r of the level-l coefficient function, read straight from the
// coefficient array in its coarse-first layout.
//
// Replaces, for D == 3, the reference's Lpk1Reo3D -> Lpk2Reo3D -> Lpk3Reo3D
// chain (Correction/LinearProcessingKernel3D.hpp:27-1090, mass_trans:
// Correction/LPKFunctor.h:47-66) with the same arithmetic in the same order, so
// results stay bit-identical, in one pass: n_l elements read, n_l/8 written.
//
// In the coarse-first layout the even (E) and odd (O) padded positions of a
// line are two contiguous vectors, and the five inputs of mass_trans at coarse
// index i are E[i-1], O[i-1], E[i], O[i], E[i+1].  So
//   f pass: a warp takes one row; every lane loads E[kf] and O[kf] (two
//           coalesced loads) and gets its neighbours' values with shuffles (the
//           warp spans 32 columns and owns the inner 30);
//   c pass: from the f-pass rows of the tile in shared memory;
//   r pass: from a five-deep register ring while the block sweeps r.
// A thread block owns a TC x TF tile of coarse (c, f) columns and a segment of
// coarse r indices; the loads of the next plane are issued before the current
// one is worked on.
#pragma once

namespace masstrans3d {

typedef long long i64;

// A warp covers 32 consecutive coarse f columns kf0-1 .. kf0+30 and OWNS the
// middle TF = 30: the two outer lanes only supply their neighbours' inputs, so
// no separate halo loads are needed.
constexpr int TC = 8, TF = 30, NT = 256, NW = NT / 32;
constexpr int NROW = 2 * TC + 3;                  // E rows kc0-1..kc0+TC, O rows kc0-1..kc0+TC-1
constexpr int RPW = (NROW + NW - 1) / NW;         // rows per warp and plane
constexpr int NST = 4;                            // raw planes in flight (power of two)

template <typename T> struct Params {
  int n[3], nc[3];  // fine / coarse level shape (r, c, f)
  i64 sin[3];       // coefficient array strides
  i64 sw[3];        // dense load-vector strides
  const T *mt[3];   // 9 x nc mass_trans tables
  int rsegs, ctiles, ftiles;
};

template <typename T>
__device__ __forceinline__ T mass_trans_k(T a, T b, T c, T d, T e, const T (&k)[9]) {
  T tb = a * k[0] + b * k[1] + c * k[2];
  T tc = b * k[2] + c * k[3] + d * k[4];
  T td = c * k[4] + d * k[5] + e * k[6];
  tc += tb * k[7] + td * k[8];
  return tc;
}

// position of the m-th even (odd = false) / odd (odd = true) padded node of a
// dimension in the coarse-first layout; -1: no such node (its value is zero)
__device__ __forceinline__ int pos(int m, bool odd, int n, int nc) {
  if (m < 0)
    return -1;
  const int p = odd ? nc + m : m;
  return (odd ? p < n : m < nc) ? p : -1;
}

template <typename T>
__global__ void __launch_bounds__(NT, 3)
masstrans3d_kernel(const Params<T> P, const T *__restrict__ in, T *__restrict__ w_out) {
  __shared__ T s_a1[2][NROW][32];
  __shared__ T s_raw[NST][NROW][2][32]; // [stage][row][E / O][lane] // f-pass rows of the current plane (double buffered)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int bid = blockIdx.x;
  const int ft = bid % P.ftiles;
  bid /= P.ftiles;
  const int ct = bid % P.ctiles;
  const int rs = bid / P.ctiles;
  const int kc0 = ct * TC, kf0 = ft * TF;
  const int nr = P.n[0], ncn = P.n[1], nf = P.n[2];
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  const int per = (rr + P.rsegs - 1) / P.rsegs;
  const int rk0 = rs * per, rk1 = min(rr, rk0 + per);
  if (rk0 >= rk1)
    return;

  // ---- per-thread constants -------------------------------------------------
  // f: this lane's column, offsets of its E and O element (-1: none -> zero)
  const int kf = kf0 - 1 + lane;
  const bool own_f = lane >= 1 && lane <= TF && kf < ff;
  const int pe = pos(kf, false, nf, ff), po = pos(kf, true, nf, ff);
  // rows of this warp: per row the offsets of the lane's E and O element from the
  // plane base (32-bit; -1: nothing to load)
  // (offsets of missing elements are 0 with a copy size of 0 = zero fill)
  int roff_e[RPW], roff_o[RPW];
  int se_even[RPW], se_odd[RPW], so_any[RPW]; // copy sizes on even-r / odd-r planes
#pragma unroll
  for (int q = 0; q < RPW; q++) {
    const int row = wid + q * NW;
    roff_e[q] = roff_o[q] = 0;
    se_even[q] = se_odd[q] = so_any[q] = 0;
    if (row < NROW) {
      const bool odd = row >= TC + 2;
      const int m = kc0 - 1 + (odd ? row - (TC + 2) : row);
      const int pc = pos(m, odd, ncn, cc);
      if (pc >= 0) {
        if (pe >= 0) {
          roff_e[q] = (int)((i64)pc * P.sin[1] + (i64)pe * P.sin[2]);
          se_odd[q] = (int)sizeof(T);
          // the all-coarse block (even r, even c, even f) counts as zero
          se_even[q] = odd ? (int)sizeof(T) : 0;
        }
        if (po >= 0) {
          roff_o[q] = (int)((i64)pc * P.sin[1] + (i64)po * P.sin[2]);
          so_any[q] = (int)sizeof(T);
        }
      }
    }
  }
  T kfc[9], kcc[9];
#pragma unroll
  for (int m = 0; m < 9; m++) {
    kfc[m] = (kf >= 0 && kf < ff) ? P.mt[2][m * ff + kf] : (T)0;
    kcc[m] = (kc0 + wid < cc) ? P.mt[1][m * cc + kc0 + wid] : (T)0;
  }
  const bool col_ok = (kc0 + wid < cc) && own_f;
  const i64 w_col = (i64)(kc0 + wid) * P.sw[1] + (i64)kf * P.sw[2];

  // ---- plane sequence: (E,O) of k = rk0-1 .. rk1-1, then E of rk1 ------------
  // plane index t: k = rk0 - 1 + t / 2, odd-r plane iff t & 1
  const int nplanes = 2 * (rk1 - rk0 + 1) + 1;
  // raw planes: ring of NST stages in shared memory filled with asynchronous
  // copies NST-1 planes ahead (a warp only ever reads the rows it requested
  // itself, so completion needs no block-wide barrier)
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(&s_raw[0][0][0][0]);
  auto issue = [&](int t) {
    const int k = rk0 - 1 + (t >> 1);
    const bool rodd = t & 1;
    const int pr = pos(k, rodd, nr, rr);
    const T *base = in + (i64)(pr >= 0 ? pr : 0) * P.sin[0];
    const unsigned stage = raw_addr + (unsigned)((t & (NST - 1)) * (NROW * 64) * (int)sizeof(T));
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      if (row < NROW) { // warp uniform
        const unsigned d = stage + (unsigned)((row * 64 + lane) * (int)sizeof(T));
        const T *ge = base + roff_e[q], *go = base + roff_o[q];
        const int se = pr >= 0 ? (rodd ? se_odd[q] : se_even[q]) : 0; // 0: zero fill
        const int so = pr >= 0 ? so_any[q] : 0;
        if (sizeof(T) == 4) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(ge), "r"(se));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d + 32 * 4), "l"(go), "r"(so));
        } else {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(ge), "r"(se));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d + 32 * 8), "l"(go), "r"(so));
        }
      }
    }
  };
  T ring[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};
#pragma unroll
  for (int t = 0; t < NST - 1; t++) {
    if (t < nplanes)
      issue(t);
    asm volatile("cp.async.commit_group;\n" ::);
  }
  const int lm = lane > 0 ? lane - 1 : 0, lp = lane < 31 ? lane + 1 : 31;
  for (int t = 0; t < nplanes; t++) {
    if (t + NST - 1 < nplanes)
      issue(t + NST - 1);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NST - 1));
    __syncwarp();
    // f pass of plane t
    T a1[RPW];
    const T(*raw)[2][32] = s_raw[t & (NST - 1)];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      a1[q] = (T)0;
      if (row < NROW) // warp uniform; lanes 0 and 31 produce values nobody uses
        a1[q] = mass_trans_k<T>(raw[row][0][lm], raw[row][1][lm], raw[row][0][lane],
                                raw[row][1][lane], raw[row][0][lp], kfc);
    }
    T(*sa)[32] = s_a1[t & 1];
#pragma unroll
    for (int q = 0; q < RPW; q++) {
      const int row = wid + q * NW;
      if (row < NROW)
        sa[row][lane] = a1[q];
    }
    __syncthreads();
    // c pass: coarse row kc0 + wid from E rows wid, wid+1, wid+2 and O rows wid, wid+1
    const T a2 = mass_trans_k<T>(sa[wid][lane], sa[TC + 2 + wid][lane], sa[wid + 1][lane],
                                 sa[TC + 2 + wid + 1][lane], sa[wid + 2][lane], kcc);
    ring[0] = ring[1];
    ring[1] = ring[2];
    ring[2] = ring[3];
    ring[3] = ring[4];
    ring[4] = a2;
    // r pass: after plane E of k+1 (t even, t >= 4) the ring holds
    // E[k-1], O[k-1], E[k], O[k], E[k+1] for k = rk0 - 1 + t/2 - 1
    if (!(t & 1) && t >= 4) {
      const int k = rk0 - 2 + (t >> 1);
      if (col_ok && k >= rk0 && k < rk1) {
        T kr[9];
#pragma unroll
        for (int m = 0; m < 9; m++)
          kr[m] = P.mt[0][m * rr + k];
        w_out[(i64)k * P.sw[0] + w_col] =
            mass_trans_k<T>(ring[0], ring[1], ring[2], ring[3], ring[4], kr);
      }
    }
  }
}

} // namespace masstrans3d

// ---------------------------------------------------------------------------------------
// Second formulation (used when the level has enough tiles to fill the GPU with warps):
// one WARP owns a TC x TF tile of coarse (c, f) columns and a segment of coarse r, and
// nothing is shared between warps - no block barrier, no exchange of f-pass results
// through shared memory.
//
//   data movement   every row of a raw plane that the tile needs is one contiguous piece of
//                   the coefficient array (E or O part, 32 nodes).  The warp's lanes fetch
//                   them with the TMA engine: one cp.async.bulk per row part (the 16-byte
//                   aligned superset of the 32 nodes; where the part starts inside it is
//                   noted next to the data), all of a plane on one mbarrier, the next
//                   plane in flight while the current one is worked on.  No per-element
//                   copy instruction, no per-element address or predicate.  (A tensor map
//                   is ruled out: its strides must be multiples of 16 bytes and a row of
//                   the array is 2049 or 513 floats.)
//   f pass          as in the first formulation: lane l holds coarse column kf0 - 1 + l;
//                   E[l-1], O[l-1], E[l], O[l], E[l+1] come straight from the staged row.
//   c pass          the lane walks down the tile's rows E0 O0 E1 O1 E2 | O2 E3 | O3 E4 ...
//                   keeping the last five f-pass values in registers: coarse row j is
//                   mass_trans of rows (E_j, O_j, E_j+1, O_j+1, E_j+2).
//   r pass          five-deep register ring per coarse row while the planes go by.
// Arithmetic and its order are those of the first formulation (mass_trans_k): bit-identical.
// Measured at 257 x 2049 x 2049 (finest level): 2.55 ms against 2.93 ms for the block
// formulation.  (Also measured: the same warp-per-tile walk with plain coalesced loads
// into registers, refilled a plane ahead, instead of the staged rows - 3.2 ms: with 128
// registers per thread only 16 warps are resident and nothing covers the load latency.)
// ---------------------------------------------------------------------------------------
namespace masstrans3d {

constexpr int W_NST = 2;          // planes staged per warp (one worked on, one in flight)
constexpr int W_NPART = 2 * NROW; // row parts of a plane: row r -> parts 2r (E), 2r + 1 (O)
constexpr int W_NW = 8;           // warps per block (independent of each other)

template <typename T> struct WLayout {
  static constexpr int A = 16 / (int)sizeof(T); // elements per 16 bytes
  static constexpr int CE = 32 + A;             // elements fetched per row part
  static constexpr int PITCH = CE + A;          // + front padding (lane -1 of the first f tile)
  static constexpr int STAGE = W_NPART * PITCH; // elements per staged plane
};

template <typename T> struct WParams {
  int n[3], nc[3];
  i64 sin[3], sw[3];
  const T *mt[3];
  int ctiles, ftiles, rsegs, per; // per: coarse r indices per segment
  i64 total; // elements of the coefficient array.  Reads stay below its 16-byte rounded end: up to
             // 12 bytes behind the array when its size is not a multiple of 16 - the library's own
             // coefficient buffer is allocated 16 bytes larger; a caller's buffer (mgb_decompose)
             // comes from an allocator with a granularity of 256 bytes or more
};

__device__ __forceinline__ unsigned w_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w_mbar_init(unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void w_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void w_bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <typename T> __host__ __device__ inline size_t warp_smem_bytes(int per) {
  typedef WLayout<T> LY;
  const size_t tab_elems = (size_t)TC * 12 + (size_t)per * 9;
  return (size_t)W_NST * LY::STAGE * sizeof(T) + W_NST * W_NPART * sizeof(int) +
         ((tab_elems * sizeof(T) + 15) & ~(size_t)15);
}
template <typename T> __host__ __device__ inline size_t warp_smem_base() {
  typedef WLayout<T> LY;
  return 128 + ((LY::PITCH * sizeof(T) + 127) & ~(size_t)127); // mbarriers | zero row
}

// EDGE: some lane of the tile has no E or O node (first / last f tile): mask what is read
template <typename T, bool EDGE>
__device__ __forceinline__ void masstrans3d_warp_body(const WParams<T> &P, const T *__restrict__ in,
                                                      T *__restrict__ w_out, unsigned char *smem_raw, int tile,
                                                      int rs) {
  typedef WLayout<T> LY;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ft = tile % P.ftiles, ct = tile / P.ftiles;
  const int kc0 = ct * TC, kf0 = ft * TF;
  const int nr = P.n[0], ncn = P.n[1], nf = P.n[2];
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  const int rk0 = rs * P.per, rk1 = min(rr, rk0 + P.per);
  if (rk0 >= rk1)
    return;
  // ---- shared memory of this warp: W_NST stages | part shifts | c / r tables ------------
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
  T *zero_row = reinterpret_cast<T *>(smem_raw + 128);
  unsigned char *mine = smem_raw + warp_smem_base<T>() + (size_t)wid * warp_smem_bytes<T>(P.per);
  T *stage0 = reinterpret_cast<T *>(mine);
  int *s_shift = reinterpret_cast<int *>(mine + (size_t)W_NST * LY::STAGE * sizeof(T));
  T *s_kc = reinterpret_cast<T *>(mine + (size_t)W_NST * LY::STAGE * sizeof(T) + W_NST * W_NPART * sizeof(int));
  T *s_kr = s_kc + TC * 12;
  const unsigned bar0 = w_smem(bars + wid * W_NST);
  // ---- per-lane constants ---------------------------------------------------------------
  const int kf = kf0 - 1 + lane;
  const bool own_f = lane >= 1 && lane <= TF && kf < ff;
  T kfc[9];
#pragma unroll
  for (int m = 0; m < 9; m++)
    kfc[m] = (kf >= 0 && kf < ff) ? P.mt[2][m * ff + kf] : (T)0;
  // validity of the five inputs of the f pass (EDGE tiles only)
  bool vEm = true, vOm = true, vE0 = true, vO0 = true, vEp = true;
  if (EDGE) {
    vEm = pos(kf - 1, false, nf, ff) >= 0;
    vOm = pos(kf - 1, true, nf, ff) >= 0;
    vE0 = pos(kf, false, nf, ff) >= 0;
    vO0 = pos(kf, true, nf, ff) >= 0;
    vEp = pos(kf + 1, false, nf, ff) >= 0;
  }
  // tables of the tile: c constants [TC][12] (9 used), r constants [per][9]
  for (int i = lane; i < TC * 9; i += 32) {
    const int j = i / 9, m = i - j * 9;
    s_kc[j * 12 + m] = (kc0 + j < cc) ? P.mt[1][m * cc + kc0 + j] : (T)0;
  }
  for (int i = lane; i < (rk1 - rk0) * 9; i += 32) {
    const int j = i / 9, m = i - j * 9;
    s_kr[j * 9 + m] = P.mt[0][m * rr + rk0 + j];
  }
  // rows that do not exist stay zero in every stage (never fetched)
  for (int i = lane; i < W_NST * LY::STAGE; i += 32)
    stage0[i] = (T)0;
  for (int i = lane; i < W_NST * W_NPART; i += 32)
    s_shift[i] = 0;
  if (lane < W_NST)
    w_mbar_init(bar0 + lane * 8);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  // ---- producer: lane L fetches parts L and 32 + L of a plane ---------------------------
  // part p = 2 row + (O ? 1 : 0); rows 0 .. TC+1 are E rows (c index kc0-1+row), rows
  // TC+2 .. NROW-1 are O rows (kc0-1+row-(TC+2))
  int poff[2];    // element offset of the part's window from the plane base; < 0: no such row
  bool p_eofe[2]; // E part of an E row: skipped on even-r planes (the all-coarse block is zero)
  const int ws_e = max(kf0 - 1, 0), ws_o = ff + kf0 - 1; // first node of the E / O window
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int part = lane + 32 * q;
    poff[q] = -1;
    p_eofe[q] = false;
    if (part < W_NPART) {
      const int row = part >> 1;
      const bool opart = part & 1, orow = row >= TC + 2;
      const int pc = pos(kc0 - 1 + (orow ? row - (TC + 2) : row), orow, ncn, cc);
      const int ws = opart ? ws_o : ws_e;
      // a part whose window starts beyond its nodes holds no node of this tile
      if (pc >= 0 && ws < nf && (opart || ws < ff))
        poff[q] = (int)((i64)pc * P.sin[1] + ws);
      p_eofe[q] = !opart && !orow;
    }
  }
  const uintptr_t hi_lim = ((uintptr_t)(in + P.total) + 15) & ~(uintptr_t)15;
  auto fetch = [&](int t) {
    const int k = rk0 - 1 + (t >> 1);
    const bool rodd = t & 1;
    const int pr = pos(k, rodd, nr, rr);
    const int st = t & (W_NST - 1);
    unsigned bytes_total = 0;
    if (pr >= 0) {
      const T *base = in + (i64)pr * P.sin[0];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int part = lane + 32 * q;
        if (poff[q] >= 0 && !(p_eofe[q] && !rodd)) {
          const uintptr_t a = (uintptr_t)(base + poff[q]);
          const uintptr_t al = a & ~(uintptr_t)15;
          unsigned bytes = (unsigned)(LY::CE * sizeof(T));
          if (al + bytes > hi_lim)
            bytes = (unsigned)(hi_lim - al);
          const unsigned dst = w_smem(stage0 + (size_t)st * LY::STAGE + part * LY::PITCH + LY::A);
          w_bulk_load(dst, (const void *)al, bytes, bar0 + st * 8);
          s_shift[st * W_NPART + part] = (int)(a - al); // bytes
          bytes_total += bytes;
        }
      }
    }
    bytes_total = __reduce_add_sync(0xffffffffu, bytes_total);
    if (lane == 0)
      w_mbar_expect(bar0 + st * 8, bytes_total);
  };
  // ---- consumer -------------------------------------------------------------------------
  T ring[TC][5];
#pragma unroll
  for (int j = 0; j < TC; j++)
#pragma unroll
    for (int m = 0; m < 5; m++)
      ring[j][m] = (T)0;
  const int nplanes = 2 * (rk1 - rk0 + 1) + 1;
  const int adj_e = kf0 == 0 ? -(int)sizeof(T) : 0; // lane 0 of the first f tile is node -1
  const unsigned zero_addr = w_smem(zero_row + LY::A);
  const i64 w_col0 = (i64)kc0 * P.sw[1] + (i64)kf * P.sw[2];
  fetch(0);
#pragma unroll 1
  for (int t = 0; t < nplanes; t++) {
    const int st = t & (W_NST - 1);
    if (t + 1 < nplanes)
      fetch(t + 1); // its stage was released at the end of the previous iteration
    w_mbar_wait(bar0 + st * 8, (t / W_NST) & 1);
    const int k = rk0 - 1 + (t >> 1);
    const bool rodd = t & 1;
    const bool plane_ok = pos(k, rodd, nr, rr) >= 0;
    const unsigned sbase = w_smem(stage0 + (size_t)st * LY::STAGE + LY::A);
    const int *sh = s_shift + st * W_NPART;
    // f pass of row `row` (compile-time), result in registers
    auto fpass = [&](int row) -> T {
      const bool orow = row >= TC + 2;
      // E part of an E row on an even-r plane, and every part of a missing plane, read zeros
      unsigned ae, ao;
      if (!plane_ok) {
        ae = ao = zero_addr + lane * (unsigned)sizeof(T);
      } else {
        ae = (!orow && !rodd) ? zero_addr + lane * (unsigned)sizeof(T)
                              : sbase + (unsigned)((2 * row) * LY::PITCH * sizeof(T)) + (unsigned)sh[2 * row] +
                                    lane * (unsigned)sizeof(T) + adj_e;
        ao = sbase + (unsigned)((2 * row + 1) * LY::PITCH * sizeof(T)) + (unsigned)sh[2 * row + 1] +
             lane * (unsigned)sizeof(T);
      }
      T a, b, c, d, e;
      if (sizeof(T) == 4) {
        asm volatile("ld.shared.f32 %0, [%1+-4];" : "=f"(*(float *)&a) : "r"(ae));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(*(float *)&c) : "r"(ae));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(*(float *)&e) : "r"(ae));
        asm volatile("ld.shared.f32 %0, [%1+-4];" : "=f"(*(float *)&b) : "r"(ao));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(*(float *)&d) : "r"(ao));
      } else {
        asm volatile("ld.shared.f64 %0, [%1+-8];" : "=d"(*(double *)&a) : "r"(ae));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(*(double *)&c) : "r"(ae));
        asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(*(double *)&e) : "r"(ae));
        asm volatile("ld.shared.f64 %0, [%1+-8];" : "=d"(*(double *)&b) : "r"(ao));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(*(double *)&d) : "r"(ao));
      }
      if (EDGE) {
        a = vEm ? a : (T)0;
        b = vOm ? b : (T)0;
        c = vE0 ? c : (T)0;
        d = vO0 ? d : (T)0;
        e = vEp ? e : (T)0;
      }
      return mass_trans_k<T>(a, b, c, d, e, kfc);
    };
    // walk down the rows: E0 O0 E1 O1 E2 -> c row 0; then (O_j+1, E_j+2) -> c row j
    T f0 = fpass(0), f1 = fpass(TC + 2), f2 = fpass(1), f3, f4;
#pragma unroll
    for (int j = 0; j < TC; j++) {
      f3 = fpass(TC + 2 + j + 1);
      f4 = fpass(j + 2);
      T kc[9];
#pragma unroll
      for (int m = 0; m < 9; m++)
        kc[m] = s_kc[j * 12 + m];
      const T a2 = mass_trans_k<T>(f0, f1, f2, f3, f4, kc);
      f0 = f2;
      f1 = f3;
      f2 = f4;
      ring[j][0] = ring[j][1];
      ring[j][1] = ring[j][2];
      ring[j][2] = ring[j][3];
      ring[j][3] = ring[j][4];
      ring[j][4] = a2;
    }
    // this stage may be refilled: every lane has read what it needs
    __syncwarp();
    // r pass: after plane E of k+1 (t even, t >= 4) the ring holds
    // E[k-1], O[k-1], E[k], O[k], E[k+1] for k = rk0 - 2 + t/2
    if (!(t & 1) && t >= 4) {
      const int ko = rk0 - 2 + (t >> 1);
      if (ko >= rk0 && ko < rk1) {
        T kr[9];
#pragma unroll
        for (int m = 0; m < 9; m++)
          kr[m] = s_kr[(ko - rk0) * 9 + m];
        if (own_f) {
#pragma unroll
          for (int j = 0; j < TC; j++)
            if (kc0 + j < cc)
              w_out[(i64)ko * P.sw[0] + w_col0 + (i64)j * P.sw[1]] =
                  mass_trans_k<T>(ring[j][0], ring[j][1], ring[j][2], ring[j][3], ring[j][4], kr);
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(W_NW * 32, sizeof(T) == 4 ? 2 : 1)
masstrans3d_warp_kernel(const WParams<T> P, const T *__restrict__ in, T *__restrict__ w_out) {
  extern __shared__ __align__(128) unsigned char mt_smem[];
  typedef WLayout<T> LY;
  // block-shared zero row
  T *zero_row = reinterpret_cast<T *>(mt_smem + 128);
  for (int i = threadIdx.x; i < LY::PITCH; i += blockDim.x)
    zero_row[i] = (T)0;
  __syncthreads();
  const int ntiles = P.ctiles * P.ftiles;
  const long long w = (long long)blockIdx.x * W_NW + (threadIdx.x >> 5);
  if (w >= (long long)ntiles * P.rsegs)
    return;
  const int tile = (int)(w % ntiles), rs = (int)(w / ntiles);
  const int kf0 = (tile % P.ftiles) * TF;
  // every lane of the tile has its E and O node and both neighbours
  const bool interior = kf0 >= 1 && kf0 + TF + 1 <= P.nc[2] && P.nc[2] + kf0 + TF + 1 <= P.n[2];
  if (interior)
    masstrans3d_warp_body<T, false>(P, in, w_out, mt_smem, tile, rs);
  else
    masstrans3d_warp_body<T, true>(P, in, w_out, mt_smem, tile, rs);
}

} // namespace masstrans3d
