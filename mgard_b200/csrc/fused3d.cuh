// Fused 3-D level kernel (sm_100a): coefficients + mass x restriction along
// f, c and r in ONE pass over the level box.
//
// Replaces, for D == 3, the reference's per-level chain
//   CopyND -> GpkReo3D -> Lpk1Reo3D -> Lpk2Reo3D -> Lpk3Reo3D
// (DataRefactoring.hpp:85-96, GridProcessingKernel3D.hpp:21-1229,
//  LinearProcessingKernel3D.hpp:27-1090) with the same arithmetic in the same
// order (lerp f,c,r; mass_trans f,c,r), so results stay bit-identical, but the
// level box is read once and only the coefficients, the coarse nodes and the
// (n/8-sized) load vector are written.
//
// Shape of the kernel: a thread block owns a (TC x TF) tile of coarse (c, f)
// columns and a segment of coarse r indices, and sweeps the nodal r planes of
// that segment.  Per plane pair it stages the raw (2TC+3) x (2TF+3) nodal tiles
// in shared memory (rows are contiguous segments of the dense level box),
// forms the coefficient plane, applies the f and c passes in shared memory and
// keeps the last five c-pass planes of its own output column in REGISTERS,
// from which the r pass produces one load-vector value per coarse plane.
// Halo re-reads exist only in c and f ((2TC+3)(2TF+3)/(4 TC TF) = 1.24x for
// 8 x 32) and hit L2.  All index arithmetic (ghost / hole mapping, coarse-first
// positions, ownership) is hoisted out of the plane loop into per-thread
// registers; the 9 mass_trans constants of a thread's fixed f column and c row
// live in registers too.
//
// MODE 0 (decomposition): input = dense nodal box; writes coefficients to the
//         coarse-first layout and the coarse nodes to the next dense box.
// MODE 1 (recomposition): input = coefficients in the coarse-first layout
//         (all-coarse block treated as zero); only the load vector is written.
#pragma once

namespace fused3d {

typedef long long i64;

constexpr int TC = 8, TF = 32, NT = 256;
constexpr int PC = 2 * TC + 3, PF = 2 * TF + 3;
constexpr int CA = TC + 2, CB = TF + 2; // 2x2 cells per plane
constexpr int NCELL = (CA * CB + NT - 1) / NT;
constexpr int LROWS = (PC + NT / 32 - 1) / (NT / 32); // rows per thread in row sweeps
constexpr int LCOLS = (PF + 31) / 32;
constexpr int NLOAD = (PC * PF + NT - 1) / NT; // load slots per thread and plane

template <typename T> struct Params {
  int n[3], nc[3];   // fine / coarse level shape (r, c, f)
  int np[3];         // padded nodal sizes 2*nc-1
  i64 sin[3];        // input strides (MODE 0: dense nodal; MODE 1: coefficient array)
  i64 sout[3];       // coefficient array strides (MODE 0)
  i64 scoarse[3];    // dense coarse box strides (MODE 0 coarse out)
  i64 sw[3];         // dense load-vector strides
  const T *ratio[3]; // level-l ratio tables
  const T *mt[3];    // 9 x nc mass_trans tables
  int rsegs;         // segments along coarse r
  int ctiles, ftiles;
};

template <typename T> __device__ __forceinline__ T lerp_ref(T v0, T v1, T t) {
  T r = v0 + v0 * t * (T)-1;
  r = r + t * v1;
  return r;
}

// mass_trans with its nine precomputed constants (LPKFunctor.h:47-66):
// k = {h1/6,(h1+h2)/3,h2/6,(h2+h3)/3,h3/6,(h3+h4)/3,h4/6,r1,r4}
template <typename T>
__device__ __forceinline__ T mass_trans_k(T a, T b, T c, T d, T e, const T (&k)[9]) {
  T tb = a * k[0] + b * k[1] + c * k[2];
  T tc = b * k[2] + c * k[3] + d * k[4];
  T td = c * k[4] + d * k[5] + e * k[6];
  tc += tb * k[7] + td * k[8];
  return tc;
}

// padded nodal index -> source nodal index; -1: hole / out of range (value 0)
__device__ __forceinline__ int src_index(int j, int n, int np) {
  if (j < 0 || j >= np)
    return -1;
  if ((n & 1) == 0) {
    if (j == n)
      return n - 1; // ghost: replicate the last node
    if (j == n - 1)
      return -1; // the slot the ghost displaced
  }
  return j;
}
// padded nodal index -> position in the coarse-first layout (valid j only)
__device__ __forceinline__ int oct_pos(int j, int nc) {
  return (j & 1) ? nc + (j >> 1) : (j >> 1);
}

__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc, int bytes,
                                         bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? bytes : 0; // src-size 0: the destination is zero filled
  if (bytes == 4)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

constexpr int NSLOT = 5; // raw plane ring: current pair (3 planes) + prefetched pair (2)

template <typename T, int MODE>
__global__ void __launch_bounds__(NT, 3)
level_kernel(const Params<T> P, const T *__restrict__ in, T *__restrict__ coef_out,
             T *__restrict__ coarse_out, T *__restrict__ w_out) {
  extern __shared__ unsigned char smem_raw[];
  T *s_raw = (T *)smem_raw;               // NSLOT planes of PC x PF
  T *s_a1 = s_raw + NSLOT * PC * PF;      // 2 x (PC x TF) f-pass outputs
  T *s_kc = s_a1 + 2 * PC * TF;           // 9 x TC
  T *s_w = s_kc + 9 * TC;                 // MODE 0: 2 coefficient planes

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  int bid = blockIdx.x;
  const int ft = bid % P.ftiles;
  bid /= P.ftiles;
  const int ct = bid % P.ctiles;
  const int rs = bid / P.ctiles;
  const int c0 = ct * TC, f0 = ft * TF;
  const int nr = P.n[0], ncn = P.n[1], nf = P.n[2];
  const int rr = P.nc[0], cc = P.nc[1], ff = P.nc[2];
  const int npr = P.np[0], npc = P.np[1], npf = P.np[2];
  const int per = (rr + P.rsegs - 1) / P.rsegs;
  const int rk0 = rs * per, rk1 = min(rr, rk0 + per);
  if (rk0 >= rk1)
    return;
  const int jc0 = 2 * c0 - 2, jf0 = 2 * f0 - 2; // tile origin (padded nodal)
  const int jstart = 2 * rk0 - 2;

  // ---- per-thread constants (hoisted out of the plane loop) ---------------
  // (1) plane loads: element e = tid + NT * q of the PC x PF tile (row-major), so
  // every slot is used and a warp reads runs of consecutive addresses.
  // ld_off[q]: offset from the plane's base (-1: hole / outside -> zero fill);
  // MODE 1: ld_even bit q set <=> row and column are both even (the all-coarse
  // block counts as zero on even planes).
  int ld_off[NLOAD];
  unsigned ld_even = 0;
#pragma unroll
  for (int q = 0; q < NLOAD; q++) {
    const int e = tid + q * NT;
    ld_off[q] = -1;
    if (e < PC * PF) {
      const int lc = e / PF, lf = e - lc * PF;
      const int jc = jc0 + lc, jf = jf0 + lf;
      const int scx = src_index(jc, ncn, npc), sf = src_index(jf, nf, npf);
      if (scx >= 0 && sf >= 0)
        ld_off[q] = (int)((MODE == 0 ? scx : oct_pos(jc, cc)) * P.sin[1] +
                          (MODE == 0 ? sf : oct_pos(jf, ff)) * P.sin[2]);
      if (!(jc & 1) && !(jf & 1))
        ld_even |= 1u << q;
    }
  }
  // (2) 2x2 cells (MODE 0)
  int cell_s[NCELL];       // smem index of the cell's (even c, even f) node; -1: none
  int cell_o[NCELL];       // offset of that node inside a plane of coef_out
  int cell_c[NCELL];       // ... inside a plane of coarse_out
  unsigned cell_fl[NCELL]; // bits 0-3 exists, 4-7 owned, 8 okf, 9 okc, 10-13 in tile
  T cell_rc[NCELL], cell_rf[NCELL];
  const int o_df = ff * (int)P.sout[2], o_dc = cc * (int)P.sout[1];
  if (MODE == 0) {
#pragma unroll
    for (int q = 0; q < NCELL; q++) {
      const int cell = tid + q * NT;
      cell_s[q] = -1;
      cell_fl[q] = 0;
      cell_c[q] = cell_o[q] = 0;
      cell_rc[q] = cell_rf[q] = (T)0;
      if (cell < CA * CB) {
        const int a = cell / CB, b = cell - a * CB;
        const int lc = 2 * a, lf = 2 * b;
        cell_s[q] = lc * PF + lf;
        unsigned fl = 0;
        if (lf + 2 < PF)
          fl |= 1u << 8;
        if (lc + 2 < PC)
          fl |= 1u << 9;
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const int llc = lc + (m >> 1), llf = lf + (m & 1);
          const int jc = jc0 + llc, jf = jf0 + llf;
          const bool in_tile = llc < PC && llf < PF;
          if (in_tile)
            fl |= 1u << (10 + m);
          const int scx = in_tile ? src_index(jc, ncn, npc) : -1;
          const int sf = in_tile ? src_index(jf, nf, npf) : -1;
          if (scx >= 0 && sf >= 0) {
            fl |= 1u << m;
            if (llc >= 2 && llc < 2 + 2 * TC && llf >= 2 && llf < 2 + 2 * TF)
              fl |= 1u << (4 + m);
          }
        }
        // positions of the (even, even) node; the other three follow by the
        // constant octant strides o_df / o_dc
        cell_o[q] = (int)(((jc0 + lc) >> 1) * P.sout[1] + ((jf0 + lf) >> 1) * P.sout[2]);
        cell_c[q] = (int)(((jc0 + lc) >> 1) * P.scoarse[1] + ((jf0 + lf) >> 1) * P.scoarse[2]);
        cell_fl[q] = fl;
        const int jco = jc0 + lc + 1, jfo = jf0 + lf + 1;
        cell_rc[q] = (jco >= 1 && jco - 1 < ncn) ? P.ratio[1][jco - 1] : (T)0;
        cell_rf[q] = (jfo >= 1 && jfo - 1 < nf) ? P.ratio[2][jfo - 1] : (T)0;
      }
    }
  }
  // (3) mass_trans constants: this thread's f column in registers, c rows in smem
  T kf[9];
#pragma unroll
  for (int m = 0; m < 9; m++)
    kf[m] = (f0 + tx < ff) ? P.mt[2][m * ff + f0 + tx] : (T)0;
  for (int k = tid; k < 9 * TC; k += NT) {
    const int m = k / TC, j = k - m * TC;
    s_kc[k] = (c0 + j < cc) ? P.mt[1][m * cc + c0 + j] : (T)0;
  }
  const bool col_ok = (c0 + ty < cc) && (f0 + tx < ff);
  const i64 w_col = (i64)(c0 + ty) * P.sw[1] + (i64)(f0 + tx) * P.sw[2];

  // ---- helpers -------------------------------------------------------------
  auto slot_idx = [&](int j) -> int { return (j - jstart) % NSLOT; };
  auto slot = [&](int j) -> T * { return s_raw + slot_idx(j) * (PC * PF); };
  const unsigned s_raw_addr = (unsigned)__cvta_generic_to_shared(s_raw);
  // asynchronous copy of nodal plane j (MODE 0) / coefficient plane j (MODE 1)
  auto issue_plane = [&](int j) {
    const unsigned buf = s_raw_addr + (unsigned)(slot_idx(j) * (PC * PF) * (int)sizeof(T));
    const int sr = src_index(j, nr, npr);
    const T *base = in + (i64)(sr >= 0 ? (MODE == 0 ? sr : oct_pos(j, rr)) : 0) * P.sin[0];
    const bool plane_ok = sr >= 0;
    const bool reven = !(j & 1);
#pragma unroll
    for (int q = 0; q < NLOAD; q++) {
      if (tid + q * NT < PC * PF) {
        bool valid = plane_ok && ld_off[q] >= 0;
        if (MODE == 1) // the all-coarse block counts as zero
          valid = valid && !(reven && (ld_even & (1u << q)));
        const T *g = valid ? base + ld_off[q] : in;
        const int sz = valid ? (int)sizeof(T) : 0; // src-size 0: zero fill
        const unsigned d = buf + (unsigned)((tid + q * NT) * (int)sizeof(T));
        if (sizeof(T) == 4)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(g), "r"(sz));
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(g), "r"(sz));
      }
    }
  };
  // MODE 0: coefficient plane `dst` from raw planes; lo/hi: even planes below/above
  auto coef_plane = [&](T *dst, const T *cur, const T *lo, const T *hi, int jr, bool owned_plane) {
    const bool podd = jr & 1;
    const int sr = src_index(jr, nr, npr);
    const T rat_r = (podd && sr >= 0) ? P.ratio[0][jr - 1] : (T)0;
    T *cplane = coef_out + (i64)(sr >= 0 ? oct_pos(jr, rr) : 0) * P.sout[0];
    T *kplane = coarse_out + (i64)(jr >> 1) * P.scoarse[0];
#pragma unroll
    for (int q = 0; q < NCELL; q++) {
      const int sb = cell_s[q];
      if (sb < 0)
        continue;
      const unsigned fl = cell_fl[q];
      const bool okf = fl & (1u << 8), okc = fl & (1u << 9);
      const T rc_ = cell_rc[q], rf_ = cell_rf[q];
      T it[4];
      if (!podd) {
        const T c00 = cur[sb];
        const T c01 = okf ? cur[sb + 2] : (T)0;
        const T c10 = okc ? cur[sb + 2 * PF] : (T)0;
        const T c11 = (okc && okf) ? cur[sb + 2 * PF + 2] : (T)0;
        const T f0_ = lerp_ref(c00, c01, rf_), f1_ = lerp_ref(c10, c11, rf_);
        it[0] = (T)0;
        it[1] = f0_;
        it[2] = lerp_ref(c00, c10, rc_);
        it[3] = lerp_ref(f0_, f1_, rc_);
      } else {
        const T l00 = lo[sb], h00 = hi[sb];
        const T l01 = okf ? lo[sb + 2] : (T)0, h01 = okf ? hi[sb + 2] : (T)0;
        const T l10 = okc ? lo[sb + 2 * PF] : (T)0, h10 = okc ? hi[sb + 2 * PF] : (T)0;
        const T l11 = (okc && okf) ? lo[sb + 2 * PF + 2] : (T)0;
        const T h11 = (okc && okf) ? hi[sb + 2 * PF + 2] : (T)0;
        const T lf0 = lerp_ref(l00, l01, rf_), lf1 = lerp_ref(l10, l11, rf_);
        const T hf0 = lerp_ref(h00, h01, rf_), hf1 = lerp_ref(h10, h11, rf_);
        it[0] = lerp_ref(l00, h00, rat_r);
        it[1] = lerp_ref(lf0, hf0, rat_r);
        it[2] = lerp_ref(lerp_ref(l00, l10, rc_), lerp_ref(h00, h10, rc_), rat_r);
        it[3] = lerp_ref(lerp_ref(lf0, lf1, rc_), lerp_ref(hf0, hf1, rc_), rat_r);
      }
#pragma unroll
      for (int m = 0; m < 4; m++) {
        if (!(fl & (1u << (10 + m))))
          continue;
        const int si = sb + (m >> 1) * PF + (m & 1);
        const bool exists = sr >= 0 && (fl & (1u << m));
        const bool coarse_node = !podd && m == 0;
        const T v = cur[si];
        T wv = (T)0;
        if (exists && !coarse_node)
          wv = v - it[m];
        dst[si] = wv;
        if (owned_plane && exists && (fl & (1u << (4 + m)))) {
          if (coarse_node)
            kplane[cell_c[q]] = v;
          else
            cplane[cell_o[q] + ((m & 1) ? o_df : 0) + ((m >> 1) ? o_dc : 0)] = wv;
        }
      }
    }
  };
  auto pass_f = [&](const T *wplane, T *a1) {
#pragma unroll
    for (int q = 0; q < LROWS; q++) {
      const int lc = ty + q * (NT / 32);
      if (lc < PC) {
        const T *w = wplane + lc * PF + 2 * tx;
        a1[lc * TF + tx] = mass_trans_k<T>(w[0], w[1], w[2], w[3], w[4], kf);
      }
    }
  };
  auto pass_c = [&](const T *a1) -> T {
    const T *a = a1 + (2 * ty) * TF + tx;
    T kc[9];
#pragma unroll
    for (int m = 0; m < 9; m++)
      kc[m] = s_kc[m * TC + ty];
    return mass_trans_k<T>(a[0], a[TF], a[2 * TF], a[3 * TF], a[4 * TF], kc);
  };

  T ring[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};
  auto push = [&](T v) {
    ring[0] = ring[1];
    ring[1] = ring[2];
    ring[2] = ring[3];
    ring[3] = ring[4];
    ring[4] = v;
  };
  auto emit = [&](int k) {
    if (col_ok && k >= rk0 && k < rk1) {
      T kr[9];
#pragma unroll
      for (int m = 0; m < 9; m++)
        kr[m] = P.mt[0][m * rr + k];
      w_out[(i64)k * P.sw[0] + w_col] =
          mass_trans_k<T>(ring[0], ring[1], ring[2], ring[3], ring[4], kr);
    }
  };

  // prologue: planes of the first pair
  issue_plane(jstart);
  issue_plane(jstart + 1);
  issue_plane(jstart + 2);
  cp_async_commit();
  // plane pairs (2k, 2k+1) for k = rk0-1 .. rk1-1, then the single plane 2*rk1
  for (int k = rk0 - 1; k <= rk1; k++) {
    const int je = 2 * k, jo = 2 * k + 1;
    const bool last = (k == rk1);
    // prefetch the next pair's new planes; their slots held planes 2k-2 / 2k-1,
    // whose last readers finished before barrier (B) of the previous iteration
    if (k + 1 < rk1) {
      issue_plane(jo + 2);
      issue_plane(jo + 3);
    } else if (k + 1 == rk1) {
      issue_plane(jo + 2); // only the single trailing even plane... (2*rk1 is je of the last iteration)
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads(); // (A) current pair's planes visible to the whole block
    const bool ev_ok = je >= 0 && je < npr;
    const bool od_ok = !last && jo >= 0 && jo < npr && src_index(jo, nr, npr) >= 0;
    const T *we = slot(je), *wo = slot(jo);
    if (MODE == 0) {
      if (ev_ok)
        coef_plane(s_w, slot(je), nullptr, nullptr, je, je >= 2 * rk0 && je < 2 * rk1);
      if (od_ok)
        coef_plane(s_w + PC * PF, slot(jo), slot(je), slot(jo + 1), jo,
                   jo >= 2 * rk0 && jo < 2 * rk1);
      we = s_w;
      wo = s_w + PC * PF;
      __syncthreads(); // (B) coefficient planes complete
    }
    if (ev_ok)
      pass_f(we, s_a1);
    if (od_ok)
      pass_f(wo, s_a1 + PC * TF);
    __syncthreads(); // (C) f pass complete
    push(ev_ok ? pass_c(s_a1) : (T)0);
    emit(k - 1);
    if (last)
      break;
    push(od_ok ? pass_c(s_a1 + PC * TF) : (T)0);
  }
  cp_async_wait<0>();
}

template <typename T> size_t smem_bytes(int mode) {
  return sizeof(T) * (size_t)(NSLOT * PC * PF + 2 * PC * TF + 9 * TC + (mode == 0 ? 2 * PC * PF : 0));
}

} // namespace fused3d
