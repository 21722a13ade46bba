// GPU Huffman: codebook construction, chunked encode, serialisation, decode
// (sm_100a).  Stream format and code assignment are those of the reference
// (include/mgard-x/Lossless/ParallelHuffman/*), so the bytes are identical
// given identical symbols:
//   codebook   GetCodebook.hpp:23-146, GenerateCL.hpp:29-520, GenerateCW.hpp:38-218
//   encode     EncodeFixedLen.hpp:22-92 + Deflate.hpp:21-77 + Condense.hpp:14-86
//   serialise  Huffman.hpp:130-262 (field order / alignment: RuntimeX/Utilities/Serializer.hpp:13-23)
//   decode     Decode.hpp:66-116, Huffman.hpp:264-362
//
// Structure here: 16-bit symbols; the code-length generation runs in ONE
// thread block (no cooperative grid syncs, no host round trips); encoding
// computes every chunk's bit count, scans the word offsets on the device and
// packs each chunk straight into its final place in the serialised block
// (no fixed-length intermediate, no separate condense pass).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace {

typedef unsigned long long u64;
typedef long long i64;

// ------------------------------- codebook ----------------------------------
struct CbWork {
  // all arrays sized dict (global scratch)
  u64 *keys_sorted;  // (freq << 32 | symbol), padded to pow2
  unsigned *lfreq;   // non-zero freqs ascending
  unsigned *CL;
  int *lleader;
  unsigned *ifreq;
  int *ileader;
  unsigned *tfreq; // temp (merged) arrays
  int *tindex;
  int *tleaf;
  u64 *cw; // codewords in ascending-length order
};

// a mod n for -n < a < 2n (every use below: sums / differences of ring indices)
__device__ __forceinline__ int modn(int a, int n) {
  int r = a >= n ? a - n : a;
  return r < 0 ? r + n : r;
}

// Single-block kernel.  d_codebook[dict], d_decodebook = first[64] entry[64] keys[dict].
// key_smem: the sort buffer lives in dynamic shared memory; ncap: capacity (in
// entries) of the shared-memory copies of the code-length work arrays, used when
// the number of non-zero symbols fits (the loop below is a chain of short
// dependent steps, so its cost is the latency of these arrays).
__global__ void __launch_bounds__(1024)
codebook_kernel(const unsigned *__restrict__ hist, int dict, int npow2, CbWork w,
                u64 *__restrict__ codebook, u64 *__restrict__ decodebook,
                int *__restrict__ status_out, int key_smem, int ncap) {
  extern __shared__ __align__(16) unsigned char cb_smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int s_front, s_rear, s_lcur, s_isize, s_curleaves, s_mfront, s_mrear,
      s_templen, s_first_nz, s_continue;
  __shared__ unsigned s_minfreq;
  __shared__ int s_gs[66], s_ge[66], s_gl[66]; // CW groups
  __shared__ u64 s_gbase[66];
  __shared__ int s_ngroups;

  // 1. sort (freq, symbol) ascending: bitonic sort (shared memory when it fits)
  u64 *key = key_smem ? reinterpret_cast<u64 *>(cb_smem) : w.keys_sorted;
  for (int i = tid; i < npow2; i += nt)
    key[i] = i < dict ? (((u64)hist[i] << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += nt) {
        int ixj = i ^ j;
        if (ixj > i) {
          u64 a = key[i], b = key[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // 2. first non-zero index (GetFirstNonzeroIndex.hpp)
  if (tid == 0)
    s_first_nz = dict;
  __syncthreads();
  for (int i = tid; i < dict; i += nt) {
    bool nzv = (key[i] >> 32) != 0;
    bool prev_zero = i == 0 || (key[i - 1] >> 32) == 0;
    if (nzv && prev_zero)
      s_first_nz = i;
  }
  __syncthreads();
  const int first_nz = s_first_nz;
  const int n = dict - first_nz;
  u64 *first = decodebook, *entry = decodebook + 64, *qkeys = decodebook + 128;
  // keys: symbols by descending frequency (ReverseArray of the sorted qcode)
  for (int i = tid; i < dict; i += nt) {
    qkeys[i] = (unsigned)(key[dict - 1 - i] & 0xffffffffu);
    codebook[i] = 0;
  }
  for (int i = tid; i < 64; i += nt) {
    first[i] = ~0ull;
    entry[i] = ~0ull;
  }
  if (n == 0) {
    if (tid == 0)
      *status_out = 1;
    return;
  }
  // 3. GenerateCL
  if (n <= ncap) {
    unsigned *b = reinterpret_cast<unsigned *>(cb_smem + (key_smem ? (size_t)npow2 * 8 : 0));
    w.lfreq = b;
    w.CL = b + ncap;
    w.lleader = reinterpret_cast<int *>(b + 2 * ncap);
    w.ifreq = b + 3 * ncap;
    w.ileader = reinterpret_cast<int *>(b + 4 * ncap);
    w.tfreq = b + 5 * ncap;
    w.tindex = reinterpret_cast<int *>(b + 6 * ncap);
    w.tleaf = reinterpret_cast<int *>(b + 7 * ncap);
  }
  for (int i = tid; i < n; i += nt) {
    w.lfreq[i] = (unsigned)(key[first_nz + i] >> 32);
    w.CL[i] = 0;
    w.lleader[i] = -1;
  }
  if (tid == 0) {
    s_front = s_rear = s_lcur = s_isize = 0;
    s_continue = 1;
  }
  __syncthreads();
  while (s_continue) {
    if (tid == 0) {
      int lcur = s_lcur, front = s_front, rear = s_rear, isize = s_isize;
      // Operation2 (GenerateCL.hpp:96-219)
      unsigned mf[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      int ml[4] = {0, 0, 0, 0};
      if (lcur < n) { mf[0] = w.lfreq[lcur]; ml[0] = 1; }
      if (lcur < n - 1) { mf[1] = w.lfreq[lcur + 1]; ml[1] = 1; }
      if (isize >= 1) { mf[2] = w.ifreq[front]; ml[2] = 0; }
      if (isize >= 2) { mf[3] = w.ifreq[modn(front + 1, n)]; ml[3] = 0; }
#define CSWAP(a, b)                                                            \
  if (mf[a] > mf[b]) {                                                         \
    unsigned tf = mf[a]; mf[a] = mf[b]; mf[b] = tf;                            \
    int tl = ml[a]; ml[a] = ml[b]; ml[b] = tl;                                 \
  }
      CSWAP(1, 3) CSWAP(0, 2) CSWAP(0, 1) CSWAP(2, 3) CSWAP(1, 2)
#undef CSWAP
      unsigned minfreq = mf[0];
      if (mf[1] < 0xffffffffu)
        minfreq += mf[1];
      w.ifreq[rear] = minfreq;
      w.ileader[rear] = -1;
      for (int k = 0; k < 2; k++) {
        if (mf[k] < 0xffffffffu) {
          if (ml[k]) {
            w.lleader[lcur] = rear;
            w.CL[lcur] += 1;
            lcur++;
          } else {
            w.ileader[front] = rear;
            front = modn(front + 1, n);
          }
        }
      }
      isize = modn(rear - front, n);
      // Operation3: leaves (from lcur) with freq <= minFreq  (sorted => prefix)
      int lo = lcur, hi = n;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (w.lfreq[mid] <= minfreq) lo = mid + 1; else hi = mid;
      }
      int curleaves = lo - lcur;
      // Operation5 (GenerateCL.hpp:237-281)
      int mrear = rear, mfront = front;
      if ((curleaves + isize) % 2 == 0) {
        front = rear;
      } else {
        // histogram[lNodesCur + curLeavesNum]: the reference reads one past
        // the end of the frequency array when every remaining leaf was taken;
        // the word that follows (the zero-initialised codebook on the device
        // layout) is defined here as 0.
        unsigned nextleaf = (lcur + curleaves < n) ? w.lfreq[lcur + curleaves] : 0u;
        if (isize != 0 &&
            (curleaves == 0 || nextleaf <= w.ifreq[modn(rear - 1, n)])) {
          mrear = modn(mrear - 1, n);
          front = modn(rear - 1, n);
        } else {
          front = rear;
          --curleaves;
        }
      }
      s_mfront = mfront;
      s_mrear = mrear;
      s_curleaves = curleaves;
      s_minfreq = minfreq;
      // leaves to merge start at the pre-advance lcur
      s_lcur = lcur; // merge phase uses s_lcur as copy base
      s_front = front;
      s_rear = modn(rear + 1, n);
      s_templen = curleaves + modn(mrear - mfront, n);
    }
    __syncthreads();
    const int A = s_curleaves, cbase = s_lcur, mfront = s_mfront;
    const int B = modn(s_mrear - mfront, n);
    const int templen = s_templen;
    const int rear = s_rear;
    if (templen > 0) {
      // merge by rank (leaf first on ties, GenerateCL.hpp:484)
      for (int k = tid; k < A; k += nt) {
        unsigned f = w.lfreq[cbase + k];
        int lo = 0, hi = B; // number of inodes with freq < f
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (w.ifreq[modn(mfront + mid, n)] < f) lo = mid + 1; else hi = mid;
        }
        int pos = k + lo;
        w.tfreq[pos] = f;
        w.tindex[pos] = cbase + k;
        w.tleaf[pos] = 1;
      }
      for (int m = tid; m < B; m += nt) {
        int bi = modn(mfront + m, n);
        unsigned f = w.ifreq[bi];
        int lo = 0, hi = A; // number of leaves with freq <= f
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (w.lfreq[cbase + mid] <= f) lo = mid + 1; else hi = mid;
        }
        int pos = m + lo;
        w.tfreq[pos] = f;
        w.tindex[pos] = bi;
        w.tleaf[pos] = 0;
      }
      __syncthreads();
      // Operation12: meld
      for (int i = tid; i < templen / 2; i += nt) {
        int ind = modn(rear + i, n);
        w.ifreq[ind] = w.tfreq[2 * i] + w.tfreq[2 * i + 1];
        w.ileader[ind] = -1;
        for (int c = 0; c < 2; c++) {
          int ti = w.tindex[2 * i + c];
          if (w.tleaf[2 * i + c]) {
            w.lleader[ti] = ind;
            w.CL[ti] += 1;
          } else {
            w.ileader[ti] = ind;
          }
        }
      }
      __syncthreads();
    }
    // Operation14: update leaders
    for (int i = tid; i < n; i += nt) {
      int ll = w.lleader[i];
      if (ll != -1) {
        int il = w.ileader[ll];
        if (il != -1) {
          w.lleader[i] = il;
          w.CL[i] += 1;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      s_lcur = s_lcur + s_curleaves;
      s_rear = modn(s_rear + templen / 2, n);
      s_isize = modn(s_rear - s_front, n);
      s_continue = (s_lcur < n) || (s_isize > 1);
    }
    __syncthreads();
  }
  // 4. GenerateCW.  Work in ascending-length order: r = n-1-k (k ascending freq)
  // groups of equal length
  // group boundaries found by warp 0 (32 ranks per step), the rest by its lane 0
  int ng = 0, ok = 0;
  if (tid < 32) {
    int start = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
      const int r = r0 + tid;
      unsigned cl = 0;
      bool last = false;
      if (r < n) {
        cl = w.CL[n - 1 - r];
        last = (r == n - 1) || (w.CL[n - 1 - (r + 1)] != cl);
      }
      unsigned m = __ballot_sync(0xffffffffu, last);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const unsigned clb = __shfl_sync(0xffffffffu, cl, b);
        if (tid == 0 && ng < 66) {
          s_gs[ng] = start;
          s_ge[ng] = r0 + b;
          s_gl[ng] = (int)clb;
        }
        ng++;
        start = r0 + b + 1;
      }
    }
    __syncwarp();
  }
  if (tid == 0) {
    unsigned maxcl = w.CL[0]; // least frequent symbol has the longest code
    if (maxcl > 56)
      ok = 2; // GetCodebook.hpp:111-121: cannot store the codeword
    if (ng > 64)
      ok = 2;
    s_ngroups = ng > 64 ? 0 : ng;
    u64 base = 0;
    for (int gi = 0; gi < s_ngroups; gi++) {
      s_gbase[gi] = base;
      if (gi + 1 < s_ngroups) {
        u64 top = base + (u64)(s_ge[gi] - s_gs[gi]);
        base = (top + 1) << (s_gl[gi + 1] - s_gl[gi]);
      }
    }
    // first / entry
    if (s_ngroups > 0) {
      int L0 = s_gl[0];
      for (int i = 0; i < L0; i++) {
        first[i] = ~0ull;
        entry[i] = 0;
      }
      entry[L0] = 0;
      if (n == 1) {
        // the loop of GenerateCW never runs: Operation2 only
        first[L0] = 0ull ^ ((1ull << L0) - 1);
        if (L0 + 1 < 64)
          entry[L0 + 1] = 1;
      } else {
        u64 cum = 0;
        for (int gi = 0; gi < s_ngroups; gi++) {
          int Lg = s_gl[gi];
          int Lnext = gi + 1 < s_ngroups ? s_gl[gi + 1] : 64;
          u64 top = s_gbase[gi] + (u64)(s_ge[gi] - s_gs[gi]);
          first[Lg] = top ^ ((1ull << Lg) - 1);
          cum += (u64)(s_ge[gi] - s_gs[gi] + 1);
          for (int i = Lg + 1; i < Lnext; i++) {
            first[i] = ~0ull;
            entry[i] = cum;
          }
          if (Lnext < 64)
            entry[Lnext] = cum;
        }
      }
    }
    *status_out = ok;
  }
  __syncthreads();
  for (int gi = 0; gi < s_ngroups; gi++) {
    const int gs = s_gs[gi], ge = s_ge[gi], Lg = s_gl[gi];
    const u64 base = s_gbase[gi];
    for (int r = gs + tid; r <= ge; r += nt) {
      u64 code = base + (u64)(ge - r);
      u64 cwv = (code | ((u64)(Lg & 0xff) << 56)) ^ ((1ull << Lg) - 1);
      // r-th most frequent symbol
      unsigned symbol = (unsigned)(key[dict - 1 - r] & 0xffffffffu);
      codebook[symbol] = cwv;
    }
  }
}

// ------------------------------- encode ------------------------------------

// bits per chunk (sum of code lengths) and words per chunk
__global__ void __launch_bounds__(256)
chunk_bits_kernel(const uint16_t *__restrict__ sym, u64 n, int chunk,
                  const u64 *__restrict__ codebook, int dict,
                  u64 *__restrict__ bits) {
  extern __shared__ unsigned char s_len[];
  __shared__ u64 s_part[8];
  for (int i = threadIdx.x; i < dict; i += blockDim.x)
    s_len[i] = (unsigned char)(codebook[i] >> 56);
  __syncthreads();
  const u64 nchunk = (n - 1) / chunk + 1;
  for (u64 c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const u64 lo = c * (u64)chunk;
    const u64 hi = min(n, lo + (u64)chunk);
    unsigned total = 0; // one thread sees at most chunk / 256 * 56 bits... kept in 64 bits below
    u64 tot64 = 0;
    const uint16_t *base = sym + lo;
    const u64 cnt = hi - lo;
    u64 done = 0;
    if ((((uintptr_t)base) & 15) == 0) {
      const u64 nv = cnt / 8;
      const uint4 *b4 = reinterpret_cast<const uint4 *>(base);
      for (u64 i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint4 v = __ldg(b4 + i);
        total += s_len[v.x & 0xffffu] + s_len[v.x >> 16] + s_len[v.y & 0xffffu] + s_len[v.y >> 16] +
                 s_len[v.z & 0xffffu] + s_len[v.z >> 16] + s_len[v.w & 0xffffu] + s_len[v.w >> 16];
        if (total > 0x7fff0000u) {
          tot64 += total;
          total = 0;
        }
      }
      done = nv * 8;
    }
    for (u64 i = done + threadIdx.x; i < cnt; i += blockDim.x)
      total += s_len[base[i]];
    tot64 += total;
    for (int o = 16; o > 0; o >>= 1)
      tot64 += __shfl_xor_sync(0xffffffffu, tot64, o);
    if ((threadIdx.x & 31) == 0)
      s_part[threadIdx.x >> 5] = tot64;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t = 0;
      for (int k = 0; k < 8; k++)
        t += s_part[k];
      bits[c] = t;
    }
    __syncthreads();
  }
}

// exclusive scan of words per chunk; writes the serialised metadata too.
// scal[1] = total words, scal[2] = overflow flag
__global__ void __launch_bounds__(1024)
chunk_scan_kernel(const u64 *__restrict__ bits, u64 nchunk, u64 *__restrict__ woff,
                  u64 *__restrict__ scal, u64 fixed_bytes, u64 cap,
                  const u64 *__restrict__ ocount_ptr, u64 ocount_fixed, u64 ocap) {
  __shared__ u64 s_warp[32];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0)
    s_carry = 0;
  __syncthreads();
  for (u64 base = 0; base < nchunk; base += blockDim.x) {
    u64 i = base + threadIdx.x;
    u64 wv = 0;
    if (i < nchunk) {
      u64 b = bits[i];
      wv = (b - 1) / 64 + 1; // Huffman.hpp:150 (wraps for b == 0 like the reference)
    }
    u64 x = wv;
    for (int o = 1; o < 32; o <<= 1) {
      u64 y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o)
        x += y;
    }
    if ((threadIdx.x & 31) == 31)
      s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      u64 v = s_warp[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        u64 y = __shfl_up_sync(0xffffffffu, v, o);
        if (threadIdx.x >= o)
          v += y;
      }
      s_warp[threadIdx.x] = v;
    }
    __syncthreads();
    u64 wprefix = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
    u64 incl = s_carry + wprefix + x;
    if (i < nchunk)
      woff[i] = incl - wv;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1)
      s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    u64 total = s_carry;
    u64 oc = ocount_ptr ? *ocount_ptr : ocount_fixed;
    scal[1] = total;
    u64 need = fixed_bytes + 8 * total + 8 + 16 * oc;
    // 1: output too small; 2: the outlier buffer overflowed (the caller grows
    // it and quantizes again) -- either way nothing more is written
    scal[2] = oc > ocap ? 2 : (need > cap ? 1 : 0);
    scal[3] = need;
    woff[nchunk] = total;
  }
}

// Block per chunk, tiles of 2048 symbols.  A thread takes 8 consecutive
// symbols (one 128-bit load), concatenates their codewords in a register and
// places them at its bit offset (block scan of the code lengths) in the tile's
// shared-memory image: words it fills completely are stored, the (at most two)
// words it shares with its neighbours are OR-ed in atomically.  Full words of
// the tile are then flushed, coalesced, to their FINAL place in the stream.
constexpr int ENC_PER = 16; // symbols per thread and tile (two 128-bit loads)
template <bool CB_SHARED>
__global__ void __launch_bounds__(256)
encode_kernel(const uint16_t *__restrict__ sym, u64 n, int chunk,
              const u64 *__restrict__ codebook, int dict,
              const u64 *__restrict__ woff, const u64 *__restrict__ scal,
              u64 *__restrict__ ddata) {
  constexpr int PER = ENC_PER, NT = 256, TILE = PER * NT;
  constexpr int TILE_WORDS = TILE * 56 / 64 + 4;
  extern __shared__ u64 smem[];
  u64 *s_cb = smem;                         // dict (if CB_SHARED)
  u64 *s_out = smem + (CB_SHARED ? dict : 0); // TILE_WORDS
  __shared__ unsigned s_scan[8];
  if (scal[2])
    return; // output too small: nothing is written
  if (CB_SHARED) {
    for (int i = threadIdx.x; i < dict; i += NT)
      s_cb[i] = codebook[i];
  }
  const u64 nchunk = (n - 1) / chunk + 1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  auto cbv = [&](unsigned sy) -> u64 { return CB_SHARED ? s_cb[sy] : __ldg(codebook + sy); };
  for (u64 c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const u64 lo = c * (u64)chunk;
    const u64 hi = min(n, lo + (u64)chunk);
    u64 *dst = ddata + woff[c];
    for (int i = tid; i < TILE_WORDS; i += NT)
      s_out[i] = 0;
    __syncthreads();
    unsigned carry_bits = 0; // bits already in s_out[0]
    u64 wdone = 0;
    for (u64 t0 = lo; t0 < hi; t0 += TILE) {
      const u64 s0 = t0 + (u64)tid * PER;
      u64 cw[PER];
      unsigned mybits = 0;
      if (s0 + PER <= hi && ((((uintptr_t)(sym + s0)) & 15) == 0)) {
        uint4 v[PER / 8];
#pragma unroll
        for (int q = 0; q < PER / 8; q++)
          v[q] = __ldcs(reinterpret_cast<const uint4 *>(sym + s0) + q);
        const unsigned *w = reinterpret_cast<const unsigned *>(v);
#pragma unroll
        for (int k = 0; k < PER; k++)
          cw[k] = cbv((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
      } else {
#pragma unroll
        for (int k = 0; k < PER; k++)
          cw[k] = s0 + k < hi ? cbv(sym[s0 + k]) : 0ull;
      }
#pragma unroll
      for (int k = 0; k < PER; k++)
        mybits += (unsigned)(cw[k] >> 56);
      // block exclusive scan of mybits
      unsigned x = mybits;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
          x += y;
      }
      if (lane == 31)
        s_scan[wid] = x;
      __syncthreads();
      unsigned wpre = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        unsigned v = s_scan[k];
        if (k < wid)
          wpre += v;
        tot += v;
      }
      const unsigned pos = carry_bits + wpre + x - mybits;
      // concatenate in a register, emit word by word
      if (mybits) {
        unsigned wi = pos >> 6, fill = pos & 63;
        bool whole = fill == 0; // the current word started with this thread
        u64 acc = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
          const unsigned len = (unsigned)(cw[k] >> 56);
          if (len) {
            const u64 code = cw[k] & 0x00ffffffffffffffull;
            const unsigned room = 64 - fill;
            if (len < room) {
              acc |= code << (room - len);
              fill += len;
            } else {
              acc |= code >> (len - room);
              if (whole)
                s_out[wi] = acc;
              else
                atomicOr(&s_out[wi], acc);
              wi++;
              whole = true;
              const unsigned rem = len - room;
              acc = rem ? code << (64 - rem) : 0ull;
              fill = rem;
            }
          }
        }
        if (fill)
          atomicOr(&s_out[wi], acc);
      }
      __syncthreads();
      const unsigned tile_bits = carry_bits + tot;
      const unsigned nfull = tile_bits >> 6;
      for (unsigned i = tid; i < nfull; i += NT)
        __stcs(dst + wdone + i, s_out[i]);
      const u64 partial = s_out[nfull];
      __syncthreads();
      for (unsigned i = tid; i <= nfull + 1 && i < (unsigned)TILE_WORDS; i += NT)
        s_out[i] = 0;
      __syncthreads();
      if (tid == 0)
        s_out[0] = partial;
      wdone += nfull;
      carry_bits = tile_bits & 63;
      __syncthreads();
    }
    if (carry_bits && tid == 0)
      dst[wdone] = s_out[0];
    __syncthreads();
  }
}

// serialised scalar fields + chunk metadata (Huffman.hpp:163-239)
__global__ void serialize_meta_kernel(unsigned char *__restrict__ out, u64 n, int dict,
                                      int chunk, u64 nchunk,
                                      const u64 *__restrict__ bits,
                                      const u64 *__restrict__ woff,
                                      const u64 *__restrict__ decodebook,
                                      const u64 *__restrict__ scal, u64 off_ddata,
                                      const u64 *__restrict__ ocount_ptr, u64 ocount_fixed,
                                      const uint64_t *__restrict__ oidx,
                                      const i64 *__restrict__ oval) {
  if (scal[2])
    return;
  u64 *o64 = (u64 *)out;
  const u64 gt = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 gs = (u64)gridDim.x * blockDim.x;
  const u64 total_words = scal[1];
  const u64 oc = ocount_ptr ? *ocount_ptr : ocount_fixed;
  if (gt == 0) {
    o64[0] = n;
    ((int *)out)[2] = dict;
    ((int *)out)[3] = chunk;
    o64[2] = 2 * nchunk;
    o64[3 + 2 * nchunk] = 8ull * 128 + 8ull * dict;
    o64[off_ddata / 8 - 1] = total_words;
    o64[off_ddata / 8 + total_words] = oc;
  }
  for (u64 i = gt; i < nchunk; i += gs) {
    o64[3 + i] = bits[i];
    o64[3 + nchunk + i] = woff[i];
  }
  u64 *db = o64 + 4 + 2 * nchunk;
  for (u64 i = gt; i < 128 + (u64)dict; i += gs)
    db[i] = decodebook[i];
  u64 *oo = o64 + off_ddata / 8 + total_words + 1;
  for (u64 i = gt; i < oc; i += gs) {
    oo[i] = oidx[i];
    oo[oc + i] = (u64)oval[i];
  }
}

// ------------------------------- decode ------------------------------------
// The reference decodes one chunk per thread, bit by bit (Decode.hpp:66-116).
// The stream format only guarantees that a CHUNK starts on a codeword (and
// word) boundary, so here a thread block owns a chunk and finds the interior
// codeword boundaries itself:
//   1. the chunk's bit string is cut into 128-bit sub-sequences; every thread
//      decodes its sub-sequences speculatively from their first bit;
//   2. each sub-sequence then restarts from where its predecessor really ended
//      until nothing changes (Huffman codes self-synchronise after a few
//      codewords, so this takes 2-3 rounds in practice and is exact after at
//      most #sub-sequences rounds);
//   3. a block scan of the symbol counts gives every sub-sequence its output
//      offset; a last decode pass writes the symbols into shared memory and the
//      chunk is flushed to global memory with coalesced 128-bit stores.
// Codewords are resolved through a 2^K-entry shared-memory table built from
// first/entry/keys, falling back to the canonical walk for longer codes.
constexpr int DEC_K = 12;     // LUT index bits
constexpr int DEC_SB = 128;   // sub-sequence length in bits
constexpr int DEC_T = 256;    // threads per block

struct BitReader {
  const u64 *w;
  u64 nw, wi, cur, nxt;
  unsigned used;
  __device__ __forceinline__ void init(const u64 *words, u64 nwords, u64 bitpos) {
    w = words;
    nw = nwords;
    wi = bitpos >> 6;
    used = (unsigned)(bitpos & 63);
    cur = wi < nw ? __ldg(w + wi) : 0ull;
    nxt = wi + 1 < nw ? __ldg(w + wi + 1) : 0ull;
  }
  __device__ __forceinline__ u64 window() const {
    return used ? ((cur << used) | (nxt >> (64 - used))) : cur;
  }
  __device__ __forceinline__ void advance(unsigned l) {
    used += l;
    if (used >= 64) {
      used -= 64;
      wi++;
      cur = nxt;
      nxt = wi + 1 < nw ? __ldg(w + wi + 1) : 0ull;
    }
  }
};

struct DecTables {
  const u64 *first, *entry; // shared
  const unsigned *lut;      // shared: symbol | len << 16 (len 0: longer than K)
  const uint16_t *keys;     // shared
  int dict, lslow;
};

__device__ __forceinline__ unsigned decode_one(const DecTables &t, u64 window, unsigned &sym) {
  unsigned e = t.lut[window >> (64 - DEC_K)];
  unsigned len = e >> 16;
  if (len) {
    sym = e & 0xffffu;
    return len;
  }
  int l = t.lslow;
  u64 v = window >> (64 - l);
  while (v < t.first[l] && l < 63) {
    l++;
    v = window >> (64 - l);
  }
  u64 ki = t.entry[l] + v - t.first[l];
  sym = ki < (u64)t.dict ? t.keys[ki] : 0u;
  return (unsigned)l;
}

// decode codewords that START in [start, limit); returns the end position
// (first boundary >= limit) and the number of codewords.
template <bool WRITE>
__device__ __forceinline__ unsigned
decode_sub(const DecTables &t, const u64 *words, u64 nw, unsigned start, unsigned limit,
           unsigned &count, uint16_t *dst, unsigned dst_cap) {
  BitReader br;
  br.init(words, nw, start);
  unsigned p = start, c = 0;
  while (p < limit) {
    unsigned sym;
    unsigned l = decode_one(t, br.window(), sym);
    if (WRITE) {
      if (c < dst_cap)
        dst[c] = (uint16_t)sym;
    }
    c++;
    p += l;
    br.advance(l);
  }
  count = c;
  return p;
}

// Flush of a decoded chunk staged in shared memory: symbols as they are, or
// (DEQ) dequantized on the way out: v = scale * (T)(symbol - dict / 2), the
// arithmetic of dequantize_linear_kernel (LinearQuantization.hpp:251-264).
template <typename OUT, int NT>
__device__ __forceinline__ void flush_chunk(const uint16_t *s_out, OUT *g, unsigned nsym, OUT scale,
                                            int half, int tid) {
  if (sizeof(OUT) == 2) {
    uint16_t *gs = reinterpret_cast<uint16_t *>(g);
    if ((((uintptr_t)gs) & 15) == 0) {
      const uint4 *s4 = reinterpret_cast<const uint4 *>(s_out);
      uint4 *g4 = reinterpret_cast<uint4 *>(gs);
      const unsigned n16 = nsym / 8;
      for (unsigned k = tid; k < n16; k += NT)
        __stcs(g4 + k, s4[k]);
      for (unsigned k = n16 * 8 + tid; k < nsym; k += NT)
        gs[k] = s_out[k];
    } else {
      for (unsigned k = tid; k < nsym; k += NT)
        gs[k] = s_out[k];
    }
  } else {
    constexpr int V = 16 / sizeof(OUT); // values per 128-bit store
    if ((((uintptr_t)g) & 15) == 0) {
      const unsigned nv = nsym / V;
      for (unsigned k = tid; k < nv; k += NT) {
        __align__(16) OUT v[V];
#pragma unroll
        for (int j = 0; j < V; j++)
          v[j] = scale * (OUT)((long long)s_out[k * V + j] - half);
        __stcs(reinterpret_cast<uint4 *>(g) + k, *reinterpret_cast<const uint4 *>(v));
      }
      for (unsigned k = nv * V + tid; k < nsym; k += NT)
        g[k] = scale * (OUT)((long long)s_out[k] - half);
    } else {
      for (unsigned k = tid; k < nsym; k += NT)
        g[k] = scale * (OUT)((long long)s_out[k] - half);
    }
  }
}

template <bool STAGE_OUT, typename OUT>
__global__ void __launch_bounds__(DEC_T)
decode_kernel(const u64 *__restrict__ ddata, u64 total_words, const u64 *__restrict__ bits,
              const u64 *__restrict__ woff, u64 nchunk, int chunk, u64 n,
              const u64 *__restrict__ decodebook, int dict, unsigned *__restrict__ sub_start,
              unsigned *__restrict__ sub_end, unsigned *__restrict__ sub_cnt,
              OUT *__restrict__ out, OUT scale, const unsigned *__restrict__ n_skipped,
              const unsigned *__restrict__ skipped, unsigned fast_bufw) {
  // fast_bufw != 0: decode_fast_kernel ran first and listed the chunks it left
  // (more than fast_bufw words); usually there are few or none
  const u64 nwork = fast_bufw ? (u64)*n_skipped : nchunk;
  if (blockIdx.x >= nwork)
    return;
  extern __shared__ u64 s_db[]; // first[64] entry[64] | lut | keys16 | outbuf
  u64 *s_first = s_db, *s_entry = s_db + 64;
  unsigned *s_lut = (unsigned *)(s_db + 128);
  uint16_t *s_keys = (uint16_t *)(s_lut + (1 << DEC_K));
  uint16_t *s_out = s_keys + ((dict + 7) & ~7);
  __shared__ unsigned s_scan[DEC_T / 32];
  __shared__ unsigned s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 128; i += DEC_T)
    s_db[i] = decodebook[i];
  for (int i = tid; i < dict; i += DEC_T)
    s_keys[i] = (uint16_t)decodebook[128 + i];
  __syncthreads();
  int lmin = 1;
  while (lmin < 63 && s_first[lmin] == ~0ull)
    lmin++;
  for (int x = tid; x < (1 << DEC_K); x += DEC_T) {
    unsigned e = 0;
    for (int l = lmin; l <= DEC_K; l++) {
      u64 v = (u64)x >> (DEC_K - l);
      if (v >= s_first[l]) {
        u64 ki = s_entry[l] + v - s_first[l];
        e = (ki < (u64)dict ? (unsigned)s_keys[ki] : 0u) | ((unsigned)l << 16);
        break;
      }
    }
    s_lut[x] = e;
  }
  __syncthreads();
  DecTables t;
  t.first = s_first;
  t.entry = s_entry;
  t.lut = s_lut;
  t.keys = s_keys;
  t.dict = dict;
  t.lslow = max(lmin, DEC_K + 1);

  for (u64 wk = blockIdx.x; wk < nwork; wk += gridDim.x) {
    const u64 c = fast_bufw ? (u64)skipped[wk] : wk;
    const u64 B64 = bits[c];
    const u64 w0 = woff[c];
    const u64 nw = (B64 - 1) / 64 + 1;
    // the per-chunk fields come from the stream: a chunk that does not lie inside the
    // bit stream (or is longer than `chunk` codewords of 64 bits can be) decodes to zeros
    if (B64 == 0 || B64 > (u64)chunk * 64 || w0 > total_words || nw > total_words - w0) {
      const u64 cnt_ = min((u64)chunk, n - c * (u64)chunk);
      for (u64 i = tid; i < cnt_; i += DEC_T)
        out[c * (u64)chunk + i] = (OUT)0;
      continue;
    }
    const u64 *src = ddata + w0;
    const unsigned B = (unsigned)B64;
    const unsigned NS = (B + DEC_SB - 1) / DEC_SB;
    const u64 sbase = (w0 * 64) / DEC_SB + c;
    unsigned *st = sub_start + sbase, *en = sub_end + sbase, *cn = sub_cnt + sbase;
    const unsigned nsym = (unsigned)min((u64)chunk, n - c * (u64)chunk);
    // 1. speculative decode
    for (unsigned i = tid; i < NS; i += DEC_T) {
      unsigned cnt;
      unsigned s0 = i * DEC_SB;
      unsigned e = decode_sub<false>(t, src, nw, s0, min(B, s0 + DEC_SB), cnt, nullptr, 0);
      st[i] = s0;
      en[i] = e;
      cn[i] = cnt;
    }
    __syncthreads();
    // 2. synchronise
    for (unsigned round = 0; round < NS; round++) {
      int changed = 0;
      for (unsigned i = tid; i < NS; i += DEC_T) {
        if (i == 0)
          continue;
        unsigned s0 = en[i - 1];
        if (s0 != st[i]) {
          unsigned cnt = 0;
          unsigned lim = min(B, (i + 1) * DEC_SB);
          unsigned e = s0 >= lim ? s0 : decode_sub<false>(t, src, nw, s0, lim, cnt, nullptr, 0);
          st[i] = s0;
          en[i] = e;
          cn[i] = cnt;
          changed = 1;
        }
      }
      if (!__syncthreads_or(changed))
        break;
    }
    // 3 + 4. offsets and final decode
    if (tid == 0)
      s_carry = 0;
    __syncthreads();
    // unstaged output is only instantiated for OUT = symbols
    uint16_t *dst_base = STAGE_OUT ? s_out : reinterpret_cast<uint16_t *>(out) + c * (u64)chunk;
    for (unsigned i0 = 0; i0 < NS; i0 += DEC_T) {
      unsigned i = i0 + tid;
      unsigned cnt = i < NS ? cn[i] : 0;
      unsigned x = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
          x += y;
      }
      if (lane == 31)
        s_scan[wid] = x;
      __syncthreads();
      unsigned wpre = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < DEC_T / 32; k++) {
        unsigned v = s_scan[k];
        if (k < wid)
          wpre += v;
        tot += v;
      }
      unsigned off = s_carry + wpre + x - cnt;
      if (i < NS && cnt) {
        unsigned dummy;
        unsigned s0 = st[i];
        unsigned cap = off < nsym ? nsym - off : 0;
        decode_sub<true>(t, src, nw, s0, min(B, (i + 1) * DEC_SB), dummy, dst_base + off, cap);
      }
      __syncthreads();
      if (tid == 0)
        s_carry += tot;
      __syncthreads();
    }
    if (STAGE_OUT) {
      flush_chunk<OUT, DEC_T>(s_out, out + c * (u64)chunk, nsym, scale, dict / 2, tid);
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
// Fast decoder: same self-synchronising scheme, but the whole chunk (bit
// stream, sub-sequence bookkeeping, output) lives in shared memory and every
// thread owns a CONTIGUOUS run of K sub-sequences, so
//   * only the first sub-sequence of a thread starts speculatively; the rest of
//     its run continues from the true previous end (one bit reader, no restarts);
//   * a synchronisation round only re-decodes from the predecessor's end until
//     the new parse lands on a recorded sub-sequence end (typically 1-2
//     sub-sequences), and most threads are final after one round;
//   * the last pass decodes each thread's run once more, writing symbols to a
//     shared staging buffer that is flushed with 128-bit stores.
// The bit reader keeps >= 33 valid bits in a 64-bit register refilled 32 bits
// at a time; codes up to DEC_K bits resolve through the LUT, longer ones start
// the canonical walk at the shortest length their DEC_K-bit prefix allows.
// Chunks that do not fit the shared buffers are counted in *n_skipped and left
// to decode_kernel.
constexpr int DF_T = 512;   // threads per block, first launch (two blocks per SM)
constexpr int DF_TBIG = 1024; // second launch (one block per SM, large buffers)
// sub-sequence length of the fast decoder: 5 words, and every thread owns an ODD
// number of them, so the threads of a warp read the bit stream at an odd word
// stride (no shared-memory bank conflicts)
constexpr int DF_SB = 160;

__device__ __forceinline__ unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds32_4(unsigned addr) { // word at addr + 4
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Shared-memory tables of the fast decoder, addressed with 32-bit shared
// addresses (no generic-pointer arithmetic in the decode loop).
struct FastTables {
  unsigned words; // chunk bit stream as 32-bit halves in STREAM order, zero padded
  unsigned lut;   // 2^DEC_K x u32: bit31 clear: symbol | len << 16; set: walk start << 16
  unsigned t32;   // [34] first[l] left aligned in 32 bits (0xffffffff: no code of length l)
  unsigned b32;   // [34] entry[l] - first[l] (mod 2^32): key index = b32[l] + code
  const u64 *first, *entry; // generic pointers to the 64-entry tables (codes > 32 bits)
  const u64 *keys;          // global (decodebook + 128)
  int dict;
};

// 32 stream bits starting at bit position p
__device__ __forceinline__ unsigned fast_window(const FastTables &t, unsigned p) {
  const unsigned a = t.words + ((p >> 3) & ~3u);
  return __funnelshift_l(lds32_4(a), lds32(a), p); // shift = p & 31
}

// codes longer than 32 bits (or a walk of more than two steps): canonical walk
template <bool WANT_SYM>
__device__ __noinline__ unsigned fast_decode_slow(const FastTables &t, unsigned p, int l,
                                                  unsigned &sym) {
  const u64 win = ((u64)fast_window(t, p) << 32) | fast_window(t, p + 32);
  u64 v = win >> (64 - l);
  while (v < t.first[l] && l < 63) {
    l++;
    v = win >> (64 - l);
  }
  if (WANT_SYM) {
    const u64 ki = t.entry[l] + v - t.first[l];
    sym = ki < (u64)t.dict ? (unsigned)__ldg(t.keys + ki) : 0u;
  }
  return (unsigned)l;
}

// length of the codeword whose first 32 bits are `hi` (and its symbol)
template <bool WANT_SYM>
__device__ __forceinline__ unsigned fast_decode_one(const FastTables &t, unsigned hi, unsigned p,
                                                    unsigned &sym) {
  const unsigned e = lds32(t.lut + ((hi >> (32 - DEC_K)) << 2));
  unsigned l = (e >> 16) & 0xffu;
  if (e & 0x80000000u) {
    // longer than DEC_K bits: l is the shortest length this prefix allows; two
    // predicated steps of the canonical walk cover nearly every code
    const unsigned ta = t.t32 + (l << 2);
    const unsigned t0 = lds32(ta), t1 = lds32_4(ta);
    const bool ok0 = hi >= t0, ok1 = hi >= t1;
    if (l >= 31 || !(ok0 || ok1))
      return fast_decode_slow<WANT_SYM>(t, p, (int)l, sym);
    l += ok0 ? 0u : 1u;
    if (WANT_SYM) {
      const unsigned ki = lds32(t.b32 + (l << 2)) + (hi >> (32 - l));
      sym = ki < (unsigned)t.dict ? (unsigned)__ldg(reinterpret_cast<const unsigned short *>(t.keys + ki)) : 0u;
    }
    return l;
  }
  sym = e & 0xffffu;
  return l;
}

template <int NT, typename OUT>
__global__ void __launch_bounds__(NT, NT <= 512 ? 2 : 1)
decode_fast_kernel(const u64 *__restrict__ ddata, u64 total_words, const u64 *__restrict__ bits,
                   const u64 *__restrict__ woff, u64 nchunk, int chunk, u64 n,
                   const u64 *__restrict__ decodebook, int dict, unsigned bufw,
                   const unsigned *__restrict__ n_work, const unsigned *__restrict__ work,
                   unsigned *__restrict__ n_skipped, unsigned *__restrict__ skipped,
                   OUT *__restrict__ out, OUT scale) {
  // work == nullptr: every chunk; otherwise the *n_work chunks listed in work[]
  // (those an earlier launch with a smaller buffer left over)
  const u64 nwork = work ? (u64)*n_work : nchunk;
  if (blockIdx.x >= nwork)
    return;
  extern __shared__ __align__(16) unsigned char df_smem[];
  u64 *s_first = reinterpret_cast<u64 *>(df_smem), *s_entry = s_first + 64;
  unsigned *s_lut = reinterpret_cast<unsigned *>(s_first + 128);
  u64 *s_words = reinterpret_cast<u64 *>(s_lut + (1 << DEC_K)); // bufw + 2 (zero padding)
  const unsigned nslot = (bufw * 64 / DF_SB + 2 * NT + 3) & ~3u; // sub-sequence slots (keeps s_out 16-byte aligned)
  unsigned *s_en = reinterpret_cast<unsigned *>(s_words + bufw + 2);
  unsigned *s_st = s_en + nslot;
  unsigned char *s_cn = reinterpret_cast<unsigned char *>(s_st + NT);
  uint16_t *s_out = reinterpret_cast<uint16_t *>(s_cn + ((nslot + 15) & ~15u));
  __shared__ unsigned s_scan[NT / 32];
  __shared__ unsigned s_t32[36], s_b32[36];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  for (int i = tid; i < 128; i += NT)
    s_first[i] = decodebook[i];
  __syncthreads();
  if (tid < 36) {
    const bool valid = tid >= 1 && tid <= 32 && s_first[tid] != ~0ull && (s_first[tid] >> tid) == 0;
    s_t32[tid] = valid ? (unsigned)(s_first[tid] << (32 - tid)) : 0xffffffffu;
    s_b32[tid] = valid ? (unsigned)(s_entry[tid] - s_first[tid]) : 0u;
  }
  int lmin = 1;
  while (lmin < 63 && s_first[lmin] == ~0ull)
    lmin++;
  for (int x = tid; x < (1 << DEC_K); x += NT) {
    unsigned e = 0xffffffffu;
    for (int l = lmin; l <= DEC_K; l++) {
      const u64 v = (u64)x >> (DEC_K - l);
      if (v >= s_first[l]) {
        const u64 ki = s_entry[l] + v - s_first[l];
        e = (ki < (u64)dict ? (unsigned)(decodebook[128 + ki] & 0xffffu) : 0u) | ((unsigned)l << 16);
        break;
      }
    }
    if (e == 0xffffffffu) {
      // longer than DEC_K bits: shortest length this prefix can still carry
      int l = max(lmin, DEC_K + 1);
      while (l < 63) {
        const u64 vmax = (((u64)x + 1) << (l - DEC_K)) - 1;
        if (s_first[l] != ~0ull && vmax >= s_first[l])
          break;
        l++;
      }
      e = 0x80000000u | ((unsigned)l << 16);
    }
    s_lut[x] = e;
  }
  __syncthreads();
  // shared-memory addresses as opaque registers: otherwise the compiler
  // rematerialises every base (S2R + LEA on the CTA's shared window) inside the
  // decode loop instead of keeping four registers alive
  auto opaque = [](unsigned x) {
    asm volatile("mov.u32 %0, %0;" : "+r"(x));
    return x;
  };
  FastTables t;
  t.words = opaque((unsigned)__cvta_generic_to_shared(s_words));
  t.lut = opaque((unsigned)__cvta_generic_to_shared(s_lut));
  t.t32 = opaque((unsigned)__cvta_generic_to_shared(s_t32));
  t.b32 = opaque((unsigned)__cvta_generic_to_shared(s_b32));
  t.first = s_first;
  t.entry = s_entry;
  t.keys = decodebook + 128;
  t.dict = dict;
  const unsigned a_en = (unsigned)__cvta_generic_to_shared(s_en);
  const unsigned a_cn = (unsigned)__cvta_generic_to_shared(s_cn);

  for (u64 wk = blockIdx.x; wk < nwork; wk += gridDim.x) {
    const u64 c = work ? (u64)work[wk] : wk;
    const u64 B64 = bits[c];
    const u64 nw64 = (B64 - 1) / 64 + 1;
    // too large for this launch's buffer, or fields (they come from the stream) that
    // point outside the bit stream: left to decode_kernel, which rejects the latter
    if (nw64 > bufw || B64 == 0 || woff[c] > total_words || nw64 > total_words - woff[c]) {
      if (tid == 0)
        skipped[atomicAdd(n_skipped, 1u)] = (unsigned)c;
      continue;
    }
    const unsigned B = (unsigned)B64, nw = (unsigned)nw64;
    const u64 *src = ddata + woff[c];
    const unsigned nsym = (unsigned)min((u64)chunk, n - c * (u64)chunk);
    // 0. chunk words -> shared memory (all copies in flight), then every thread
    //    swaps the halves of its own words into stream order
    for (unsigned i = tid; i < nw; i += NT) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(s_words + i);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;\n" ::);
    if (tid < 2)
      s_words[nw + tid] = 0; // the reader may look up to two words past the end
    // run geometry: K (odd) sub-sequences of DF_SB bits per thread
    const unsigned NS = (B + DF_SB - 1) / DF_SB;
    const unsigned K = ((NS + NT - 1) / NT) | 1u;
    const unsigned run0 = tid * K * DF_SB; // first bit of this thread's run
    const bool active = run0 < B;
    asm volatile("cp.async.wait_group 0;\n" ::);
    for (unsigned i = tid; i < nw; i += NT) {
      const uint2 w = reinterpret_cast<uint2 *>(s_words)[i];
      reinterpret_cast<uint2 *>(s_words)[i] = make_uint2(w.y, w.x);
    }
    __syncthreads();
    // 1. speculative pass over the run
    if (active) {
      unsigned p = run0;
      for (unsigned j = 0; j < K; j++) {
        const unsigned i = tid * K + j;
        const unsigned lim = min(B, (i + 1) * DF_SB);
        unsigned cnt = 0;
        while (p < lim) {
          unsigned sym;
          p += fast_decode_one<false>(t, fast_window(t, p), p, sym);
          cnt++;
        }
        s_en[i] = p;
        s_cn[i] = (unsigned char)cnt;
      }
      s_st[tid] = run0;
    }
    __syncthreads();
    // 2. synchronise: restart from the predecessor's end until the parse meets
    //    a recorded sub-sequence end
    for (unsigned round = 0; round <= NT; round++) {
      int changed = 0;
      if (active && tid > 0) {
        const unsigned s0 = s_en[tid * K - 1];
        if (s0 != s_st[tid]) {
          changed = 1;
          s_st[tid] = s0;
          unsigned p = s0;
          for (unsigned j = 0; j < K; j++) {
            const unsigned i = tid * K + j;
            const unsigned lim = min(B, (i + 1) * DF_SB);
            unsigned cnt = 0;
            while (p < lim) {
              unsigned sym;
              p += fast_decode_one<false>(t, fast_window(t, p), p, sym);
              cnt++;
            }
            const bool met = (p == s_en[i]);
            s_en[i] = p;
            s_cn[i] = (unsigned char)cnt;
            if (met)
              break;
          }
        }
      }
      if (!__syncthreads_or(changed))
        break;
    }
    // 3. output offsets (block scan of the per-run symbol counts) + final pass
    unsigned cnt = 0;
    if (active)
      for (unsigned j = 0; j < K; j++)
        cnt += lds8(a_cn + tid * K + j);
    unsigned x = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o)
        x += y;
    }
    if (lane == 31)
      s_scan[wid] = x;
    __syncthreads();
    unsigned wpre = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++)
      if (k < wid)
        wpre += s_scan[k];
    unsigned off = wpre + x - cnt;
    if (active && cnt) {
      unsigned p = tid ? lds32(a_en + ((tid * K - 1) << 2)) : 0u;
      const unsigned lim = min(B, (tid + 1) * K * DF_SB);
      const unsigned a_out = (unsigned)__cvta_generic_to_shared(s_out);
      while (p < lim) {
        unsigned sym;
        p += fast_decode_one<true>(t, fast_window(t, p), p, sym);
        if (off < nsym)
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_out + (off << 1)), "h"((unsigned short)sym));
        off++;
      }
    }
    __syncthreads();
    flush_chunk<OUT, NT>(s_out, out + c * (u64)chunk, nsym, scale, dict / 2, tid);
    __syncthreads();
  }
}

// decodebook_size, ddata word count and outlier count of a serialised block
// (Huffman.hpp:293-312); ~0 marks a truncated stream.
__global__ void parse_sizes_kernel(const unsigned char *__restrict__ p, u64 off, u64 size,
                                   u64 dict, u64 *__restrict__ dst) {
  dst[0] = dst[1] = dst[2] = ~0ull;
  if (off + 8 > size)
    return;
  u64 db = *(const u64 *)(p + off);
  dst[0] = db;
  if (db != 1024 + 8 * dict || off + 8 + db + 8 > size)
    return;
  u64 o2 = off + 8 + db;
  u64 tw = *(const u64 *)(p + o2);
  if (tw > (size - o2 - 8) / 8 || o2 + 8 + 8 * tw + 8 > size)
    return;
  dst[1] = tw;
  dst[2] = *(const u64 *)(p + o2 + 8 + 8 * tw);
}

} // namespace
#include "huffman_serial.cuh"
namespace {

// inputs with at least this many chunks take the chunk-serial kernels
// (huffman_serial.cuh); MGB_SERIAL_MIN_CHUNKS overrides (0: always, huge: never)
u64 g_serial_min_chunks = getenv("MGB_SERIAL_MIN_CHUNKS") ? strtoull(getenv("MGB_SERIAL_MIN_CHUNKS"), nullptr, 10)
                                                          : 16384ull;
u64 serial_min_chunks() { return g_serial_min_chunks; }
// the ring formulation of the thread-per-chunk decoder (MGB_TUNE_RING_DECODER; 0: first formulation)
bool g_ring_decoder = !getenv("MGB_NO_RING_DECODER");
bool g_sub_encoder = !getenv("MGB_NO_SUB_ENCODER");
int g_ring_lanes = getenv("MGB_RING_LANES") ? std::min(32, std::max(1, atoi(getenv("MGB_RING_LANES")))) : 32;

// Launches the decoders for one serialised block.  OUT = uint16_t: symbols;
// OUT = float / double: values dequantized with `scale` while a chunk is flushed.
template <typename OUT>
int launch_decoders(mgb_plan *p, const u64 *ddata, u64 total_words, const u64 *bits,
                    const u64 *woff, u64 nchunk, int chunk, u64 n, const u64 *decodebook, int dict,
                    OUT *out, OUT scale, cudaStream_t st) {
  const bool out_vec = ((uintptr_t)out & 31) == 0 && ((size_t)chunk * sizeof(OUT)) % 32 == 0;
  if (nchunk >= serial_min_chunks() && g_ring_decoder && out_vec && dict <= 65536 &&
      serial::ring_smem_bytes<OUT>(dict, 4096, 64) <= 200 * 1024) {
    // thread per chunk, stream through a shared-memory ring (huffman_serial.cuh)
    const size_t budget = 200 * 1024;
    if (!p->d_declut)
      MGB_CUDA_CHECK(cudaMalloc(&p->d_declut, std::max(serial::ring_tab_bytes<float>(dict, serial::RL2_MAX_BYTES),
                                                       serial::tab_bytes(dict))));
    static bool configured[64] = {};
    if (mgb_first_use_on_device(configured))
      cudaFuncSetAttribute(serial::decode_ring_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    // Chunks per block: the chunks of one wave spread evenly over the SMs, k blocks on each,
    // for the smallest k whose blocks fit (tables + 128 bytes of ring per chunk); what is left
    // of the shared memory goes to second-level tables.
    const unsigned L = (unsigned)g_ring_lanes; // lanes of a warp that take a chunk
    cudaFuncAttributes fa;
    MGB_CUDA_CHECK(cudaFuncGetAttributes(&fa, serial::decode_ring_kernel<OUT>));
    const unsigned max_thr = std::min(2048u, 65536u / (unsigned)std::max(fa.numRegs, 32) / 32 * 32); // per SM
    const unsigned act_max = serial::RING_T / 32 * L;
    const size_t sub_min = 16 * 1024; // second-level tables a block should have at least
    unsigned act = act_max, resident = 0;      // chunks per block | blocks per SM
    for (unsigned k = 1; k <= 8; k++) {
      unsigned t = (unsigned)((nchunk + 148ull * k - 1) / (148ull * k));
      t = std::max(2 * L, (t + L - 1) / L * L);
      if (t <= act_max && t / L * 32 * k <= max_thr &&
          budget / serial::ring_smem_bytes<OUT>(dict, sub_min, (int)t) >= k) {
        act = t;
        resident = k;
        break;
      }
    }
    const size_t base = serial::ring_smem_bytes<OUT>(dict, 0, (int)act);
    if (base + 4096 > budget)
      return MGB_FAILURE;
    // blocks per SM (several waves: as many as fit with the least tables)
    const size_t per_sm = resident ? resident
                                   : std::max<size_t>(1, std::min<size_t>(budget / (base + sub_min), max_thr / (act / L * 32)));
    const unsigned sub_bytes = (unsigned)(std::min<size_t>(serial::RL2_MAX_BYTES, budget / per_sm - base) & ~(size_t)127);
    // warps work in groups of RING_S on L * RING_S chunks: whole groups
    const unsigned threads = act / L * 32;
    const u64 warps = (nchunk + (u64)L * serial::RING_S - 1) / ((u64)L * serial::RING_S) * serial::RING_S;
    const unsigned blocks = (unsigned)((warps + threads / 32 - 1) / (threads / 32));
    MGB_LAUNCH(MGB_K_PARSE, st,
               (serial::build_ring_lut_kernel<OUT><<<1, 1024, 0, st>>>(decodebook, dict, sub_bytes,
                                                                       (unsigned char *)p->d_declut, scale)));
    MGB_LAUNCH(MGB_K_DECODE, st,
               (serial::decode_ring_kernel<OUT><<<blocks, threads, serial::ring_smem_bytes<OUT>(dict, sub_bytes, (int)act), st>>>(
                   ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict, sub_bytes, (int)L,
                   (const unsigned char *)p->d_declut, out, scale)));
    MGB_CUDA_CHECK(cudaGetLastError());
    return MGB_SUCCESS;
  }
  if (nchunk >= serial_min_chunks() && serial::tab_bytes(dict) <= 200 * 1024) {
    // thread per chunk (huffman_serial.cuh)
    const size_t tabb = serial::tab_bytes(dict);
    if (!p->d_declut)
      MGB_CUDA_CHECK(cudaMalloc(&p->d_declut, std::max(tabb, serial::ring_tab_bytes<float>(dict, serial::RL2_MAX_BYTES))));
    static bool configured[64] = {};
    if (mgb_first_use_on_device(configured)) {
      cudaFuncSetAttribute(serial::decode_serial_kernel<OUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(serial::decode_serial_kernel<OUT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    MGB_LAUNCH(MGB_K_PARSE, st, (serial::build_lut_kernel<<<1, 1024, 0, st>>>(decodebook, dict, (unsigned char *)p->d_declut)));
    // Threads (= chunks) per block so that the chunks spread evenly over the SMs: the
    // tables allow two blocks per SM, and a launch is as slow as its fullest SM (52 700
    // chunks of a C5 slab in blocks of 256: 58 SMs with 512 threads, 90 with 256)
    const u64 slots = 148ull * 2;
    const u64 waves = (nchunk + slots * serial::DS_T - 1) / (slots * serial::DS_T);
    unsigned threads = (unsigned)((nchunk + slots * waves - 1) / (slots * waves));
    threads = std::min<unsigned>(serial::DS_T, std::max<unsigned>(64, (threads + 31) & ~31u));
    const unsigned blocks = (unsigned)((nchunk + threads - 1) / threads);
    const bool vec = ((uintptr_t)out & 31) == 0 && ((size_t)chunk * sizeof(OUT)) % 32 == 0;
    if (vec)
      MGB_LAUNCH(MGB_K_DECODE, st,
                 (serial::decode_serial_kernel<OUT, true><<<blocks, threads, tabb, st>>>(
                     ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict,
                     (const unsigned char *)p->d_declut, out, scale)));
    else
      MGB_LAUNCH(MGB_K_DECODE, st,
                 (serial::decode_serial_kernel<OUT, false><<<blocks, threads, tabb, st>>>(
                     ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict,
                     (const unsigned char *)p->d_declut, out, scale)));
    MGB_CUDA_CHECK(cudaGetLastError());
    return MGB_SUCCESS;
  }
  unsigned *sub = p->d_dec_sub;
  const u64 subn = p->dec_sub_cap;
  // fast path: chunks staged in shared memory.  First launch: 2 blocks of 512
  // threads per SM, word buffer sized from the average chunk (+25 %); second
  // launch (1 block per SM, the largest buffer that fits) for the chunks the
  // first one left; whatever still does not fit goes to decode_kernel.
  unsigned fast_bufw = 0, big_bufw = 0;
  unsigned *cnt1 = (unsigned *)((u64 *)p->d_scalars + 15), *cnt2 = cnt1 + 1;
  unsigned *list1 = sub + 3 * subn, *list2 = list1 + nchunk;
  auto fast_smem = [&](unsigned bw, int nt) {
    const unsigned nslot = (bw * 64 / DF_SB + 2 * nt + 3) & ~3u;
    return 128 * 8 + (size_t)(1 << DEC_K) * 4 + ((size_t)bw + 2) * 8 + (size_t)nslot * 4 +
           (size_t)nt * 4 + ((nslot + 15) & ~15u) + (((size_t)chunk * 2 + 15) & ~(size_t)15);
  };
  // largest word buffer whose launch fits `budget` bytes of shared memory
  // (per word: 8 B data + 64/DF_SB x (4 B end slot + 1 B count))
  auto max_words = [&](size_t budget, int nt) -> u64 {
    const size_t fixed = fast_smem(0, nt) + 64;
    return fixed >= budget ? 0 : (budget - fixed) / 10;
  };
  {
    const u64 avg = total_words / nchunk + 1;
    const u64 small_max = max_words(113 * 1024 - 512, DF_T);
    const u64 big_max = max_words(200 * 1024, DF_TBIG);
    u64 want = avg + avg / 4 + 256;
    if (want > small_max && avg + avg / 16 + 32 <= small_max)
      want = small_max;
    if (want <= small_max)
      fast_bufw = (unsigned)(want & ~(u64)1);
    if (big_max > 64) {
      big_bufw = (unsigned)(std::min<u64>(big_max, (u64)chunk * 56 / 64 + 2) & ~(u64)1);
      if (big_bufw <= fast_bufw)
        big_bufw = 0;
    }
  }
  if (fast_bufw || big_bufw) {
    static bool configured[64] = {};
    if (mgb_first_use_on_device(configured)) {
      MGB_CUDA_CHECK(cudaFuncSetAttribute(decode_fast_kernel<DF_T, OUT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
      cudaFuncSetAttribute(decode_fast_kernel<DF_T, OUT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      MGB_CUDA_CHECK(cudaFuncSetAttribute(decode_fast_kernel<DF_TBIG, OUT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      cudaFuncSetAttribute(decode_fast_kernel<DF_TBIG, OUT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
    }
    MGB_CUDA_CHECK(cudaMemsetAsync(cnt1, 0, 2 * sizeof(unsigned), st));
  }
  if (fast_bufw) {
    unsigned fblocks = (unsigned)std::min<u64>(nchunk, 148 * 2);
    MGB_LAUNCH(MGB_K_DECODE, st,
               (decode_fast_kernel<DF_T, OUT><<<fblocks, DF_T, fast_smem(fast_bufw, DF_T), st>>>(
                   ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict, fast_bufw, nullptr,
                   nullptr, cnt1, list1, out, scale)));
  }
  if (big_bufw) {
    unsigned fblocks = (unsigned)std::min<u64>(nchunk, 148);
    MGB_LAUNCH(MGB_K_DECODE, st,
               (decode_fast_kernel<DF_TBIG, OUT><<<fblocks, DF_TBIG, fast_smem(big_bufw, DF_TBIG), st>>>(
                   ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict, big_bufw,
                   fast_bufw ? cnt1 : nullptr, fast_bufw ? list1 : nullptr, cnt2, list2, out, scale)));
  }
  // what decode_kernel has to look at: the second list, else the first
  unsigned *n_skipped = big_bufw ? cnt2 : cnt1;
  unsigned *skip_list = big_bufw ? list2 : list1;
  if (big_bufw)
    fast_bufw = big_bufw;
  size_t smem_tab = 128 * 8 + (size_t)(1 << DEC_K) * 4 + (size_t)((dict + 7) & ~7) * 2;
  size_t smem_out = (size_t)chunk * 2 + 16;
  unsigned blocks = (unsigned)std::min<u64>(nchunk, 148 * 8);
  if (smem_tab + smem_out <= 160 * 1024) {
    size_t smem = smem_tab + smem_out;
    cudaFuncSetAttribute(decode_kernel<true, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    MGB_LAUNCH(MGB_K_DECODE, st,
               (decode_kernel<true, OUT><<<blocks, DEC_T, smem, st>>>(
                   ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict, sub,
                   sub + subn, sub + 2 * subn, out, scale, n_skipped, skip_list, fast_bufw)));
  } else {
    if (sizeof(OUT) != 2)
      return MGB_FAILURE; // the caller only fuses the dequantizer when chunks can be staged
    cudaFuncSetAttribute(decode_kernel<false, uint16_t>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tab);
    MGB_LAUNCH(MGB_K_DECODE, st,
               (decode_kernel<false, uint16_t><<<blocks, DEC_T, smem_tab, st>>>(
                   ddata, total_words, bits, woff, nchunk, chunk, n, decodebook, dict, sub,
                   sub + subn, sub + 2 * subn, reinterpret_cast<uint16_t *>(out), (uint16_t)0,
                   n_skipped, skip_list, fast_bufw)));
  }
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

int ensure_huff_workspace(mgb_plan *p) {
  if (p->d_codebook)
    return MGB_SUCCESS;
  const int dict = p->cfg.huff_dict_size;
  const u64 nchunk = (p->N - 1) / p->cfg.huff_block_size + 1;
  int npow2 = 1;
  while (npow2 < dict)
    npow2 <<= 1;
  MGB_CUDA_CHECK(cudaMalloc(&p->d_codebook, dict * sizeof(u64)));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_decodebook, (128 + dict) * sizeof(u64)));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_chunk_bits, nchunk * sizeof(u64)));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_chunk_woff, (nchunk + 1) * sizeof(u64)));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_chunk_sub, nchunk * serial::ESUB * sizeof(unsigned)));
  MGB_CUDA_CHECK(cudaMalloc(&p->d_scalars, 24 * sizeof(u64))); // [16..17]: grid barrier of the outlier sort
  MGB_CUDA_CHECK(cudaMemset(p->d_scalars, 0, 24 * sizeof(u64)));
  size_t cbw = npow2 * sizeof(u64) + (size_t)dict * (9 * 4 + 8) + 256;
  MGB_CUDA_CHECK(cudaMalloc(&p->d_cbwork, cbw));
  MGB_CUDA_CHECK(cudaMallocHost(&p->h_pinned, 16 * sizeof(u64)));
  return MGB_SUCCESS;
}

} // namespace

int mgb_huff_workspace(mgb_plan *p) { return ensure_huff_workspace(p); }

extern "C" int mgb_tune(int key, long long value) {
  switch (key) {
  case MGB_TUNE_SERIAL_MIN_CHUNKS:
    g_serial_min_chunks = value < 0 ? ~0ull : (u64)value;
    return MGB_SUCCESS;
  case MGB_TUNE_RING_DECODER:
    g_ring_decoder = value != 0;
    return MGB_SUCCESS;
  case MGB_TUNE_SUB_ENCODER:
    g_sub_encoder = value != 0;
    return MGB_SUCCESS;
  default:
    return MGB_BAD_ARGUMENT;
  }
}

extern "C" int mgb_codebook(mgb_plan *p, const uint32_t *d_hist, uint64_t *d_codebook,
                            uint64_t *d_decodebook, void *stream) {
  if (!p || !d_hist || !d_codebook || !d_decodebook)
    return MGB_BAD_ARGUMENT;
  int rc = ensure_huff_workspace(p);
  if (rc)
    return rc;
  const int dict = p->cfg.huff_dict_size;
  int npow2 = 1;
  while (npow2 < dict)
    npow2 <<= 1;
  CbWork w;
  unsigned char *b = p->d_cbwork;
  w.keys_sorted = (u64 *)b; b += (size_t)npow2 * 8;
  w.cw = (u64 *)b; b += (size_t)dict * 8;
  w.lfreq = (unsigned *)b; b += (size_t)dict * 4;
  w.CL = (unsigned *)b; b += (size_t)dict * 4;
  w.lleader = (int *)b; b += (size_t)dict * 4;
  w.ifreq = (unsigned *)b; b += (size_t)dict * 4;
  w.ileader = (int *)b; b += (size_t)dict * 4;
  w.tfreq = (unsigned *)b; b += (size_t)dict * 4;
  w.tindex = (int *)b; b += (size_t)dict * 4;
  w.tleaf = (int *)b; b += (size_t)dict * 4;
  // shared memory: sort buffer (8 B x npow2) + 8 work arrays of ncap entries
  const size_t smem_max = 200 * 1024;
  const int key_smem = (size_t)npow2 * 8 <= 128 * 1024;
  size_t left = smem_max - (key_smem ? (size_t)npow2 * 8 : 0);
  int ncap = (int)std::min<size_t>(left / 32, (size_t)dict);
  const size_t smem = (key_smem ? (size_t)npow2 * 8 : 0) + (size_t)ncap * 32;
  static bool configured[64] = {};
  if (mgb_first_use_on_device(configured)) {
    MGB_CUDA_CHECK(cudaFuncSetAttribute(codebook_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_max));
  }
  static const int cb_threads = getenv("MGB_CB_THREADS") ? atoi(getenv("MGB_CB_THREADS")) : 1024;
  MGB_LAUNCH(MGB_K_CODEBOOK, (cudaStream_t)stream,
             (codebook_kernel<<<1, cb_threads, smem, (cudaStream_t)stream>>>(
                 d_hist, dict, npow2, w, (u64 *)d_codebook, (u64 *)d_decodebook,
                 (int *)(p->d_scalars + 8), key_smem, ncap)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

// Internal: everything up to (not including) the final size read-back.
// ocount is read from d_ocount_ptr on the device when non-null.
int mgb_huffman_compress_async(mgb_plan *p, const uint16_t *d_sym, uint64_t n,
                               const uint32_t *d_hist,
                               const unsigned long long *d_ocount_ptr,
                               uint64_t ocount_fixed, const uint64_t *d_oidx,
                               const int64_t *d_oval, uint8_t *d_out, uint64_t cap,
                               cudaStream_t st) {
  int rc = ensure_huff_workspace(p);
  if (rc)
    return rc;
  const int dict = p->cfg.huff_dict_size, chunk = p->cfg.huff_block_size;
  const u64 nchunk = (n - 1) / chunk + 1;
  rc = mgb_codebook(p, d_hist, (uint64_t *)p->d_codebook, (uint64_t *)p->d_decodebook, st);
  if (rc)
    return rc;
  const u64 fixed = 8 + 4 + 4 + 8 + 16 * nchunk + 8 + (1024 + 8ull * dict) + 8;
  u64 *scal = (u64 *)p->d_scalars;
  unsigned gb = (unsigned)std::min<u64>(nchunk, 148 * 8);
  // thread-per-chunk kernels, eight threads per chunk when the chunk divides that way
  // (huffman_serial.cuh); MGB_NO_SUB_ENCODER / MGB_TUNE_SUB_ENCODER = 0: one thread per chunk
  // (eight threads per chunk reach the thread count that pays with an eighth of the chunks)
  const bool serial_enc = nchunk >= serial_min_chunks();
  const bool sub_enc = g_sub_encoder && chunk >= 1024 && chunk % (8 * serial::ESUB) == 0 && (size_t)dict <= 65536 &&
                       (serial_enc || (serial_min_chunks() != ~0ull && nchunk * serial::ESUB >= serial_min_chunks()));
  if (sub_enc)
    MGB_LAUNCH(MGB_K_CHUNK_BITS, st,
               (serial::chunk_bits_sub_kernel<<<gb, 256, dict, st>>>(d_sym, n, chunk, p->d_codebook, dict,
                                                                     (u64 *)p->d_chunk_bits, p->d_chunk_sub)));
  else
    MGB_LAUNCH(MGB_K_CHUNK_BITS, st,
               (chunk_bits_kernel<<<gb, 256, dict, st>>>(d_sym, n, chunk, p->d_codebook, dict,
                                                        (u64 *)p->d_chunk_bits)));
  MGB_LAUNCH(MGB_K_CHUNK_SCAN, st,
             (chunk_scan_kernel<<<1, 1024, 0, st>>>((u64 *)p->d_chunk_bits, nchunk,
                                                   (u64 *)p->d_chunk_woff, scal, fixed, cap,
                                                   (const u64 *)d_ocount_ptr, ocount_fixed,
                                                   d_ocount_ptr ? p->outlier_cap : ~0ull)));
  u64 *ddata = (u64 *)(d_out + fixed);
  constexpr int TILE_WORDS = ENC_PER * 256 * 56 / 64 + 4;
  // the codebook is read through L1 (measured faster than a shared-memory copy,
  // which limits the kernel to two blocks per SM); MGB_ENC_SHARED_CB=1 for A/B runs
  static const bool enc_shared = getenv("MGB_ENC_SHARED_CB") != nullptr;
  if (sub_enc) {
    const unsigned blocks = (unsigned)((nchunk * serial::ESUB + serial::ESUB_T - 1) / serial::ESUB_T);
    const size_t smem = (size_t)dict * 8;
    if (smem <= 72 * 1024) {
      static bool configured[64] = {};
      if (mgb_first_use_on_device(configured))
        cudaFuncSetAttribute(serial::encode_sub_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
      MGB_LAUNCH(MGB_K_ENCODE, st,
                 (serial::encode_sub_kernel<true><<<blocks, serial::ESUB_T, smem, st>>>(
                     d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, p->d_chunk_sub, scal, ddata)));
    } else {
      MGB_LAUNCH(MGB_K_ENCODE, st,
                 (serial::encode_sub_kernel<false><<<blocks, serial::ESUB_T, 0, st>>>(
                     d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, p->d_chunk_sub, scal, ddata)));
    }
  } else if (serial_enc) {
    // thread per chunk (huffman_serial.cuh); codebook in shared memory when it fits
    const unsigned blocks = (unsigned)((nchunk + serial::ES_T - 1) / serial::ES_T);
    const bool vec = ((uintptr_t)d_sym & 31) == 0 && ((size_t)chunk * 2) % 32 == 0;
    const size_t smem = (size_t)dict * 8;
    static const bool cb_global = getenv("MGB_ENC_SERIAL_GLOBAL_CB") != nullptr;
    if (smem <= 72 * 1024 && !cb_global) {
      static bool configured[64] = {};
      if (mgb_first_use_on_device(configured)) {
        cudaFuncSetAttribute(serial::encode_serial_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        cudaFuncSetAttribute(serial::encode_serial_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
      }
      if (vec)
        MGB_LAUNCH(MGB_K_ENCODE, st,
                   (serial::encode_serial_kernel<true, true><<<blocks, serial::ES_T, smem, st>>>(
                       d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, scal, ddata)));
      else
        MGB_LAUNCH(MGB_K_ENCODE, st,
                   (serial::encode_serial_kernel<true, false><<<blocks, serial::ES_T, smem, st>>>(
                       d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, scal, ddata)));
    } else if (vec) {
      MGB_LAUNCH(MGB_K_ENCODE, st,
                 (serial::encode_serial_kernel<false, true><<<blocks, serial::ES_T, 0, st>>>(
                     d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, scal, ddata)));
    } else {
      MGB_LAUNCH(MGB_K_ENCODE, st,
                 (serial::encode_serial_kernel<false, false><<<blocks, serial::ES_T, 0, st>>>(
                     d_sym, n, chunk, p->d_codebook, dict, (u64 *)p->d_chunk_woff, scal, ddata)));
    }
  } else if (dict <= 16384 && enc_shared) {
    size_t smem = ((size_t)dict + TILE_WORDS) * 8;
    cudaFuncSetAttribute(encode_kernel<true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MGB_LAUNCH(MGB_K_ENCODE, st,
               (encode_kernel<true><<<gb, 256, smem, st>>>(d_sym, n, chunk, p->d_codebook, dict,
                                                          (u64 *)p->d_chunk_woff, scal, ddata)));
  } else {
    size_t smem = (size_t)TILE_WORDS * 8;
    MGB_LAUNCH(MGB_K_ENCODE, st,
               (encode_kernel<false><<<gb, 256, smem, st>>>(d_sym, n, chunk, p->d_codebook, dict,
                                                           (u64 *)p->d_chunk_woff, scal, ddata)));
  }
  MGB_LAUNCH(MGB_K_SERIALIZE, st,
             (serialize_meta_kernel<<<148, 256, 0, st>>>(
                 d_out, n, dict, chunk, nchunk, (u64 *)p->d_chunk_bits, (u64 *)p->d_chunk_woff,
                 p->d_decodebook, scal, fixed, (const u64 *)d_ocount_ptr, ocount_fixed, d_oidx,
                 (const i64 *)d_oval)));
  MGB_CUDA_CHECK(cudaGetLastError());
  return MGB_SUCCESS;
}

// reads back {need bytes, overflow flag, codebook status}; synchronises.
int mgb_huffman_finish(mgb_plan *p, uint64_t *size, cudaStream_t st) {
  MGB_CUDA_CHECK(cudaMemcpyAsync(p->h_pinned, p->d_scalars, 16 * sizeof(u64),
                                 cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  int cbstatus = (int)(p->h_pinned[8] & 0xffffffffu);
  if (cbstatus == 2)
    return MGB_FAILURE;
  *size = p->h_pinned[3];
  if (p->h_pinned[2] == 1)
    return MGB_OUTPUT_TOO_LARGE;
  if (p->h_pinned[2] == 2)
    return MGB_FAILURE; // outlier buffer overflow: see mgb_compress_lowlevel
  return MGB_SUCCESS;
}

extern "C" int mgb_huffman_compress(mgb_plan *p, const uint16_t *d_sym, uint64_t n,
                                    const uint32_t *d_hist, uint64_t ocount,
                                    const uint64_t *d_oidx, const int64_t *d_oval,
                                    uint8_t *d_out, uint64_t cap, uint64_t *size,
                                    void *stream) {
  if (!p || !d_sym || !d_hist || !d_out || !size || n == 0)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = mgb_huffman_compress_async(p, d_sym, n, d_hist, nullptr, ocount, d_oidx,
                                      d_oval, d_out, cap, st);
  if (rc)
    return rc;
  return mgb_huffman_finish(p, size, st);
}

// d_deq != nullptr: write values dequantized with deq_scale to d_deq instead of
// symbols to d_sym when possible (*fused = 1 then).
int mgb_huffman_decompress_impl(mgb_plan *p, const uint8_t *d_in, uint64_t size,
                                uint16_t *d_sym, uint64_t n, uint64_t *ocount,
                                const uint64_t **d_oidx, const int64_t **d_oval, void *stream,
                                void *d_deq, double deq_scale, int *fused) {
  if (fused)
    *fused = 0;
  if (!p || !d_in || !d_sym || size < 32)
    return MGB_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_huff_workspace(p);
  if (rc)
    return rc;
  // Huffman.hpp:264-320: scalar fields come back to the host to size the views
  unsigned char head[24];
  MGB_CUDA_CHECK(cudaMemcpyAsync(head, d_in, 24, cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  u64 primary_count, huffmeta;
  int dict, chunk;
  memcpy(&primary_count, head, 8);
  memcpy(&dict, head + 8, 4);
  memcpy(&chunk, head + 12, 4);
  memcpy(&huffmeta, head + 16, 8);
  if (primary_count != n || dict < 2 || dict > 65536 || chunk < 1)
    return MGB_BAD_STREAM;
  const u64 nchunk = (n - 1) / chunk + 1;
  if (huffmeta != 2 * nchunk)
    return MGB_BAD_STREAM;
  u64 off = 24;
  const u64 *bits = (const u64 *)(d_in + off);
  off += 8 * nchunk;
  const u64 *woff = (const u64 *)(d_in + off);
  off += 8 * nchunk;
  if (off + 8 > size)
    return MGB_BAD_STREAM;
  // remaining size fields are read on the device in one go
  u64 *tmp = (u64 *)p->d_scalars + 12;
  MGB_LAUNCH(MGB_K_PARSE, st, (parse_sizes_kernel<<<1, 1, 0, st>>>(d_in, off, size, (u64)dict, tmp)));
  u64 hs[3];
  MGB_CUDA_CHECK(cudaMemcpyAsync(hs, tmp, 24, cudaMemcpyDeviceToHost, st));
  MGB_CUDA_CHECK(cudaStreamSynchronize(st));
  const u64 dbsize = hs[0], total_words = hs[1], oc = hs[2];
  if (dbsize != 1024 + 8ull * dict || total_words == ~0ull || oc == ~0ull)
    return MGB_BAD_STREAM;
  off += 8;
  const u64 *decodebook = (const u64 *)(d_in + off);
  off += dbsize + 8;
  const u64 *ddata = (const u64 *)(d_in + off);
  // sizes come from the stream: no multiplication that could wrap
  if (total_words > (size - off) / 8 || size - off - 8 * total_words < 8)
    return MGB_BAD_STREAM;
  off += 8 * total_words + 8;
  if (oc > n || oc > (size - off) / 16)
    return MGB_BAD_STREAM;
  if (ocount)
    *ocount = oc;
  if (d_oidx)
    *d_oidx = (const uint64_t *)(d_in + off);
  if (d_oval)
    *d_oval = (const int64_t *)(d_in + off + 8 * oc);
  // scratch for the sub-sequence bookkeeping (3 x u32 per 128 stream bits) of the
  // block-per-chunk decoders
  if (nchunk < serial_min_chunks()) {
    u64 need = (total_words * 64) / DEC_SB + nchunk + 8;
    if (p->dec_sub_cap < need) {
      cudaFree(p->d_dec_sub);
      p->d_dec_sub = nullptr;
      p->dec_sub_cap = 0;
      MGB_CUDA_CHECK(cudaMalloc(&p->d_dec_sub, (need * 3 + 2 * nchunk) * sizeof(unsigned)));
      p->dec_sub_cap = need;
    }
  }
  // dequantize while flushing (s = inf) when the chunks can be staged in shared memory
  const bool can_stage = 128 * 8 + (size_t)(1 << DEC_K) * 4 + (size_t)((dict + 7) & ~7) * 2 +
                             (size_t)chunk * 2 + 16 <= 160 * 1024;
  if (d_deq && can_stage) {
    if (fused)
      *fused = 1;
    if (p->dtype == MGB_F32)
      return launch_decoders<float>(p, ddata, total_words, bits, woff, nchunk, chunk, n, decodebook,
                                    dict, (float *)d_deq, (float)deq_scale, st);
    return launch_decoders<double>(p, ddata, total_words, bits, woff, nchunk, chunk, n, decodebook,
                                   dict, (double *)d_deq, deq_scale, st);
  }
  return launch_decoders<uint16_t>(p, ddata, total_words, bits, woff, nchunk, chunk, n, decodebook,
                                   dict, d_sym, (uint16_t)0, st);
}

extern "C" int mgb_huffman_decompress(mgb_plan *p, const uint8_t *d_in, uint64_t size,
                                      uint16_t *d_sym, uint64_t n, uint64_t *ocount,
                                      const uint64_t **d_oidx, const int64_t **d_oval,
                                      void *stream) {
  return mgb_huffman_decompress_impl(p, d_in, size, d_sym, n, ocount, d_oidx, d_oval, stream,
                                     nullptr, 0.0, nullptr);
}
